// qz_mcts.cu -- K5-K8: batched PUCT MCTS over many concurrent games, trees in flat device arrays.
//
// Reference semantics restated (file:line into the reference):
//   TreeNode            mcts.py:12-80  (dup pure_mcts.py:19-56)   -> one SoA slot per node (= per edge)
//   MCTS._playout       mcts.py:103-127 / pure_mcts.py:66-83      -> select kernel + expand_backup kernel
//   get_move_probs      mcts.py:129-144                           -> root_stats kernel
//   update_with_move    mcts.py:146-151                           -> reroot kernel (subtree compaction)
//   MCTSPlayer.choose_action  mcts.py:172-196                     -> choose kernel
//
// Layout: game g owns the node slots [g*node_cap, (g+1)*node_cap).  A node's children are CONTIGUOUS and
// stored in the reference's actions() order, so a warp scans them with coalesced loads and "first max wins"
// (mcts.py:42) is "lowest child index wins".  Per node: prior f32, visits i32, Q f64, child_base i32
// (-1 = leaf), meta u32 (action | n_child << 8 | inflight << 16).
//
// Exactness: Q and u are evaluated in float64 with numpy's promotion rules (float32 prior * weak Python
// scalar c_puct rounds to float32 first; mcts.py:69), so with one leaf per game per
// wave (K = 1) visit counts equal the reference's bit for bit.  K > 1 adds virtual loss (documented deviation).
#include <math.h>

#include "qz_common.cuh"
#include "qz_philox.cuh"
#include "qz_warp.cuh"

__device__ __forceinline__ int qz_meta_action(uint32_t m) { return (int)(m & 0xFFu); }
__device__ __forceinline__ int qz_meta_nchild(uint32_t m) { return (int)((m >> 8) & 0xFFu); }
__device__ __forceinline__ int qz_meta_inflight(uint32_t m) { return (int)(m >> 16); }

static int qz_tree_check(const qz_tree *t, const char *fn) {
    if (t == nullptr) return qz_fail(QZ_E_NULL, "%s: tree is NULL", fn);
    if (t->n_games < 0 || t->node_cap < 1 || t->max_depth < 2 || t->leaves_per_game < 1)
        return qz_fail(QZ_E_RANGE, "%s: bad tree dimensions", fn);
    if (!t->prior || !t->visits || !t->q || !t->child_base || !t->node_meta || !t->root || !t->n_nodes ||
        !t->root_state || !t->leaf_node || !t->leaf_state || !t->path || !t->path_len || !t->leaf_flags)
        return qz_fail(QZ_E_NULL, "%s: a tree array is NULL", fn);
    return 0;
}

// ------------------------------------------------------------------------------------------ init
// MCTS.__init__ (mcts.py:97) / update_with_move(-1) (mcts.py:150-151): a fresh root with prior 1.0.
__global__ void qz_mcts_init_kernel(qz_tree t, const qz_state *__restrict__ root_states, const uint8_t *__restrict__ sel) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.n_games) return;
    if (sel != nullptr && !sel[g]) return;
    const int64_t o = g * t.node_cap;
    t.prior[o] = 1.0f; t.visits[o] = 0; t.q[o] = 0.0; t.child_base[o] = -1; t.node_meta[o] = 0;
    t.root[g] = 0;
    t.n_nodes[g] = 1;
    if (root_states != nullptr) qz_store_state(t.root_state + g, qz_load_state(root_states + g));
}

extern "C" int qz_mcts_init(const qz_tree *tree, const qz_state *root_states, const uint8_t *select, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_init");
    if (rc) return rc;
    if (tree->n_games == 0) return 0;
    qz_mcts_init_kernel<<<qz_blocks_for(tree->n_games, 256), 256, 0, (cudaStream_t)stream>>>(*tree, root_states, select);
    return qz_check_launch("qz_mcts_init");
}

// ------------------------------------------------------------------------------------------ select
// One warp per game; k_leaves sequential descents (virtual loss between them).
#define QZ_SEL_MAX_CHILDREN 144        // a node has at most 140 children (12 pawn ids + 128 walls)
template <bool UNIFORM_PRIOR>
__global__ void __launch_bounds__(128, 7) qz_mcts_select_kernel(qz_tree t, double c_puct, int k_leaves) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const int64_t o = g * t.node_cap;
    float *__restrict__ prior = t.prior + o;
    int32_t *__restrict__ visits = t.visits + o;
    double *__restrict__ q = t.q + o;
    int32_t *__restrict__ child_base = t.child_base + o;
    uint32_t *__restrict__ meta = t.node_meta + o;
    const QzState root_state = qz_load_state(t.root_state + g);
    const int root = t.root[g];
    const int K = t.leaves_per_game;
    // The k_leaves descents of a game all start at the root, whose children's statistics only change by the
    // in-flight marks this warp sets itself: they are read once into shared memory, so the first level of every
    // descent costs no global round trip (the select pass is a chain of dependent loads, ~4 per level).
    __shared__ int32_t s_n[4][QZ_SEL_MAX_CHILDREN];
    __shared__ uint32_t s_meta[4][QZ_SEL_MAX_CHILDREN];
    __shared__ double s_q[4][QZ_SEL_MAX_CHILDREN];
    __shared__ float s_prior[UNIFORM_PRIOR ? 1 : 4][UNIFORM_PRIOR ? 1 : QZ_SEL_MAX_CHILDREN];
    const int wib = threadIdx.x >> 5;
    const int rbase = child_base[root];
    uint32_t rmeta = 0;
    int rnc = 0, rvis = 0, root_marks = 0;
    if (rbase >= 0) {
        rmeta = meta[root];
        rnc = qz_meta_nchild(rmeta);
        rvis = visits[root];
        for (int j = lane; j < rnc; j += 32) {
            s_n[wib][j] = visits[rbase + j];
            s_meta[wib][j] = meta[rbase + j];
            s_q[wib][j] = q[rbase + j];
            if (!UNIFORM_PRIOR) s_prior[wib][j] = prior[rbase + j];
        }
    }
    __syncwarp();
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        if (k >= k_leaves) {
            // an unused leaf slot gets a finished position: the legality sweep and the rollout kernels that run over
            // all n*K slots then drop it at once (the one-leaf first wave of a search would otherwise evaluate the
            // stale positions of 63 slots per game -- a whole wave's work)
            if (lane == 0) {
                t.leaf_flags[L] = QZ_LEAF_INACTIVE; t.path_len[L] = 0; t.leaf_node[L] = -1;
                QzState idle;
                idle.H = 0; idle.V = 0; idle.meta = qz_pack_meta(4, 76, 0, 0, 1, QZ_FLAG_DONE, 0);
                qz_store_state(t.leaf_state + L, idle);
            }
            continue;
        }
        int32_t *__restrict__ path = t.path + L * t.max_depth;
        QzState s = root_state;
        int node = root, depth = 0;
        unsigned flags = 0;
        if (lane == 0) { path[0] = node; meta[root] += (1u << 16); }   // every node on the path carries one in-flight mark
        root_marks++;
        __syncwarp();
        for (;;) {
            const bool at_root = depth == 0;
            const int base = at_root ? rbase : child_base[node];
            if (base < 0) break;                                        // is_leaf (mcts.py:76)
            const uint32_t pm = at_root ? rmeta + ((uint32_t)root_marks << 16) : meta[node];
            const int nc = qz_meta_nchild(pm);
            const int np_eff = (at_root ? rvis : visits[node]) + qz_meta_inflight(pm) - 1;  // minus this descent's own mark
            const double sq = sqrt((double)np_eff);                     // np.sqrt(parent._n_visits)
            const double uni = UNIFORM_PRIOR ? c_puct * (1.0 / (double)nc) : 0.0;   // pure_mcts.py:15
            double best = -INFINITY;
            int bj = 0x7FFFFFFF;
            for (int j = lane; j < nc; j += 32) {
                const int c = base + j;
                const uint32_t cm = at_root ? s_meta[wib][j] : meta[c];
                const int n = at_root ? s_n[wib][j] : visits[c], infl = qz_meta_inflight(cm);
                double qv = at_root ? s_q[wib][j] : q[c];
                if (infl > 0) {                                         // virtual loss (K > 1 only)
                    // a zero numerator (e.g. one win, one mark) would send the FP64 division down its ~100-instruction
                    // special-case path, which was a quarter of this kernel's instructions; 0 / x is +0 either way
                    const double num = qv * (double)n - (double)infl;
                    qv = num == 0.0 ? 0.0 : num / (double)(n + infl);
                }
                double cp;
                if (UNIFORM_PRIOR) cp = uni;
                else cp = (double)((float)c_puct * (at_root ? s_prior[wib][j] : prior[c]));   // float32 product first (numpy weak scalar)
                const double u = cp * sq / (double)(1 + n + infl);      // mcts.py:69
                const double v = qv + u;
                if (v > best) { best = v; bj = j; }                     // ascending j per lane: first max kept
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(QZ_FULL_MASK, best, off);
                const int oj = __shfl_xor_sync(QZ_FULL_MASK, bj, off);
                if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
            }
            if (bj == 0x7FFFFFFF) bj = 0;                                // all-NaN guard; never in practice
            node = base + bj;
            const uint32_t cm = at_root ? s_meta[wib][bj] : meta[node];
            s = qz_apply(s, qz_meta_action(cm));                         // game.step(action), mcts.py:113
            depth++;
            __syncwarp();                                                // everyone has read s_meta[bj]
            if (lane == 0) {
                path[depth] = node;
                meta[node] = cm + (1u << 16);
                if (at_root) s_meta[wib][bj] = cm + (1u << 16);
            }
            __syncwarp();
            if (depth >= t.max_depth - 1) { flags |= QZ_LEAF_DEPTH_OVERFLOW; break; }
        }
        if (qz_done(s.meta)) flags |= QZ_LEAF_TERMINAL;
        if (lane == 0) {
            t.leaf_node[L] = node;
            t.path_len[L] = depth + 1;
            t.leaf_flags[L] = (uint8_t)flags;
            qz_store_state(t.leaf_state + L, s);
        }
        __syncwarp();
    }
}

extern "C" int qz_mcts_select(const qz_tree *tree, double c_puct, int uniform_prior, int k_leaves, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_select");
    if (rc) return rc;
    QZ_REQUIRE(k_leaves >= 0 && k_leaves <= tree->leaves_per_game);
    if (tree->n_games == 0) return 0;
    const unsigned blocks = qz_blocks_for(tree->n_games, 4);
    if (uniform_prior) qz_mcts_select_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(*tree, c_puct, k_leaves);
    else qz_mcts_select_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(*tree, c_puct, k_leaves);
    return qz_check_launch("qz_mcts_select");
}

// ------------------------------------------------------------------------------------------ expand + backup
// TreeNode.expand (mcts.py:27-35) with the (action, prob) pairs of policy_value_fn (policy_value_net.py:162:
// probs[legal], NOT renormalised), then update_recursive(-leaf_value) (mcts.py:44-62,127).
// One warp per game, its leaves in order, so a game's tree is only ever touched by one warp: no atomics.
struct QzExpandArgs {
    const uint64_t *mask3;      // [n*K,3] legal masks of the leaves (qz_env_legal_mask on leaf_state)
    const float *priors;        // [n*K,140] or NULL (uniform 1/len, pure_mcts.py:13-16)
    const float *value_f32;     // [n*K] leaf value for the side to move (net), or NULL
    const double *value_f64;    // [n*K] same in float64 (stubs), or NULL
    const int8_t *value_i8;     // [n*K] same as +1/0/-1 (rollouts), or NULL
    int fix_terminal_sign;      // 0 = reference behaviour (a winning move backs up as a loss, mcts.py:125)
    int32_t *overflow_count;    // nullable: incremented when an expansion did not fit the arena
};

__global__ void __launch_bounds__(128) qz_mcts_expand_backup_kernel(qz_tree t, QzExpandArgs a) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const int64_t o = g * t.node_cap;
    float *__restrict__ prior = t.prior + o;
    int32_t *__restrict__ visits = t.visits + o;
    double *__restrict__ q = t.q + o;
    int32_t *__restrict__ child_base = t.child_base + o;
    uint32_t *__restrict__ meta = t.node_meta + o;
    const int K = t.leaves_per_game;
    int n_nodes = t.n_nodes[g];
    const int root = t.root[g];
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        unsigned flags = t.leaf_flags[L];
        if (flags & QZ_LEAF_INACTIVE) continue;
        const int node = t.leaf_node[L];
        const int node_cb = child_base[node];
        __syncwarp();
        double v;
        bool pending = false;
        if (flags & QZ_LEAF_TERMINAL) {
            v = a.fix_terminal_sign ? -1.0 : 1.0;                       // mcts.py:125 (mover not rotated => +1)
        } else {
            v = a.value_f64 ? a.value_f64[L] : (a.value_f32 ? (double)a.value_f32[L] : (double)a.value_i8[L]);
            // a rollout still running in the deferred pass: expand now, back up later (qz_mcts_backup_pending);
            // the path keeps its in-flight marks meanwhile
            pending = a.value_i8 && a.value_i8[L] == (int8_t)QZ_ROLLOUT_PENDING;
            if (node_cb < 0 && !(flags & QZ_LEAF_DEPTH_OVERFLOW)) {
                uint32_t pawn; uint64_t hl, vl;
                const uint64_t mk[3] = {a.mask3[3 * L], a.mask3[3 * L + 1], a.mask3[3 * L + 2]};
                qz_unpack_mask(mk, pawn, hl, vl);
                const int cnt = qz_popc32(pawn) + qz_popc64(hl) + qz_popc64(vl);
                if (cnt > 0) {
                    if (n_nodes + cnt <= t.node_cap) {
                        const int base = n_nodes;
                        n_nodes += cnt;
                        const float up = 1.0f / (float)cnt;
#pragma unroll
                        for (int r = 0; r < 5; r++) {
                            const int act = lane + 32 * r;
                            if (act >= QZ_N_ACTIONS) break;
                            const bool legal = act < 12 ? (pawn >> act) & 1u
                                                        : (act < 76 ? (hl >> (act - 12)) & 1ull : (vl >> (act - 76)) & 1ull);
                            if (!legal) continue;
                            const int c = base + qz_action_rank(pawn, hl, vl, act);
                            prior[c] = a.priors ? a.priors[L * QZ_N_ACTIONS + act] : up;
                            visits[c] = 0; q[c] = 0.0; child_base[c] = -1; meta[c] = (uint32_t)act;
                        }
                        if (lane == 0) {
                            child_base[node] = base;
                            meta[node] = (meta[node] & 0xFFFF00FFu) | ((uint32_t)cnt << 8);
                        }
                    } else {
                        flags |= QZ_LEAF_ARENA_OVERFLOW;
                        if (lane == 0 && a.overflow_count) atomicAdd(a.overflow_count, 1);
                    }
                }
            } else if (node_cb >= 0) {
                flags |= QZ_LEAF_DUPLICATE;
            }
        }
        __syncwarp();
        if (pending) {
            if (lane == 0) t.leaf_flags[L] = (uint8_t)(flags | QZ_LEAF_PENDING);
            __syncwarp();
            continue;
        }
        {
            // update_recursive(-leaf_value), mcts.py:56-62,127: the nodes of one path are distinct, so lane d updates
            // the node at depth d (one memory latency per leaf instead of one per level); leaves are still backed up
            // one after the other, so every node sees its updates in playout order as in the reference
            const int32_t *path = t.path + L * t.max_depth;
            const int len = t.path_len[L];
            for (int d0 = 0; d0 < len; d0 += 32) {
                const int d = d0 + lane;
                if (d < len) {
                    const int nd = path[d];
                    const double x = ((len - 1 - d) & 1) ? v : -v;       // the sign flips per level (mcts.py:61)
                    const int nv = visits[nd] + 1;
                    visits[nd] = nv;
                    const double qo = q[nd];
                    const double num = 1.0 * (x - qo);
                    q[nd] = num == 0.0 ? qo + 0.0 : qo + num / (double)nv;   // mcts.py:53 (0 / n spared: FP64 special-case path)
                    meta[nd] -= (1u << 16);                              // clear this descent's virtual loss
                }
            }
            if (lane == 0) t.leaf_flags[L] = (uint8_t)flags;
        }
        __syncwarp();
    }
    if (lane == 0) t.n_nodes[g] = n_nodes;
}

extern "C" int qz_mcts_expand_backup(const qz_tree *tree, const uint64_t *mask3, const float *priors,
                                     const float *value_f32, const double *value_f64, const int8_t *value_i8,
                                     int fix_terminal_sign, int32_t *overflow_count, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_expand_backup");
    if (rc) return rc;
    QZ_REQUIRE_PTR(mask3);
    if (!value_f32 && !value_f64 && !value_i8) return qz_fail(QZ_E_NULL, "qz_mcts_expand_backup: no value array");
    if (tree->n_games == 0) return 0;
    QzExpandArgs a;
    a.mask3 = mask3; a.priors = priors; a.value_f32 = value_f32; a.value_f64 = value_f64; a.value_i8 = value_i8;
    a.fix_terminal_sign = fix_terminal_sign; a.overflow_count = overflow_count;
    qz_mcts_expand_backup_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*tree, a);
    return qz_check_launch("qz_mcts_expand_backup");
}

// Back up the leaves whose rollout was deferred (QZ_LEAF_PENDING) once their values are known.  One thread per
// game, its pending leaves in order; same arithmetic as qz_mcts_expand_backup_kernel.
__global__ void qz_mcts_backup_pending_kernel(qz_tree t, const int8_t *__restrict__ value_i8, int fix_terminal_sign) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.n_games) return;
    const int64_t o = g * t.node_cap;
    int32_t *__restrict__ visits = t.visits + o;
    double *__restrict__ q = t.q + o;
    uint32_t *__restrict__ meta = t.node_meta + o;
    const int K = t.leaves_per_game;
    const int root = t.root[g];
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        const unsigned flags = t.leaf_flags[L];
        if (!(flags & QZ_LEAF_PENDING)) continue;
        const double v = (flags & QZ_LEAF_TERMINAL) ? (fix_terminal_sign ? -1.0 : 1.0) : (double)value_i8[L];
        const int32_t *path = t.path + L * t.max_depth;
        const int len = t.path_len[L];
        double x = -v;
        for (int d = len - 1; d >= 0; d--) {
            const int nd = path[d];
            const int nv = visits[nd] + 1;
            visits[nd] = nv;
            const double qo = q[nd];
            q[nd] = qo + 1.0 * (x - qo) / (double)nv;
            if (d > 0) meta[nd] -= (1u << 16);
            x = -x;
        }
        meta[root] -= (1u << 16);
        t.leaf_flags[L] = (uint8_t)(flags & ~QZ_LEAF_PENDING);
    }
}

extern "C" int qz_mcts_backup_pending(const qz_tree *tree, const int8_t *value_i8, int fix_terminal_sign, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_backup_pending");
    if (rc) return rc;
    QZ_REQUIRE_PTR(value_i8);
    if (tree->n_games == 0) return 0;
    qz_mcts_backup_pending_kernel<<<qz_blocks_for(tree->n_games, 128), 128, 0, (cudaStream_t)stream>>>(*tree, value_i8,
                                                                                                       fix_terminal_sign);
    return qz_check_launch("qz_mcts_backup_pending");
}

// ------------------------------------------------------------------------------------------ root statistics
// get_move_probs (mcts.py:141-144): visits of the root's children, scattered by action id into 140-vectors,
// and softmax(1/temp * log(visits + 1e-10)) (mcts.py:6-9,143) in float64.  One warp per game.
__global__ void __launch_bounds__(128) qz_mcts_root_stats_kernel(qz_tree t, double temp, int32_t *__restrict__ visits_out,
                                                                 double *__restrict__ q_out, double *__restrict__ probs_out,
                                                                 int32_t *__restrict__ root_n_out) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const int64_t o = g * t.node_cap;
    const int root = t.root[g];
    const int base = t.child_base[o + root];
    const int nc = base < 0 ? 0 : qz_meta_nchild(t.node_meta[o + root]);
    for (int act = lane; act < QZ_N_ACTIONS; act += 32) {
        if (visits_out) visits_out[g * QZ_N_ACTIONS + act] = 0;
        if (q_out) q_out[g * QZ_N_ACTIONS + act] = 0.0;
        if (probs_out) probs_out[g * QZ_N_ACTIONS + act] = 0.0;
    }
    if (lane == 0 && root_n_out) root_n_out[g] = t.visits[o + root];
    __syncwarp();
    double mx = -INFINITY;
    for (int j = lane; j < nc; j += 32) {
        const double x = 1.0 / temp * log((double)t.visits[o + base + j] + 1e-10);
        mx = fmax(mx, x);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(QZ_FULL_MASK, mx, off));
    double sum = 0.0;
    for (int j = lane; j < nc; j += 32) sum += exp(1.0 / temp * log((double)t.visits[o + base + j] + 1e-10) - mx);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(QZ_FULL_MASK, sum, off);
    for (int j = lane; j < nc; j += 32) {
        const int c = base + j;
        const int act = qz_meta_action(t.node_meta[o + c]);
        const int n = t.visits[o + c];
        if (visits_out) visits_out[g * QZ_N_ACTIONS + act] = n;
        if (q_out) q_out[g * QZ_N_ACTIONS + act] = t.q[o + c];
        if (probs_out) probs_out[g * QZ_N_ACTIONS + act] = exp(1.0 / temp * log((double)n + 1e-10) - mx) / sum;
    }
}

extern "C" int qz_mcts_root_stats(const qz_tree *tree, double temp, int32_t *visits_out, double *q_out, double *probs_out,
                                  int32_t *root_n_out, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_root_stats");
    if (rc) return rc;
    QZ_REQUIRE(temp > 0.0);
    if (tree->n_games == 0) return 0;
    qz_mcts_root_stats_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(
        *tree, temp, visits_out, q_out, probs_out, root_n_out);
    return qz_check_launch("qz_mcts_root_stats");
}

// ------------------------------------------------------------------------------------------ choose a move
// MCTSPlayer.choose_action (mcts.py:172-196) / pure_mcts.MCTS.get_move (:115).
//   mode 0: first-max of visits (pure_mcts.py:115; also the deterministic choice used by parity tests)
//   mode 1: sample from probs                     (mcts.py:185, np.random.choice(acts, p=probs))
//   mode 2: sample from 0.75*probs + 0.25*Dir(0.3) (mcts.py:181, self-play)
// Randomness: Philox keyed by (seed, game_id[g], ply) -- independent of sharding; the reference uses the
// global numpy RNG, so parity for modes 1/2 is distributional only.
__device__ __forceinline__ double qz_u01(uint32_t a, uint32_t b) {   // (0,1) from 53 random bits
    const uint64_t x = ((uint64_t)a << 21) ^ (uint64_t)b;
    return ((double)(x & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// Gamma(alpha<1) via Marsaglia-Tsang on alpha+1 and the U^(1/alpha) boost; counter-based draws.
__device__ double qz_gamma_small(double alpha, uint64_t seed, uint64_t rid, uint32_t c2) {
    const double d = alpha + 1.0 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t it = 0; it < 64; it++) {
        const QzPhilox4 r = qz_philox(seed, rid, c2, 0x40000000u + it);
        const QzPhilox4 r2 = qz_philox(seed, rid, c2, 0x50000000u + it);
        const double u1 = qz_u01(r.x, r.y), u2 = qz_u01(r.z, r.w);
        const double nrm = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);     // Box-Muller
        const double vv = 1.0 + c * nrm;
        if (vv <= 0.0) continue;
        const double v3 = vv * vv * vv;
        const double u = qz_u01(r2.x, r2.y);
        if (log(u) < 0.5 * nrm * nrm + d - d * v3 + d * log(v3)) {
            const double boost = pow(qz_u01(r2.z, r2.w), 1.0 / alpha);
            return d * v3 * boost;
        }
    }
    return alpha;
}

__global__ void __launch_bounds__(128) qz_mcts_choose_kernel(qz_tree t, int mode, double temp, double noise_eps,
                                                             double dir_alpha, uint64_t seed,
                                                             const int64_t *__restrict__ game_id,
                                                             int32_t *__restrict__ moves_out) {
    __shared__ double pbuf[4][QZ_N_ACTIONS];
    const int w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const int64_t o = g * t.node_cap;
    const int root = t.root[g];
    const int base = t.child_base[o + root];
    const int nc = base < 0 ? 0 : qz_meta_nchild(t.node_meta[o + root]);
    if (nc == 0) { if (lane == 0) moves_out[g] = -1; return; }          // mcts.py:195-196 ("board is full")
    if (mode == 0) {
        int bv = -1, bj = 0x7FFFFFFF;
        for (int j = lane; j < nc; j += 32) {
            const int n = t.visits[o + base + j];
            if (n > bv) { bv = n; bj = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int ov = __shfl_xor_sync(QZ_FULL_MASK, bv, off), oj = __shfl_xor_sync(QZ_FULL_MASK, bj, off);
            if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        if (lane == 0) moves_out[g] = qz_meta_action(t.node_meta[o + base + bj]);
        return;
    }
    const uint64_t rid = game_id ? (uint64_t)game_id[g] : (uint64_t)g;
    const uint32_t ply = qz_ply(t.root_state[g].meta);
    // probabilities exactly as root_stats
    double mx = -INFINITY;
    for (int j = lane; j < nc; j += 32) mx = fmax(mx, 1.0 / temp * log((double)t.visits[o + base + j] + 1e-10));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(QZ_FULL_MASK, mx, off));
    double sum = 0.0, gsum = 0.0;
    for (int j = lane; j < nc; j += 32) {
        const double e = exp(1.0 / temp * log((double)t.visits[o + base + j] + 1e-10) - mx);
        pbuf[w][j] = e;
        sum += e;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(QZ_FULL_MASK, sum, off);
    double gam[5];
    if (mode == 2) {
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int j = lane + 32 * r;
            gam[r] = j < nc ? qz_gamma_small(dir_alpha, seed ^ 0xD1B54A32D192ED03ull, rid, (ply << 8) | (uint32_t)j) : 0.0;
            gsum += gam[r];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gsum += __shfl_xor_sync(QZ_FULL_MASK, gsum, off);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int j = lane + 32 * r;
        if (j < nc) {
            double p = pbuf[w][j] / sum;
            if (mode == 2) p = (1.0 - noise_eps) * p + noise_eps * (gam[r] / gsum);
            pbuf[w][j] = p;
        }
    }
    __syncwarp();
    if (lane == 0) {
        const QzPhilox4 r = qz_philox(seed ^ 0x2545F4914F6CDD1Dull, rid, ply, 0x60000000u);
        const double u = qz_u01(r.x, r.y);
        double acc = 0.0;
        int pick = nc - 1;
        for (int j = 0; j < nc; j++) {
            acc += pbuf[w][j];
            if (u < acc) { pick = j; break; }
        }
        moves_out[g] = qz_meta_action(t.node_meta[o + base + pick]);
    }
}

extern "C" int qz_mcts_choose(const qz_tree *tree, int mode, double temp, double noise_eps, double dir_alpha,
                              uint64_t seed, const int64_t *game_id, int32_t *moves_out, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_choose");
    if (rc) return rc;
    QZ_REQUIRE_PTR(moves_out);
    QZ_REQUIRE(mode >= 0 && mode <= 2 && temp > 0.0 && dir_alpha > 0.0 && noise_eps >= 0.0 && noise_eps <= 1.0);
    if (tree->n_games == 0) return 0;
    qz_mcts_choose_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(
        *tree, mode, temp, noise_eps, dir_alpha, seed, game_id, moves_out);
    return qz_check_launch("qz_mcts_choose");
}

// ------------------------------------------------------------------------------------------ re-root
// MCTS.update_with_move (mcts.py:146-151): the chosen child becomes the root and keeps its statistics; an
// unknown move (-1) gives a fresh root.  The kept subtree is COMPACTED from the current arena (`src`) into the
// other one (`dst`), breadth first, so children stay contiguous and the arena never fragments.  While a dst
// node waits to be processed its child_base holds the index of its src twin.  One warp per game.  The
// root state advances by the move (Quoridor.step, quoridor.py:159-186).
__global__ void __launch_bounds__(128) qz_mcts_reroot_kernel(qz_tree src, qz_tree dst, const int32_t *__restrict__ moves,
                                                             int apply_move) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= src.n_games) return;
    const int lane = threadIdx.x & 31;
    const int64_t so = g * src.node_cap, dofs = g * dst.node_cap;
    const int move = moves[g];
    QzState s = qz_load_state(src.root_state + g);
    if (apply_move && move >= 0) s = qz_apply(s, move);
    if (lane == 0) qz_store_state(dst.root_state + g, s);
    // find the child of the root that carries `move`
    const int root = src.root[g];
    const int rbase = src.child_base[so + root];
    const int rnc = rbase < 0 ? 0 : qz_meta_nchild(src.node_meta[so + root]);
    int found = -1;
    for (int j = lane; j < rnc; j += 32)
        if (move >= 0 && qz_meta_action(src.node_meta[so + rbase + j]) == move) found = rbase + j;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) found = max(found, __shfl_xor_sync(QZ_FULL_MASK, found, off));
    if (lane == 0) dst.root[g] = 0;
    if (found < 0) {                                                     // fresh tree (mcts.py:150-151)
        if (lane == 0) {
            dst.prior[dofs] = 1.0f; dst.visits[dofs] = 0; dst.q[dofs] = 0.0; dst.child_base[dofs] = -1;
            dst.node_meta[dofs] = 0; dst.n_nodes[g] = 1;
        }
        return;
    }
    if (lane == 0) {
        dst.prior[dofs] = src.prior[so + found];
        dst.visits[dofs] = src.visits[so + found];
        dst.q[dofs] = src.q[so + found];
        dst.node_meta[dofs] = src.node_meta[so + found] & 0x0000FFFFu;   // drop stale in-flight marks
        dst.child_base[dofs] = src.child_base[so + found] >= 0 ? found : -1;   // src twin, pending
    }
    __syncwarp();
    int tail = 1;
    for (int head = 0; head < tail;) {
        const int scan = min(32, tail - head);                           // nodes appended below are scanned later
        const int i = head + lane;
        int twin = -1;
        if (lane < scan) twin = dst.child_base[dofs + i];                // >= 0: src index whose children to copy
        unsigned pending = __ballot_sync(QZ_FULL_MASK, twin >= 0);
        while (pending) {
            const int l = __ffs(pending) - 1;
            pending &= pending - 1;
            const int tw = __shfl_sync(QZ_FULL_MASK, twin, l);
            const int sb = src.child_base[so + tw];
            const int nc = qz_meta_nchild(src.node_meta[so + tw]);
            const int db = tail;
            for (int j = lane; j < nc; j += 32) {
                const int sc = sb + j, dc = db + j;
                dst.prior[dofs + dc] = src.prior[so + sc];
                dst.visits[dofs + dc] = src.visits[so + sc];
                dst.q[dofs + dc] = src.q[so + sc];
                dst.node_meta[dofs + dc] = src.node_meta[so + sc] & 0x0000FFFFu;
                dst.child_base[dofs + dc] = src.child_base[so + sc] >= 0 ? sc : -1;
            }
            if (lane == 0) dst.child_base[dofs + head + l] = db;
            tail += nc;
        }
        head += scan;
        __syncwarp();
    }
    if (lane == 0) dst.n_nodes[g] = tail;
}

extern "C" int qz_mcts_reroot(const qz_tree *src, const qz_tree *dst, const int32_t *moves, int apply_move, void *stream) {
    int rc = qz_tree_check(src, "qz_mcts_reroot");
    if (rc) return rc;
    rc = qz_tree_check(dst, "qz_mcts_reroot");
    if (rc) return rc;
    QZ_REQUIRE_PTR(moves);
    QZ_REQUIRE(src->n_games == dst->n_games && dst->node_cap >= src->node_cap);
    QZ_REQUIRE(src->prior != dst->prior && src->child_base != dst->child_base);
    if (src->n_games == 0) return 0;
    qz_mcts_reroot_kernel<<<qz_blocks_for(src->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*src, *dst, moves, apply_move);
    return qz_check_launch("qz_mcts_reroot");
}

// ------------------------------------------------------------------------------------------ deterministic stubs
// The parity stubs of tests/golden/stubs.py (S1 uniform, S2 hash, S3 hash/8) evaluated on the device, so
// that stub-driven MCTS runs entirely on the GPU.  One warp per leaf.
__device__ __forceinline__ uint64_t qz_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(128) qz_stub_eval_kernel(const qz_state *__restrict__ states,
                                                           const uint64_t *__restrict__ mask3, int kind,
                                                           float *__restrict__ priors, double *__restrict__ values,
                                                           int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const QzState s = qz_load_state(states + i);
    const uint64_t m0 = mask3[3 * i], m1 = mask3[3 * i + 1], m2 = mask3[3 * i + 2];
    const int cnt = qz_popc64(m0) + qz_popc64(m1) + qz_popc64(m2);
    const uint64_t meta = s.meta & 0xFFFFFFFFFFull;                      // p1,p2,w1,w2,cur (bytes 0..4)
    const uint64_t key = qz_splitmix64(s.H ^ qz_splitmix64(s.V ^ qz_splitmix64(meta)));
    for (int a = lane; a < QZ_N_ACTIONS; a += 32) {
        const bool legal = ((a < 64 ? m0 : (a < 128 ? m1 : m2)) >> (a & 63)) & 1ull;
        float p = 0.0f;
        if (legal) {
            if (kind == 1) p = 1.0f / (float)(cnt > 0 ? cnt : 1);
            else p = (float)((qz_splitmix64(key + (uint64_t)a * 0x9E3779B97F4A7C15ull) >> 40) + 1) * 0x1p-30f;
        }
        priors[i * QZ_N_ACTIONS + a] = p;
    }
    if (lane == 0) {
        double v = 0.0;
        if (kind != 1) {
            v = (double)(qz_splitmix64(key ^ 0xABCDEFull) >> 40) / 8388608.0 - 1.0;
            if (kind == 3) v /= 8.0;
        }
        values[i] = v;
    }
}

extern "C" int qz_stub_eval(const qz_state *states, const uint64_t *mask3, int kind, float *priors, double *values,
                            int64_t n, void *stream) {
    QZ_REQUIRE(n >= 0 && kind >= 1 && kind <= 3);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(mask3);
    QZ_REQUIRE_PTR(priors);
    QZ_REQUIRE_PTR(values);
    qz_stub_eval_kernel<<<qz_blocks_for(n, 4), 128, 0, (cudaStream_t)stream>>>(states, mask3, kind, priors, values, n);
    return qz_check_launch("qz_stub_eval");
}
