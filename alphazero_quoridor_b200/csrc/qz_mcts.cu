// qz_mcts.cu -- K5-K8: batched PUCT MCTS over many concurrent games, trees in flat device arrays.
//
// Reference semantics restated (file:line into the reference):
//   TreeNode            mcts.py:12-80  (dup pure_mcts.py:19-56)   -> one SoA slot per VISITED child + one block header per
//                                                                    expanded node
//   MCTS._playout       mcts.py:103-127 / pure_mcts.py:66-83      -> select kernel + expand_backup kernel
//   get_move_probs      mcts.py:129-144                           -> root_stats kernel
//   update_with_move    mcts.py:146-151                           -> reroot kernel (subtree compaction)
//   MCTSPlayer.choose_action  mcts.py:172-196                     -> choose kernel
//
// Layout (lazy children).  Game g owns the slots [g*node_cap, (g+1)*node_cap) of five arrays: prior f32, visits i32,
// Q f64, child_base i32, meta u32.  TreeNode.expand (mcts.py:27-35) creates up to ~131 children of which a search of a
// thousand playouts ever visits a handful, so an expansion does NOT write them.  It writes a BLOCK: a 3-slot header
// holding the node's 140-bit legal mask, the number of children in existence (m), the block's capacity and the number
// of legal actions, followed by room for `cap` child slots.  A child slot comes into existence the first time the
// descent picks that child.  This is exact: a child never visited has Q = 0 and n = 0, so its PUCT value is
// c_puct * P * sqrt(N) (mcts.py:69) and the reference's first-max over all children (mcts.py:42) can only ever pick,
// among the unvisited ones, the one with the largest c_puct * P (lowest actions() rank on ties) -- the header keeps that
// candidate, and the descent compares it with the best existing child by (value, rank).  With uniform priors
// (pure MCTS) the candidate is simply the next action in actions() order.  A block that is full moves to a larger one
// (doubling); its old slots keep a forwarding index, so the paths other leaves of the wave already recorded stay valid.
//   child slot : prior, visits, q, child_base (-1 = is_leaf(), >= 0 = its block, <= -2 = moved to slot -2 - x),
//                meta = action | actions() rank << 8 | in-flight (virtual loss) count << 16
//   header b   : q[b..b+2] = legal mask bits; visits[b] = m, visits[b+1] = cap, visits[b+2] = legal count;
//                net priors only: child_base[b] = offset of the node's priors in the game's prior pool (one float per
//                legal action, in actions() rank order), child_base[b+1] = rank of the candidate (-1 = none), prior[b] =
//                its prior, {meta[b], meta[b+1], meta[b+2], child_base[b+2], bits of prior[b+1]} = the set of ranks that
//                already have a slot
// 24 B per visited child + ~7 slots per expansion: ~0.3 MB per game for 1000 playouts (3.4 MB with eager children).
//
// Exactness: Q and u are evaluated in float64 with numpy's promotion rules (float32 prior * weak Python
// scalar c_puct rounds to float32 first; mcts.py:69), so with one leaf per game per
// wave (K = 1) visit counts equal the reference's bit for bit.  K > 1 adds virtual loss (documented deviation).
#include <math.h>

#include "qz_common.cuh"
#include "qz_philox.cuh"
#include "qz_warp.cuh"

#define QZ_HDR 3                       // header slots of a block
#define QZ_CAP0 4                      // initial capacity of a non-root block
// child_base of a slot: >= 0 its block; -1 never evaluated (is_leaf()); QZ_CB_LAZY evaluated once -- the reference
// expanded it then (pure_mcts.py:77-79) -- but its block is only built when a playout comes back (lazy expansion);
// -2 - x (x < node_cap) moved to slot x
#define QZ_CB_LAZY ((int32_t)0x80000000)

__device__ __forceinline__ int qz_meta_action(uint32_t m) { return (int)(m & 0xFFu); }
__device__ __forceinline__ int qz_meta_rank(uint32_t m) { return (int)((m >> 8) & 0xFFu); }
__device__ __forceinline__ int qz_meta_inflight(uint32_t m) { return (int)(m >> 16); }

static int qz_tree_check(const qz_tree *t, const char *fn) {
    if (t == nullptr) return qz_fail(QZ_E_NULL, "%s: tree is NULL", fn);
    if (t->n_games < 0 || t->node_cap < 1 + QZ_HDR + QZ_N_ACTIONS || t->max_depth < 2 || t->leaves_per_game < 1 || t->pool_cap < 0)
        return qz_fail(QZ_E_RANGE, "%s: bad tree dimensions", fn);
    if (!t->prior || !t->visits || !t->q || !t->child_base || !t->node_meta || !t->root || !t->n_nodes ||
        !t->root_state || !t->leaf_node || !t->leaf_state || !t->path || !t->path_len || !t->leaf_flags)
        return qz_fail(QZ_E_NULL, "%s: a tree array is NULL", fn);
    if (t->pool_cap > 0 && (!t->prior_pool || !t->n_pool)) return qz_fail(QZ_E_NULL, "%s: prior pool is NULL", fn);
    return 0;
}

// per-game view of the arrays
struct QzGameTree {
    float *prior; int32_t *visits; double *q; int32_t *child_base; uint32_t *meta; float *pool;
};
__device__ __forceinline__ QzGameTree qz_game_tree(const qz_tree &t, int64_t g) {
    const int64_t o = g * t.node_cap;
    QzGameTree v;
    v.prior = t.prior + o; v.visits = t.visits + o; v.q = t.q + o; v.child_base = t.child_base + o; v.meta = t.node_meta + o;
    v.pool = t.prior_pool ? t.prior_pool + g * (int64_t)t.pool_cap : nullptr;
    return v;
}
// follow the forwarding indices a relocated block left behind
__device__ __forceinline__ int qz_resolve(const int32_t *child_base, int i) {
    int cb;
    while ((cb = child_base[i]) <= -2 && cb != QZ_CB_LAZY) i = -2 - cb;
    return i;
}
__device__ __forceinline__ void qz_header_mask(const QzGameTree &v, int b, uint32_t &pawn, uint64_t &hl, uint64_t &vl) {
    const uint64_t mk[3] = {(uint64_t)__double_as_longlong(v.q[b]), (uint64_t)__double_as_longlong(v.q[b + 1]),
                            (uint64_t)__double_as_longlong(v.q[b + 2])};
    qz_unpack_mask(mk, pawn, hl, vl);
}
__device__ __forceinline__ bool qz_is_legal(uint32_t pawn, uint64_t hl, uint64_t vl, int act) {
    return act < 12 ? (pawn >> act) & 1u : (act < 76 ? (hl >> (act - 12)) & 1ull : (vl >> (act - 76)) & 1ull);
}
// the action with actions() rank k (quoridor.py:157,420-430: pawn ids ascending, then H(ix), V(ix) interleaved); whole warp
__device__ __forceinline__ int qz_action_of_rank(uint32_t pawn, uint64_t hl, uint64_t vl, int k) {
    const int np = __popc(pawn);
    if (k < np) return qz_nth_bit64((uint64_t)pawn, k);
    k -= np;
    const int lane = threadIdx.x & 31;
    int found = -1;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int ix = lane + 32 * half;
        const uint64_t below = (1ull << ix) - 1ull;
        const int before = __popcll(hl & below) + __popcll(vl & below);
        const int hb = (int)((hl >> ix) & 1ull), vb = (int)((vl >> ix) & 1ull);
        if (k >= before && k < before + hb + vb) found = (hb && k == before) ? 12 + ix : 76 + ix;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) found = max(found, __shfl_xor_sync(QZ_FULL_MASK, found, off));
    return found;
}
// net priors: the unvisited child the reference's first-max would reach first = largest float32 product c_puct * P
// (mcts.py:69 promotes that product to float64), lowest rank on ties, among the children without a slot.  `row` holds the
// node's priors in actions() RANK order (`total` of them), `mat` the ranks that already have a slot.
__device__ __forceinline__ void qz_next_candidate(const float *__restrict__ row, int total, const uint32_t mat[5],
                                                  float c_puct_f, int &nrank, float &nprior) {
    const int lane = threadIdx.x & 31;
    float bcp = -INFINITY, bp = 0.0f;
    int brank = -1;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int rank = lane + 32 * r;
        if (rank >= total || ((mat[r] >> lane) & 1u)) continue;
        const float p = row[rank], cp = c_puct_f * p;
        if (brank < 0 || cp > bcp) { bcp = cp; brank = rank; bp = p; }   // ascending rank per lane: first max kept
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ocp = __shfl_xor_sync(QZ_FULL_MASK, bcp, off), op = __shfl_xor_sync(QZ_FULL_MASK, bp, off);
        const int orank = __shfl_xor_sync(QZ_FULL_MASK, brank, off);
        if (orank >= 0 && (brank < 0 || ocp > bcp || (ocp == bcp && orank < brank))) { bcp = ocp; brank = orank; bp = op; }
    }
    nrank = brank; nprior = bp;
}
// header words holding the "has a slot" set by rank: bit (r & 31) of word r >> 5
__device__ __forceinline__ void qz_mat_load(const QzGameTree &v, int b, uint32_t mat[5]) {
    mat[0] = v.meta[b]; mat[1] = v.meta[b + 1]; mat[2] = v.meta[b + 2]; mat[3] = (uint32_t)v.child_base[b + 2];
    mat[4] = __float_as_uint(v.prior[b + 1]);
}
__device__ __forceinline__ void qz_mat_store(const QzGameTree &v, int b, const uint32_t mat[5]) {
    v.meta[b] = mat[0]; v.meta[b + 1] = mat[1]; v.meta[b + 2] = mat[2]; v.child_base[b + 2] = (int32_t)mat[3];
    v.prior[b + 1] = __uint_as_float(mat[4]);
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
    if (t.n_pool) t.n_pool[g] = 0;
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
// FAST_ROOT (used when a wave collects more than one leaf per game): the k_leaves descents of a game re-evaluate all of the
// root's children (up to 131) every time, two float64 divisions each; the root's per-child terms 1/(1+n+inflight) and the
// virtual-loss value are therefore kept in shared memory, refreshed for the one child a descent picks, and a child's
// value becomes one fused multiply-add.  Multiplying by a stored reciprocal is not bit-identical to the reference's
// division (mcts.py:69), which only the one-leaf-per-wave search reproduces bit for bit anyway -- that search keeps the
// literal arithmetic (FAST_ROOT = false).
template <bool UNIFORM_PRIOR, bool FAST_ROOT>
__global__ void __launch_bounds__(128, 6) qz_mcts_select_kernel(qz_tree t, double c_puct, int k_leaves, int lazy_expand,
                                                                int32_t *__restrict__ overflow_count) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    float *__restrict__ prior = v.prior;
    int32_t *__restrict__ visits = v.visits;
    double *__restrict__ q = v.q;
    int32_t *__restrict__ child_base = v.child_base;
    uint32_t *__restrict__ meta = v.meta;
    const QzState root_state = qz_load_state(t.root_state + g);
    const int root = t.root[g];
    const int K = t.leaves_per_game;
    const float c_puct_f = (float)c_puct;
    int n_nodes = t.n_nodes[g];
    // The k_leaves descents of a game all start at the root, whose children's statistics only change by the
    // in-flight marks (and new slots) this warp makes itself: they are read once into shared memory, so the first
    // level of every descent costs no global round trip (the select pass is a chain of dependent loads).
    __shared__ int32_t s_n[4][QZ_SEL_MAX_CHILDREN];
    __shared__ uint32_t s_meta[4][QZ_SEL_MAX_CHILDREN];
    __shared__ double s_q[4][QZ_SEL_MAX_CHILDREN];
    __shared__ float s_prior[UNIFORM_PRIOR ? 1 : 4][UNIFORM_PRIOR ? 1 : QZ_SEL_MAX_CHILDREN];
    __shared__ double s_rcp[FAST_ROOT ? 4 : 1][FAST_ROOT ? QZ_SEL_MAX_CHILDREN : 1];    // 1 / (1 + n + inflight)
    __shared__ double s_qv[FAST_ROOT ? 4 : 1][FAST_ROOT ? QZ_SEL_MAX_CHILDREN : 1];     // Q with the virtual loss applied
    const int wib = threadIdx.x >> 5;
    int rb = child_base[root];
    uint32_t rmeta = 0;
    int rm = 0, rtotal = 0, rvis = 0, root_marks = 0;
    if (rb >= 0) {
        rmeta = meta[root];
        rm = visits[rb]; rtotal = visits[rb + 2];
        rvis = visits[root];
        for (int j = lane; j < rm; j += 32) {
            const int c = rb + QZ_HDR + j;
            s_n[wib][j] = visits[c];
            s_meta[wib][j] = meta[c];
            s_q[wib][j] = q[c];
            if (!UNIFORM_PRIOR) s_prior[wib][j] = prior[c];
            if (FAST_ROOT) {
                const int n = s_n[wib][j], infl = qz_meta_inflight(s_meta[wib][j]);
                const double num = s_q[wib][j] * (double)n - (double)infl;
                s_qv[wib][j] = infl > 0 ? (num == 0.0 ? 0.0 : num / (double)(n + infl)) : s_q[wib][j];
                s_rcp[wib][j] = 1.0 / (double)(1 + n + infl);
            }
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
            int b = at_root ? rb : child_base[node];
            if (b < 0) {                                                // is_leaf (mcts.py:76) -- or, with lazy expansion,
                // a node that WAS expanded (evaluated before) but whose legal mask has not been needed yet:
                // qz_mcts_extend builds the block from the mask and takes this descent one level further
                if (UNIFORM_PRIOR && lazy_expand && b == QZ_CB_LAZY && !qz_done(s.meta)) flags |= QZ_LEAF_NEEDS_MASK;
                break;
            }
            // below the root only the block index is known at this point: the header, the legal mask and the first 32
            // child slots are requested in ONE round of loads (slots past m hold garbage that is masked below) -- the
            // descent is a chain of dependent loads, and this takes two links per level out of it
            int hm = 0, hcap = 0, htot = 0;
            double hq0 = 0.0, hq1 = 0.0, hq2 = 0.0, pq = 0.0;
            uint32_t pcm = 0;
            int pn = 0;
            float pp = 0.0f;
            if (!at_root) {
                hm = visits[b]; hcap = visits[b + 1]; htot = visits[b + 2];
                hq0 = q[b]; hq1 = q[b + 1]; hq2 = q[b + 2];
                const int c0 = b + QZ_HDR + lane;
                if (c0 < t.node_cap) {
                    pcm = meta[c0]; pn = visits[c0]; pq = q[c0];
                    if (!UNIFORM_PRIOR) pp = prior[c0];
                }
            }
            const int m = at_root ? rm : hm, total = at_root ? rtotal : htot;
            const uint32_t pm = at_root ? rmeta + ((uint32_t)root_marks << 16) : meta[node];
            const int np_eff = (at_root ? rvis : visits[node]) + qz_meta_inflight(pm) - 1;  // minus this descent's own mark
            const double sq = sqrt((double)np_eff);                     // np.sqrt(parent._n_visits)
            const double uni = UNIFORM_PRIOR ? c_puct * (1.0 / (double)total) : 0.0;   // pure_mcts.py:15
            double best = -INFINITY;
            int brank = 0x7FFFFFFF, bj = -1;
            if (FAST_ROOT && at_root) {
                for (int j = lane; j < m; j += 32) {
                    const double cp = UNIFORM_PRIOR ? uni : (double)(c_puct_f * s_prior[wib][j]);
                    const double val = fma(cp * sq, s_rcp[wib][j], s_qv[wib][j]);
                    const int rank = qz_meta_rank(s_meta[wib][j]);
                    if (val > best || (val == best && rank < brank)) { best = val; brank = rank; bj = j; }
                }
            } else
            for (int j = lane; j < m; j += 32) {
                const int c = b + QZ_HDR + j;
                const bool pre = !at_root && j == lane;                 // first round below the root: loaded above
                const uint32_t cm = at_root ? s_meta[wib][j] : (pre ? pcm : meta[c]);
                const int n = at_root ? s_n[wib][j] : (pre ? pn : visits[c]), infl = qz_meta_inflight(cm);
                double qv = at_root ? s_q[wib][j] : (pre ? pq : q[c]);
                if (infl > 0) {                                         // virtual loss (K > 1 only)
                    // a zero numerator (e.g. one win, one mark) would send the FP64 division down its ~100-instruction
                    // special-case path, which was a quarter of this kernel's instructions; 0 / x is +0 either way
                    const double num = qv * (double)n - (double)infl;
                    qv = num == 0.0 ? 0.0 : num / (double)(n + infl);
                }
                double cp;
                if (UNIFORM_PRIOR) cp = uni;
                else cp = (double)(c_puct_f * (at_root ? s_prior[wib][j] : (pre ? pp : prior[c])));   // float32 product first (numpy weak scalar)
                const double u = cp * sq / (double)(1 + n + infl);      // mcts.py:69
                const double val = qv + u;
                const int rank = qz_meta_rank(cm);
                if (val > best || (val == best && rank < brank)) { best = val; brank = rank; bj = j; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(QZ_FULL_MASK, best, off);
                const int orank = __shfl_xor_sync(QZ_FULL_MASK, brank, off), oj = __shfl_xor_sync(QZ_FULL_MASK, bj, off);
                if (oj >= 0 && (bj < 0 || ob > best || (ob == best && orank < brank))) { best = ob; brank = orank; bj = oj; }
            }
            // the best child WITHOUT a slot: never visited, so Q = 0, n = 0 and its value is c_puct * P * sqrt(N)
            bool take_new = false;
            int uact = -1, urank = 0;
            float uprior = 0.0f;
            uint32_t pawn = 0; uint64_t hl = 0, vl = 0;
            if (m < total) {
                if (at_root) {
                    qz_header_mask(v, b, pawn, hl, vl);
                } else {
                    const uint64_t mk[3] = {(uint64_t)__double_as_longlong(hq0), (uint64_t)__double_as_longlong(hq1),
                                            (uint64_t)__double_as_longlong(hq2)};
                    qz_unpack_mask(mk, pawn, hl, vl);
                }
                double cpu;
                if (UNIFORM_PRIOR) {
                    cpu = uni; urank = m;                                // slots are made in actions() order
                } else {
                    uprior = prior[b]; urank = child_base[b + 1];
                    cpu = (double)(c_puct_f * uprior);
                }
                const double u0 = cpu * sq / (double)(1 + 0 + 0);
                take_new = bj < 0 || u0 > best || (u0 == best && urank < brank);
            }
            if (take_new) {
                uact = qz_action_of_rank(pawn, hl, vl, urank);
                if (UNIFORM_PRIOR) uprior = 1.0f / (float)total;
                int cap = at_root ? visits[b + 1] : hcap;
                bool room = true;
                if (m == cap) {
                    // the block is full: move it to one of twice the size; the old slots forward to the new ones
                    const int newcap = min(total, max(2 * cap, QZ_CAP0));
                    if (n_nodes + QZ_HDR + newcap <= t.node_cap) {
                        const int nb = n_nodes;
                        n_nodes += QZ_HDR + newcap;
                        for (int j = lane; j < QZ_HDR + m; j += 32) {
                            const int src = b + j, dst = nb + j;
                            prior[dst] = prior[src]; visits[dst] = visits[src]; q[dst] = q[src];
                            child_base[dst] = child_base[src]; meta[dst] = meta[src];
                        }
                        __syncwarp();
                        for (int j = lane; j < m; j += 32) child_base[b + QZ_HDR + j] = -2 - (nb + QZ_HDR + j);
                        if (lane == 0) { visits[nb + 1] = newcap; child_base[node] = nb; }
                        __syncwarp();
                        b = nb;
                        if (at_root) rb = nb;
                    } else {
                        room = false;
                    }
                }
                if (!room) {
                    if (lane == 0 && overflow_count) atomicAdd(overflow_count, 1);
                    if (bj < 0) { flags |= QZ_LEAF_ARENA_OVERFLOW; break; }     // nothing to descend into: stop here
                    take_new = false;                                   // fall back to the best existing child
                } else {
                    const int c = b + QZ_HDR + m;
                    const uint32_t cm = (uint32_t)uact | ((uint32_t)urank << 8) | (1u << 16);
                    if (lane == 0) {
                        prior[c] = uprior; visits[c] = 0; q[c] = 0.0; child_base[c] = -1; meta[c] = cm;
                        visits[b] = m + 1;
                    }
                    if (!UNIFORM_PRIOR) {
                        uint32_t mat[5];
                        qz_mat_load(v, b, mat);
#pragma unroll
                        for (int r = 0; r < 5; r++) if (r == (urank >> 5)) mat[r] |= 1u << (urank & 31);
                        int nrank; float nprior;
                        qz_next_candidate(v.pool + child_base[b], total, mat, c_puct_f, nrank, nprior);
                        __syncwarp();
                        if (lane == 0) { qz_mat_store(v, b, mat); child_base[b + 1] = nrank; prior[b] = nprior; }
                    }
                    if (at_root) {
                        if (lane == 0) {
                            s_n[wib][m] = 0; s_meta[wib][m] = cm; s_q[wib][m] = 0.0;
                            if (!UNIFORM_PRIOR) s_prior[wib][m] = uprior;
                            if (FAST_ROOT) { s_qv[wib][m] = -1.0; s_rcp[wib][m] = 0.5; }   // n = 0, one mark: -1/1, 1/(1+0+1)
                        }
                        rm = m + 1;
                    }
                    node = c;
                    s = qz_apply(s, uact);                               // game.step(action), mcts.py:113
                    depth++;
                    if (lane == 0) path[depth] = node;
                    __syncwarp();
                    if (depth >= t.max_depth - 1) { flags |= QZ_LEAF_DEPTH_OVERFLOW; break; }
                    continue;                                            // the new slot has no block: the loop ends there
                }
            }
            if (bj < 0) bj = 0;                                          // all-NaN guard; never in practice
            node = b + QZ_HDR + bj;
            const uint32_t cm = at_root ? s_meta[wib][bj] : meta[node];
            s = qz_apply(s, qz_meta_action(cm));                         // game.step(action), mcts.py:113
            depth++;
            __syncwarp();                                                // everyone has read s_meta[bj]
            if (lane == 0) {
                path[depth] = node;
                meta[node] = cm + (1u << 16);
                if (at_root) {
                    s_meta[wib][bj] = cm + (1u << 16);
                    if (FAST_ROOT) {
                        const int n = s_n[wib][bj], infl = qz_meta_inflight(cm) + 1;
                        const double num = s_q[wib][bj] * (double)n - (double)infl;
                        s_qv[wib][bj] = num == 0.0 ? 0.0 : num / (double)(n + infl);
                        s_rcp[wib][bj] = 1.0 / (double)(1 + n + infl);
                    }
                }
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
    if (lane == 0) t.n_nodes[g] = n_nodes;
}

extern "C" int qz_mcts_select(const qz_tree *tree, double c_puct, int uniform_prior, int k_leaves, int lazy_expand,
                              int32_t *overflow_count, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_select");
    if (rc) return rc;
    QZ_REQUIRE(k_leaves >= 0 && k_leaves <= tree->leaves_per_game);
    if (!uniform_prior && tree->pool_cap <= 0) return qz_fail(QZ_E_NULL, "qz_mcts_select: stored priors need a prior pool");
    if (tree->n_games == 0) return 0;
    const unsigned blocks = qz_blocks_for(tree->n_games, 4);
    QZ_REQUIRE(!lazy_expand || uniform_prior);
    const bool fast = tree->leaves_per_game > 1;      // one leaf per wave = the reference's arithmetic, literally
    cudaStream_t st = (cudaStream_t)stream;
    if (uniform_prior && fast) qz_mcts_select_kernel<true, true><<<blocks, 128, 0, st>>>(*tree, c_puct, k_leaves, lazy_expand, overflow_count);
    else if (uniform_prior) qz_mcts_select_kernel<true, false><<<blocks, 128, 0, st>>>(*tree, c_puct, k_leaves, lazy_expand, overflow_count);
    else if (fast) qz_mcts_select_kernel<false, true><<<blocks, 128, 0, st>>>(*tree, c_puct, k_leaves, 0, overflow_count);
    else qz_mcts_select_kernel<false, false><<<blocks, 128, 0, st>>>(*tree, c_puct, k_leaves, 0, overflow_count);
    return qz_check_launch("qz_mcts_select");
}

// ------------------------------------------------------------------------------------------ extend (lazy expansion)
// Pure MCTS (uniform priors) never needs a node's legal actions until the node is visited a SECOND time: the first visit
// only rolls out from it (pure_mcts.py:75-83 expands there and then, but the children it creates stay untouched until a
// later playout descends through the node), and of a thousand playouts ~85 % end in a node that is never seen again
// before the tree is thrown away (pure_mcts.py:142).  So with lazy_expand the 128-candidate legality sweep
// (quoridor.py:420-528, the dearest thing a playout does after its rollout) is not run on every leaf: qz_mcts_select
// stops at a visited node that has no block yet (QZ_LEAF_NEEDS_MASK), and this kernel -- one warp per game, its flagged
// leaves in playout order -- computes that node's legal set, builds its block, and takes the descent one PUCT level
// further (mcts.py:37-42,64-70), exactly what the eager form would have done from a block built at the first visit.
__global__ void __launch_bounds__(128, 4) qz_mcts_extend_kernel(qz_tree t, double c_puct, const uint64_t *__restrict__ mask3,
                                                                int32_t *__restrict__ overflow_count) {
    __shared__ uint32_t scratch[4][QZ_WARP_SCRATCH_WORDS];
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    float *__restrict__ prior = v.prior;
    int32_t *__restrict__ visits = v.visits;
    double *__restrict__ q = v.q;
    int32_t *__restrict__ child_base = v.child_base;
    uint32_t *__restrict__ meta = v.meta;
    const int K = t.leaves_per_game;
    const int root = t.root[g];
    int n_nodes = t.n_nodes[g];
    bool touched = false;
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        unsigned flags = t.leaf_flags[L];
        if (!(flags & QZ_LEAF_NEEDS_MASK)) continue;
        flags &= ~QZ_LEAF_NEEDS_MASK;
        touched = true;
        const bool swept = !(flags & QZ_LEAF_DUPLICATE);               // its mask3 row was computed (see the flagged sweep)
        flags &= ~QZ_LEAF_DUPLICATE;
        const int X = qz_resolve(child_base, t.leaf_node[L]);
        QzState s = qz_load_state(t.leaf_state + L);
        const int len = t.path_len[L];
        int b = child_base[X];
        uint32_t pawn; uint64_t hl, vl;
        __syncwarp();
        if (b < 0) {                                                    // first descent to come back to X: expand it now
            // swept in parallel by qz_env_legal_mask_flagged -- unless this leaf was skipped there as the duplicate of an
            // earlier one whose node then could NOT be given a block (no legal action, arena full): swept here
            if (mask3 != nullptr && swept) {
                const uint64_t mk[3] = {mask3[3 * L], mask3[3 * L + 1], mask3[3 * L + 2]};
                qz_unpack_mask(mk, pawn, hl, vl);
            } else {
                qz_warp_legal(s, pawn, hl, vl, scratch[threadIdx.x >> 5]);  // Quoridor.actions(), quoridor.py:138-157
            }
            const int cnt = qz_popc32(pawn) + qz_popc64(hl) + qz_popc64(vl);
            const int cap = X == root ? cnt : min(cnt, QZ_CAP0);
            if (cnt == 0 || n_nodes + QZ_HDR + cap > t.node_cap) {
                // no legal action (stalemate: the leaf stays X) or no room (counted; the playout is evaluated at X)
                if (cnt != 0) { flags |= QZ_LEAF_ARENA_OVERFLOW; if (lane == 0 && overflow_count) atomicAdd(overflow_count, 1); }
                if (lane == 0) t.leaf_flags[L] = (uint8_t)flags;
                continue;
            }
            b = n_nodes;
            n_nodes += QZ_HDR + cap;
            uint64_t mk[3];
            qz_pack_mask(pawn, hl, vl, mk);
            if (lane < QZ_HDR) {
                q[b + lane] = __longlong_as_double((long long)(lane == 0 ? mk[0] : (lane == 1 ? mk[1] : mk[2])));
                meta[b + lane] = 0; prior[b + lane] = 0.0f;
                visits[b + lane] = lane == 0 ? 0 : (lane == 1 ? cap : cnt);
                child_base[b + lane] = lane == 2 ? 0 : -1;
            }
            __syncwarp();
            if (lane == 0) child_base[X] = b;
            __syncwarp();
        } else {
            qz_header_mask(v, b, pawn, hl, vl);
        }
        // one level of TreeNode.select (mcts.py:37-42) at X under uniform priors
        const int m = visits[b], total = visits[b + 2];
        const uint32_t pm = meta[X];
        const int np_eff = visits[X] + qz_meta_inflight(pm) - 1;          // minus this descent's own mark
        const double sq = sqrt((double)np_eff);
        const double uni = c_puct * (1.0 / (double)total);               // pure_mcts.py:15
        double best = -INFINITY;
        int brank = 0x7FFFFFFF, bj = -1;
        for (int j = lane; j < m; j += 32) {
            const int c = b + QZ_HDR + j;
            const uint32_t cm = meta[c];
            const int n = visits[c], infl = qz_meta_inflight(cm);
            double qv = q[c];
            if (infl > 0) {
                const double num = qv * (double)n - (double)infl;
                qv = num == 0.0 ? 0.0 : num / (double)(n + infl);
            }
            const double val = qv + uni * sq / (double)(1 + n + infl);
            const int rank = qz_meta_rank(cm);
            if (val > best || (val == best && rank < brank)) { best = val; brank = rank; bj = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(QZ_FULL_MASK, best, off);
            const int orank = __shfl_xor_sync(QZ_FULL_MASK, brank, off), oj = __shfl_xor_sync(QZ_FULL_MASK, bj, off);
            if (oj >= 0 && (bj < 0 || ob > best || (ob == best && orank < brank))) { best = ob; brank = orank; bj = oj; }
        }
        bool take_new = false;
        if (m < total) {
            const double u0 = uni * sq / (double)(1 + 0 + 0);
            take_new = bj < 0 || u0 > best || (u0 == best && m < brank);
        }
        int child = -1, act = -1;
        if (take_new) {
            act = qz_action_of_rank(pawn, hl, vl, m);
            bool room = true;
            if (m == visits[b + 1]) {                                   // block full: move it (see qz_mcts_select_kernel)
                const int newcap = min(total, max(2 * m, QZ_CAP0));
                if (n_nodes + QZ_HDR + newcap <= t.node_cap) {
                    const int nb = n_nodes;
                    n_nodes += QZ_HDR + newcap;
                    for (int j = lane; j < QZ_HDR + m; j += 32) {
                        prior[nb + j] = prior[b + j]; visits[nb + j] = visits[b + j]; q[nb + j] = q[b + j];
                        child_base[nb + j] = child_base[b + j]; meta[nb + j] = meta[b + j];
                    }
                    __syncwarp();
                    for (int j = lane; j < m; j += 32) child_base[b + QZ_HDR + j] = -2 - (nb + QZ_HDR + j);
                    if (lane == 0) { visits[nb + 1] = newcap; child_base[X] = nb; }
                    __syncwarp();
                    b = nb;
                } else {
                    room = false;
                    if (lane == 0 && overflow_count) atomicAdd(overflow_count, 1);
                }
            }
            if (room) {
                child = b + QZ_HDR + m;
                if (lane == 0) {
                    prior[child] = 1.0f / (float)total; visits[child] = 0; q[child] = 0.0; child_base[child] = -1;
                    meta[child] = (uint32_t)act | ((uint32_t)m << 8) | (1u << 16);
                    visits[b] = m + 1;
                }
            } else if (bj < 0) {
                flags |= QZ_LEAF_ARENA_OVERFLOW;                        // nowhere to go: the playout is evaluated at X
                if (lane == 0) t.leaf_flags[L] = (uint8_t)flags;
                __syncwarp();
                continue;
            }
        }
        if (child < 0) {                                                // an existing child
            if (bj < 0) bj = 0;
            child = b + QZ_HDR + bj;
            const uint32_t cm = meta[child];
            act = qz_meta_action(cm);
            __syncwarp();
            if (lane == 0) meta[child] = cm + (1u << 16);
        }
        s = qz_apply(s, act);                                            // game.step(action), mcts.py:113
        if (qz_done(s.meta)) flags |= QZ_LEAF_TERMINAL;
        if (len + 1 >= t.max_depth) flags |= QZ_LEAF_DEPTH_OVERFLOW;
        if (lane == 0) {
            if (len < t.max_depth) { t.path[L * t.max_depth + len] = child; t.path_len[L] = len + 1; }
            t.leaf_node[L] = child;
            t.leaf_flags[L] = (uint8_t)flags;
            qz_store_state(t.leaf_state + L, s);
        }
        __syncwarp();
    }
    if (touched && lane == 0) t.n_nodes[g] = n_nodes;
}

// End of a lazily expanded search: a root that was evaluated but never came back to (n_playout = 1) gets its block, so
// that root statistics and the move choice see its children (all unvisited) as the reference's would.
__global__ void __launch_bounds__(128, 4) qz_mcts_build_root_kernel(qz_tree t, int32_t *__restrict__ overflow_count) {
    __shared__ uint32_t scratch[4][QZ_WARP_SCRATCH_WORDS];
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    const int root = t.root[g];
    if (v.child_base[root] != QZ_CB_LAZY) return;
    const QzState s = qz_load_state(t.root_state + g);
    uint32_t pawn; uint64_t hl, vl;
    qz_warp_legal(s, pawn, hl, vl, scratch[threadIdx.x >> 5]);
    const int cnt = qz_popc32(pawn) + qz_popc64(hl) + qz_popc64(vl);
    const int n_nodes = t.n_nodes[g];
    if (cnt == 0) return;
    if (n_nodes + QZ_HDR + cnt > t.node_cap) { if (lane == 0 && overflow_count) atomicAdd(overflow_count, 1); return; }
    const int b = n_nodes;
    uint64_t mk[3];
    qz_pack_mask(pawn, hl, vl, mk);
    if (lane < QZ_HDR) {
        v.q[b + lane] = __longlong_as_double((long long)(lane == 0 ? mk[0] : (lane == 1 ? mk[1] : mk[2])));
        v.meta[b + lane] = 0; v.prior[b + lane] = 0.0f;
        v.visits[b + lane] = lane == 0 ? 0 : cnt;
        v.child_base[b + lane] = lane == 2 ? 0 : -1;
    }
    __syncwarp();
    if (lane == 0) { v.child_base[root] = b; t.n_nodes[g] = n_nodes + QZ_HDR + cnt; }
}

extern "C" int qz_mcts_extend(const qz_tree *tree, double c_puct, const uint64_t *mask3, int root_only, int32_t *overflow_count,
                              void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_extend");
    if (rc) return rc;
    if (tree->n_games == 0) return 0;
    if (root_only)
        qz_mcts_build_root_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*tree, overflow_count);
    else
        qz_mcts_extend_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*tree, c_puct, mask3, overflow_count);
    return qz_check_launch("qz_mcts_extend");
}

// ------------------------------------------------------------------------------------------ expand + backup
// TreeNode.expand (mcts.py:27-35) with the (action, prob) pairs of policy_value_fn (policy_value_net.py:162:
// probs[legal], NOT renormalised), then update_recursive(-leaf_value) (mcts.py:44-62,127).
// One warp per game, its leaves in order, so a game's tree is only ever touched by one warp: no atomics.
struct QzExpandArgs {
    const uint64_t *mask3;      // [n*K,3] legal masks of the leaves (qz_env_legal_mask on leaf_state); NULL = lazy expansion
    const float *priors;        // [n*K,140] or NULL (uniform 1/len, pure_mcts.py:13-16)
    const float *value_f32;     // [n*K] leaf value for the side to move (net), or NULL
    const double *value_f64;    // [n*K] same in float64 (stubs), or NULL
    const int8_t *value_i8;     // [n*K] same as +1/0/-1 (rollouts), or NULL
    int fix_terminal_sign;      // 0 = reference behaviour (a winning move backs up as a loss, mcts.py:125)
    int32_t *overflow_count;    // nullable: incremented when an expansion did not fit the arena
    float c_puct_f;             // float32 c_puct (orders the unvisited children of a net-prior node)
};

__global__ void __launch_bounds__(128) qz_mcts_expand_backup_kernel(qz_tree t, QzExpandArgs a) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree tv = qz_game_tree(t, g);
    float *__restrict__ prior = tv.prior;
    int32_t *__restrict__ visits = tv.visits;
    double *__restrict__ q = tv.q;
    int32_t *__restrict__ child_base = tv.child_base;
    uint32_t *__restrict__ meta = tv.meta;
    const int K = t.leaves_per_game;
    int n_nodes = t.n_nodes[g];
    int n_pool = t.n_pool ? t.n_pool[g] : 0;
    const int root = t.root[g];
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        unsigned flags = t.leaf_flags[L];
        if (flags & QZ_LEAF_INACTIVE) continue;
        const int node = qz_resolve(child_base, t.leaf_node[L]);
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
            if (a.mask3 != nullptr && node_cb < 0 && !(flags & (QZ_LEAF_DEPTH_OVERFLOW | QZ_LEAF_ARENA_OVERFLOW))) {
                uint32_t pawn; uint64_t hl, vl;
                const uint64_t mk[3] = {a.mask3[3 * L], a.mask3[3 * L + 1], a.mask3[3 * L + 2]};
                qz_unpack_mask(mk, pawn, hl, vl);
                const int cnt = qz_popc32(pawn) + qz_popc64(hl) + qz_popc64(vl);
                if (cnt > 0) {
                    // every child of the root is visited sooner or later; elsewhere a handful are
                    const int cap = node == root ? cnt : min(cnt, QZ_CAP0);
                    if (n_nodes + QZ_HDR + cap <= t.node_cap && (!a.priors || n_pool + cnt <= t.pool_cap)) {
                        const int b = n_nodes;
                        n_nodes += QZ_HDR + cap;
                        if (lane < QZ_HDR) {
                            q[b + lane] = __longlong_as_double((long long)mk[lane]);
                            meta[b + lane] = 0; prior[b + lane] = 0.0f;
                            visits[b + lane] = lane == 0 ? 0 : (lane == 1 ? cap : cnt);
                            child_base[b + lane] = lane == 2 ? 0 : -1;
                        }
                        if (a.priors) {
                            // the node's priors in actions() rank order: `cnt` floats of the game's prior pool
                            float *row = tv.pool + n_pool;
#pragma unroll
                            for (int r = 0; r < 5; r++) {
                                const int act = lane + 32 * r;
                                if (act < QZ_N_ACTIONS && qz_is_legal(pawn, hl, vl, act))
                                    row[qz_action_rank(pawn, hl, vl, act)] = a.priors[L * QZ_N_ACTIONS + act];
                            }
                            __syncwarp();
                            const uint32_t mat[5] = {0, 0, 0, 0, 0};
                            int nrank; float nprior;
                            qz_next_candidate(row, cnt, mat, a.c_puct_f, nrank, nprior);
                            __syncwarp();
                            if (lane == 0) { child_base[b] = n_pool; child_base[b + 1] = nrank; prior[b] = nprior; }
                            n_pool += cnt;
                        }
                        __syncwarp();
                        if (lane == 0) child_base[node] = b;
                    } else {
                        flags |= QZ_LEAF_ARENA_OVERFLOW;
                        if (lane == 0 && a.overflow_count) atomicAdd(a.overflow_count, 1);
                    }
                }
            } else if (a.mask3 == nullptr && node_cb == -1 && !(flags & QZ_LEAF_DEPTH_OVERFLOW)) {
                // lazy expansion: remember that the reference expanded this node now; its block comes when needed
                __syncwarp();
                if (lane == 0) child_base[node] = QZ_CB_LAZY;
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
                    const int nd = qz_resolve(child_base, path[d]);
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
    if (lane == 0) { t.n_nodes[g] = n_nodes; if (t.n_pool) t.n_pool[g] = n_pool; }
}

extern "C" int qz_mcts_expand_backup(const qz_tree *tree, const uint64_t *mask3, const float *priors,
                                     const float *value_f32, const double *value_f64, const int8_t *value_i8,
                                     double c_puct, int fix_terminal_sign, int32_t *overflow_count, void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_expand_backup");
    if (rc) return rc;
    if (mask3 == nullptr && priors != nullptr) return qz_fail(QZ_E_NULL, "qz_mcts_expand_backup: priors without legal masks");
    if (!value_f32 && !value_f64 && !value_i8) return qz_fail(QZ_E_NULL, "qz_mcts_expand_backup: no value array");
    if (priors && tree->pool_cap <= 0) return qz_fail(QZ_E_NULL, "qz_mcts_expand_backup: priors need a prior pool");
    if (tree->n_games == 0) return 0;
    QzExpandArgs a;
    a.mask3 = mask3; a.priors = priors; a.value_f32 = value_f32; a.value_f64 = value_f64; a.value_i8 = value_i8;
    a.fix_terminal_sign = fix_terminal_sign; a.overflow_count = overflow_count; a.c_puct_f = (float)c_puct;
    qz_mcts_expand_backup_kernel<<<qz_blocks_for(tree->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*tree, a);
    return qz_check_launch("qz_mcts_expand_backup");
}

// Back up the leaves whose rollout was deferred (QZ_LEAF_PENDING) once their values are known.  One thread per
// game, its pending leaves in order; same arithmetic as qz_mcts_expand_backup_kernel.
__global__ void qz_mcts_backup_pending_kernel(qz_tree t, const int8_t *__restrict__ value_i8, int fix_terminal_sign) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.n_games) return;
    const QzGameTree tv = qz_game_tree(t, g);
    const int K = t.leaves_per_game;
    for (int k = 0; k < K; k++) {
        const int64_t L = g * K + k;
        const unsigned flags = t.leaf_flags[L];
        if (!(flags & QZ_LEAF_PENDING)) continue;
        const double v = (flags & QZ_LEAF_TERMINAL) ? (fix_terminal_sign ? -1.0 : 1.0) : (double)value_i8[L];
        const int32_t *path = t.path + L * t.max_depth;
        const int len = t.path_len[L];
        double x = -v;
        for (int d = len - 1; d >= 0; d--) {
            const int nd = qz_resolve(tv.child_base, path[d]);
            const int nv = tv.visits[nd] + 1;
            tv.visits[nd] = nv;
            const double qo = tv.q[nd];
            tv.q[nd] = qo + 1.0 * (x - qo) / (double)nv;
            tv.meta[nd] -= (1u << 16);
            x = -x;
        }
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
// and softmax(1/temp * log(visits + 1e-10)) (mcts.py:6-9,143) in float64.  One warp per game.  Children without a
// slot are the reference's never-visited children: 0 visits, Q = 0, and they take part in the softmax.
__global__ void __launch_bounds__(128) qz_mcts_root_stats_kernel(qz_tree t, double temp, int32_t *__restrict__ visits_out,
                                                                 double *__restrict__ q_out, double *__restrict__ probs_out,
                                                                 int32_t *__restrict__ root_n_out) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    const int root = t.root[g];
    const int b = v.child_base[root];
    const int m = b < 0 ? 0 : v.visits[b], total = b < 0 ? 0 : v.visits[b + 2];
    uint32_t pawn = 0; uint64_t hl = 0, vl = 0;
    if (b >= 0) qz_header_mask(v, b, pawn, hl, vl);
    if (lane == 0 && root_n_out) root_n_out[g] = v.visits[root];
    const double x0 = 1.0 / temp * log(0.0 + 1e-10);
    double mx = total > m ? x0 : -INFINITY;
    for (int j = lane; j < m; j += 32) mx = fmax(mx, 1.0 / temp * log((double)v.visits[b + QZ_HDR + j] + 1e-10));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(QZ_FULL_MASK, mx, off));
    double sum = 0.0;
    for (int j = lane; j < m; j += 32) sum += exp(1.0 / temp * log((double)v.visits[b + QZ_HDR + j] + 1e-10) - mx);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(QZ_FULL_MASK, sum, off);
    const double e0 = total > m ? exp(x0 - mx) : 0.0;
    sum += (double)(total - m) * e0;
    for (int act = lane; act < QZ_N_ACTIONS; act += 32) {
        const bool legal = b >= 0 && qz_is_legal(pawn, hl, vl, act);
        if (visits_out) visits_out[g * QZ_N_ACTIONS + act] = 0;
        if (q_out) q_out[g * QZ_N_ACTIONS + act] = 0.0;
        if (probs_out) probs_out[g * QZ_N_ACTIONS + act] = legal ? e0 / sum : 0.0;
    }
    __syncwarp();
    for (int j = lane; j < m; j += 32) {
        const int c = b + QZ_HDR + j;
        const int act = qz_meta_action(v.meta[c]);
        const int n = v.visits[c];
        if (visits_out) visits_out[g * QZ_N_ACTIONS + act] = n;
        if (q_out) q_out[g * QZ_N_ACTIONS + act] = v.q[c];
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

// ------------------------------------------------------------------------------------------ one node's children
// The reference's `node._children` (mcts.py:19-25,33-35) for ONE node of ONE game, in actions() order, children
// without a slot included (visits 0, Q 0, their prior): what the host-side TreeNode view reads.  One warp.
__global__ void qz_mcts_node_children_kernel(qz_tree t, int64_t g, int32_t node, int32_t *__restrict__ out_i,
                                             double *__restrict__ out_d) {
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    node = qz_resolve(v.child_base, node);
    const int b = v.child_base[node];
    // out_i: [0] = count, [1] = node visits, [2] = resolved node index; then per rank r: [4+4r] action, [5+4r] slot (-1 = none),
    // [6+4r] visits, [7+4r] in-flight;  out_d: [0] = node Q, [1] = node prior; then per rank r: [2+2r] Q, [3+2r] prior
    if (lane == 0) {
        out_i[1] = v.visits[node]; out_i[2] = node; out_d[0] = v.q[node]; out_d[1] = (double)v.prior[node];
    }
    if (b < 0) { if (lane == 0) out_i[0] = 0; return; }
    const int m = v.visits[b], total = v.visits[b + 2];
    uint32_t pawn; uint64_t hl, vl;
    qz_header_mask(v, b, pawn, hl, vl);
    const int pidx = v.child_base[b];
    if (lane == 0) out_i[0] = total;
    for (int act = lane; act < QZ_N_ACTIONS; act += 32) {
        if (!qz_is_legal(pawn, hl, vl, act)) continue;
        const int r = qz_action_rank(pawn, hl, vl, act);
        out_i[4 + 4 * r] = act; out_i[5 + 4 * r] = -1; out_i[6 + 4 * r] = 0; out_i[7 + 4 * r] = 0;
        out_d[2 + 2 * r] = 0.0;
        out_d[3 + 2 * r] = (double)((pidx >= 0 && v.pool) ? v.pool[pidx + r] : 1.0f / (float)total);
    }
    __syncwarp();
    for (int j = lane; j < m; j += 32) {
        const int c = b + QZ_HDR + j;
        const uint32_t cm = v.meta[c];
        const int r = qz_meta_rank(cm);
        out_i[5 + 4 * r] = c; out_i[6 + 4 * r] = v.visits[c]; out_i[7 + 4 * r] = qz_meta_inflight(cm);
        out_d[2 + 2 * r] = v.q[c]; out_d[3 + 2 * r] = (double)v.prior[c];
    }
}

extern "C" int qz_mcts_node_children(const qz_tree *tree, int64_t game, int32_t node, int32_t *out_i, double *out_d,
                                     void *stream) {
    int rc = qz_tree_check(tree, "qz_mcts_node_children");
    if (rc) return rc;
    QZ_REQUIRE_PTR(out_i);
    QZ_REQUIRE_PTR(out_d);
    QZ_REQUIRE(game >= 0 && game < tree->n_games && node >= 0 && node < tree->node_cap);
    qz_mcts_node_children_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*tree, game, node, out_i, out_d);
    return qz_check_launch("qz_mcts_node_children");
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
    __shared__ double pbuf[4][QZ_SEL_MAX_CHILDREN];
    __shared__ uint8_t act_of_rank[4][QZ_SEL_MAX_CHILDREN];
    const int w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    if (g >= t.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree v = qz_game_tree(t, g);
    const int root = t.root[g];
    const int b = v.child_base[root];
    const int total = b < 0 ? 0 : v.visits[b + 2];
    if (total == 0) { if (lane == 0) moves_out[g] = -1; return; }      // mcts.py:195-196 ("board is full")
    const int m = v.visits[b];
    uint32_t pawn; uint64_t hl, vl;
    qz_header_mask(v, b, pawn, hl, vl);
    if (mode == 0) {
        // max(children, key=visits): the first child in actions() order among the most visited
        int bv = 0, brank = 0;                                          // a child without a slot has 0 visits: rank 0 wins at 0
        for (int j = lane; j < m; j += 32) {
            const int n = v.visits[b + QZ_HDR + j], rank = qz_meta_rank(v.meta[b + QZ_HDR + j]);
            if (n > bv || (n == bv && rank < brank)) { bv = n; brank = rank; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int ov = __shfl_xor_sync(QZ_FULL_MASK, bv, off), orank = __shfl_xor_sync(QZ_FULL_MASK, brank, off);
            if (ov > bv || (ov == bv && orank < brank)) { bv = ov; brank = orank; }
        }
        const int act = qz_action_of_rank(pawn, hl, vl, brank);
        if (lane == 0) moves_out[g] = act;
        return;
    }
    const uint64_t rid = game_id ? (uint64_t)game_id[g] : (uint64_t)g;
    const uint32_t ply = qz_ply(t.root_state[g].meta);
    for (int act = lane; act < QZ_N_ACTIONS; act += 32)
        if (qz_is_legal(pawn, hl, vl, act)) act_of_rank[w][qz_action_rank(pawn, hl, vl, act)] = (uint8_t)act;
    // probabilities exactly as root_stats, indexed by actions() rank
    const double x0 = 1.0 / temp * log(0.0 + 1e-10);
    double mx = total > m ? x0 : -INFINITY;
    for (int j = lane; j < m; j += 32) mx = fmax(mx, 1.0 / temp * log((double)v.visits[b + QZ_HDR + j] + 1e-10));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(QZ_FULL_MASK, mx, off));
    const double e0 = total > m ? exp(x0 - mx) : 0.0;
    for (int r = lane; r < total; r += 32) pbuf[w][r] = e0;
    __syncwarp();
    double sum = 0.0, gsum = 0.0;
    for (int j = lane; j < m; j += 32) {
        const int c = b + QZ_HDR + j;
        const double e = exp(1.0 / temp * log((double)v.visits[c] + 1e-10) - mx);
        pbuf[w][qz_meta_rank(v.meta[c])] = e;
        sum += e;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(QZ_FULL_MASK, sum, off);
    sum += (double)(total - m) * e0;
    double gam[5];
    if (mode == 2) {
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int j = lane + 32 * r;
            gam[r] = j < total ? qz_gamma_small(dir_alpha, seed ^ 0xD1B54A32D192ED03ull, rid, (ply << 8) | (uint32_t)j) : 0.0;
            gsum += gam[r];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gsum += __shfl_xor_sync(QZ_FULL_MASK, gsum, off);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int j = lane + 32 * r;
        if (j < total) {
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
        int pick = total - 1;
        for (int j = 0; j < total; j++) {
            acc += pbuf[w][j];
            if (u < acc) { pick = j; break; }
        }
        moves_out[g] = (int)act_of_rank[w][pick];
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
// unknown move (-1), or a child that never got a slot (never visited: the reference's node would be a fresh leaf with
// only its prior), gives a fresh root.  The kept subtree is COMPACTED from the current arena (`src`) into the other one
// (`dst`), breadth first, block by block (blocks shrink to fit, the new root's block gets room for all its children),
// so the arena never fragments.  While a dst slot waits to be processed its child_base holds the index of its src
// block.  One warp per game.  The root state advances by the move (Quoridor.step, quoridor.py:159-186).
__global__ void __launch_bounds__(128) qz_mcts_reroot_kernel(qz_tree src, qz_tree dst, const int32_t *__restrict__ moves,
                                                             int apply_move, int32_t *__restrict__ overflow_count) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= src.n_games) return;
    const int lane = threadIdx.x & 31;
    const QzGameTree sv = qz_game_tree(src, g), dv = qz_game_tree(dst, g);
    const int move = moves[g];
    QzState s = qz_load_state(src.root_state + g);
    if (apply_move && move >= 0) s = qz_apply(s, move);
    if (lane == 0) qz_store_state(dst.root_state + g, s);
    // find the child of the root that carries `move`
    const int root = src.root[g];
    const int rb = sv.child_base[root];
    const int rm = rb < 0 ? 0 : sv.visits[rb];
    int found = -1;
    for (int j = lane; j < rm; j += 32)
        if (move >= 0 && qz_meta_action(sv.meta[rb + QZ_HDR + j]) == move) found = rb + QZ_HDR + j;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) found = max(found, __shfl_xor_sync(QZ_FULL_MASK, found, off));
    if (lane == 0) { dst.root[g] = 0; if (dst.n_pool) dst.n_pool[g] = 0; }
    if (found < 0) {                                                     // fresh tree (mcts.py:150-151)
        if (lane == 0) {
            dv.prior[0] = 1.0f; dv.visits[0] = 0; dv.q[0] = 0.0; dv.child_base[0] = -1; dv.meta[0] = 0; dst.n_nodes[g] = 1;
        }
        return;
    }
    if (lane == 0) {
        dv.prior[0] = sv.prior[found];
        dv.visits[0] = sv.visits[found];
        dv.q[0] = sv.q[found];
        dv.meta[0] = sv.meta[found] & 0x0000FFFFu;                      // drop stale in-flight marks
        const int fcb = sv.child_base[found];
        dv.child_base[0] = fcb >= 0 ? fcb : (fcb == QZ_CB_LAZY ? QZ_CB_LAZY : -1);  // >= 0: src block, pending
    }
    __syncwarp();
    // breadth first over BLOCKS: the root slot first, then the child slots of every block in the order the blocks were
    // appended to dst ([cursor, cursor + 3 + cap) is one block; only its first m child slots exist)
    int tail = 1, n_pool = 0;
    bool full = false;
    int cursor = 0, first = 0, count = 1;                               // the slots [first, first + count) to look at
    for (;;) {
        for (int c0 = 0; c0 < count; c0 += 32) {
            const int i = first + c0 + lane;
            int twin = -1;
            if (c0 + lane < count) twin = dv.child_base[i];              // >= 0: src block whose children to copy
            unsigned pending = __ballot_sync(QZ_FULL_MASK, twin >= 0);
            while (pending) {
                const int l = __ffs(pending) - 1;
                pending &= pending - 1;
                const int slot = first + c0 + l;
                const int sb = __shfl_sync(QZ_FULL_MASK, twin, l);
                const int m = sv.visits[sb], total = sv.visits[sb + 2];
                const int cap = slot == 0 ? total : min(total, max(m + (m >> 1) + 1, QZ_CAP0));
                const int spool = sv.child_base[sb];
                if (full || tail + QZ_HDR + cap > dst.node_cap || (spool >= 0 && n_pool + total > dst.pool_cap)) {
                    // cannot happen with arenas of equal shape (the kept subtree is a subset); be safe: cut the subtree here
                    full = true;
                    __syncwarp();
                    if (lane == 0) { dv.child_base[slot] = -1; if (overflow_count) atomicAdd(overflow_count, 1); }
                    continue;
                }
                const int db = tail;
                if (lane < QZ_HDR) {
                    dv.q[db + lane] = sv.q[sb + lane];                   // legal mask
                    dv.meta[db + lane] = sv.meta[sb + lane];             // "has a slot" set (net priors)
                    dv.prior[db + lane] = sv.prior[sb + lane];           // candidate prior / set bits
                    dv.visits[db + lane] = lane == 1 ? cap : sv.visits[sb + lane];
                    dv.child_base[db + lane] = lane == 0 ? (spool >= 0 ? n_pool : -1) : sv.child_base[sb + lane];
                }
                if (spool >= 0) {
                    const float *srow = sv.pool + spool;
                    float *drow = dv.pool + n_pool;
                    for (int a = lane; a < total; a += 32) drow[a] = srow[a];
                    n_pool += total;
                }
                for (int j = lane; j < m; j += 32) {
                    const int sc = sb + QZ_HDR + j, dc = db + QZ_HDR + j;
                    dv.prior[dc] = sv.prior[sc];
                    dv.visits[dc] = sv.visits[sc];
                    dv.q[dc] = sv.q[sc];
                    dv.meta[dc] = sv.meta[sc] & 0x0000FFFFu;
                    const int scb = sv.child_base[sc];
                    dv.child_base[dc] = scb >= 0 ? scb : (scb == QZ_CB_LAZY ? QZ_CB_LAZY : -1);
                }
                __syncwarp();
                if (lane == 0) dv.child_base[slot] = db;
                tail += QZ_HDR + cap;
            }
            __syncwarp();
        }
        // next block of dst (blocks start at slot 1 and follow each other)
        cursor = cursor == 0 ? 1 : cursor + QZ_HDR + dv.visits[cursor + 1];
        if (cursor >= tail) break;
        first = cursor + QZ_HDR;
        count = dv.visits[cursor];
    }
    if (lane == 0) { dst.n_nodes[g] = tail; if (dst.n_pool) dst.n_pool[g] = n_pool; }
}

extern "C" int qz_mcts_reroot(const qz_tree *src, const qz_tree *dst, const int32_t *moves, int apply_move,
                              int32_t *overflow_count, void *stream) {
    int rc = qz_tree_check(src, "qz_mcts_reroot");
    if (rc) return rc;
    rc = qz_tree_check(dst, "qz_mcts_reroot");
    if (rc) return rc;
    QZ_REQUIRE_PTR(moves);
    QZ_REQUIRE(src->n_games == dst->n_games && dst->node_cap >= src->node_cap && dst->pool_cap >= src->pool_cap);
    QZ_REQUIRE(src->prior != dst->prior && src->child_base != dst->child_base);
    if (src->n_games == 0) return 0;
    qz_mcts_reroot_kernel<<<qz_blocks_for(src->n_games, 4), 128, 0, (cudaStream_t)stream>>>(*src, *dst, moves, apply_move,
                                                                                            overflow_count);
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
