// qz_rollout.cu -- K4: random rollouts to terminal, one thread per rollout, persistent with refill.
// Replaces pure_mcts.MCTS._evaluate_rollout (pure_mcts.py:86-108) + rollout_policy_fn (:7-10); also serves
// BASELINE config 1 (uniform-random legal play from reset()).
//
// Work distribution: rollouts have very uneven lengths (tens to ~1000 plies), so a thread that finishes
// immediately claims the next unstarted rollout from a global counter (warp-aggregated atomicAdd) and the
// warp stays converged at the top of one "advance every live lane by one ply" loop.  A rollout is cut in two
// PHASES run by two kernels -- plies while walls remain (flood fills, ~2000 instructions each) and pawn-only
// plies (~150 instructions each) -- so that the lanes of a warp always execute the same kind of ply; the
// first capture (profiles/r1a_rollout_ncu_full.txt) of the single-kernel form showed 4.2 active lanes per
// instruction because a warp mixed both kinds.  No tensor cores, no shared memory: the game lives in
// registers and HBM sees 24 B in, 48 B through `mid`, and 1..29 B out per ROLLOUT.
#include "qz_common.cuh"
#include "qz_sample.cuh"
#include "qz_warp.cuh"

struct QzRolloutArgs {
    const qz_state *states;        // [n_states]
    const int32_t *state_index;    // nullable: rollout r starts from states[state_index[r]] (else r / per_state)
    const uint64_t *rids;          // nullable: RNG id of rollout r (else rid_base + r)
    uint64_t rid_base;
    uint64_t seed;
    int64_t n_rollouts;
    int32_t per_state;
    int32_t limit;
    int8_t *result;                // [n_rollouts] +1/-1/0 from the starting mover's view
    int32_t *plies;                // nullable [n_rollouts]
    qz_state *final_states;        // nullable [n_rollouts]
    unsigned long long *counter;   // [0] phase-1 work counter, [1] cumulative plies, [2] phase-2 work counter
    qz_state *mid;                 // [n_rollouts] state at the end of the wall phase
};

// Warp-aggregated claim of the next unstarted rollout for every idle lane; returns -1 when none is left.
__device__ __forceinline__ int64_t qz_claim(unsigned long long *counter, bool want, int64_t total) {
    const int lane = threadIdx.x & 31;
    const unsigned wmask = __ballot_sync(QZ_FULL_MASK, want);
    if (!wmask) return -1;
    unsigned long long base = 0;
    const int leader = __ffs(wmask) - 1;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(wmask));
    base = __shfl_sync(QZ_FULL_MASK, base, leader);
    if (!want) return -1;
    const int64_t cand = (int64_t)base + __popc(wmask & ((1u << lane) - 1u));
    return cand < total ? cand : -1;
}

__device__ __forceinline__ int64_t qz_start_index(const QzRolloutArgs &a, int64_t r) {
    return a.state_index ? (int64_t)__ldg(a.state_index + r) : r / a.per_state;
}

// ---- phase 1: the wall phase ------------------------------------------------------------------------------
// While either player still owns a wall almost every ply draws a wall and runs the two flood fills of the
// path check.  Every lane of every warp of this kernel is in that phase, so the expensive code is executed
// by full warps; a rollout leaves the kernel as soon as no wall is left (or the game / the limit ends) and its
// state is parked in `mid` for phase 2.
//
// A rollout whose mover keeps drawing walls that fail the path check ("stuck": walls in hand, next to no
// legal placement -- under 1% of rollouts but 60% of all path checks when every draw costs two flood fills)
// is switched to warp-cooperative plies: its state is broadcast, the whole warp computes the exact legal set
// with one 128-candidate sweep (qz_warp_legal), and the owner lane replays the SAME draw sequence against
// that table.  Same action as the per-lane path, bit for bit.
#define QZ_MAX_REJECTS 2u
#define QZ_MAX_SWEEPS_PER_ITER 6      // more stuck lanes than this in one warp: cheaper to let every lane grind on

// out of line: keeps the second copy of the flood-fill code (and its registers) out of the per-lane hot loop
__device__ __noinline__ void qz_warp_legal_call(uint64_t H, uint64_t V, uint64_t meta, uint32_t *pawn, uint64_t *hl,
                                                uint64_t *vl) {
    QzState t;
    t.H = H; t.V = V; t.meta = meta;
    uint32_t p; uint64_t a, b;
    qz_warp_legal(t, p, a, b);
    *pawn = p; *hl = a; *vl = b;
}

__global__ void __launch_bounds__(128, 4) qz_rollout_wall_kernel(QzRolloutArgs a) {
    const int lane = threadIdx.x & 31;
    QzState s;
    QzRng rng;
    int64_t r = -1;
    int steps = 0;
    bool exhausted = false, stuck = false;
    s.H = s.V = s.meta = 0;
    rng = qz_rng_init(0, 0);
    for (;;) {
        const bool want = (r < 0) && !exhausted;
        const int64_t got = qz_claim(a.counter, want, a.n_rollouts);
        if (want) {
            if (got >= 0) {
                r = got;
                s = qz_load_state(a.states + qz_start_index(a, r));
                rng = qz_rng_init(a.seed, a.rids ? __ldg(a.rids + r) : a.rid_base + (uint64_t)r);
                steps = 0;
                stuck = false;
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(QZ_FULL_MASK, r < 0)) break;
        bool leave = false, need_sweep = false;
        if (r >= 0) leave = qz_done(s.meta) || steps >= a.limit - 1 || (qz_w1(s.meta) + qz_w2(s.meta)) == 0;
        const bool live = r >= 0 && !leave;
        // lanes already known to be stuck go straight to the warp sweep -- unless the warp is full of them
        const bool pre = live && stuck && qz_mover_walls(s.meta) > 0;
        const unsigned pre_mask = __ballot_sync(QZ_FULL_MASK, pre);
        const bool crowded = __popc(pre_mask) > QZ_MAX_SWEEPS_PER_ITER;
        if (live) {
            if (pre && !crowded) {
                need_sweep = true;
            } else {
                const int act = qz_sample_action_capped(s, rng, (uint32_t)steps, crowded ? 0xFFFFFFFFu : QZ_MAX_REJECTS);
                if (act == -2) { stuck = true; need_sweep = true; }
                else if (act < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; leave = true; }
                else { s = qz_apply(s, act); steps++; }
            }
        }
        unsigned sweep_mask = __ballot_sync(QZ_FULL_MASK, need_sweep);
        while (sweep_mask) {
            const int l = __ffs(sweep_mask) - 1;
            sweep_mask &= sweep_mask - 1;
            QzState t;
            t.H = __shfl_sync(QZ_FULL_MASK, s.H, l);
            t.V = __shfl_sync(QZ_FULL_MASK, s.V, l);
            t.meta = __shfl_sync(QZ_FULL_MASK, s.meta, l);
            uint32_t pawn; uint64_t hl, vl;
            qz_warp_legal_call(t.H, t.V, t.meta, &pawn, &hl, &vl);
            if (lane == l) {
                const int act = qz_sample_action_known(s, rng, (uint32_t)steps, pawn, hl, vl);
                if (act < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; leave = true; }
                else { s = qz_apply(s, act); steps++; }
            }
        }
        if (r >= 0 && leave) {
            qz_store_state(a.mid + r, s);
            r = -1;
        }
    }
}

// ---- phase 2: the pawn phase ------------------------------------------------------------------------------
// No wall can be placed any more, so the twelve corner masks of the position are built ONCE per rollout and
// every remaining ply is a branch-free pawn-move query, one Philox word, a nth-set-bit pick and a position
// update: all 32 lanes of a warp run the same ~150 instructions per ply whatever tile they stand on.
__global__ void __launch_bounds__(128) qz_rollout_pawn_kernel(QzRolloutArgs a) {
    const int lane = threadIdx.x & 31;
    QzState s;
    QzRng rng;
    QzPawnCtx ctx;
    int64_t r = -1;
    int player0 = 0, steps = 0;
    unsigned long long my_plies = 0;
    bool exhausted = false;
    s.H = s.V = s.meta = 0;
    rng = qz_rng_init(0, 0);
    ctx = qz_ctx_build(0, 0);
    for (;;) {
        const bool want = (r < 0) && !exhausted;
        const int64_t got = qz_claim(a.counter + 2, want, a.n_rollouts);
        if (want) {
            if (got >= 0) {
                r = got;
                s = qz_load_state(a.mid + r);
                const uint64_t m0 = __ldg(reinterpret_cast<const uint64_t *>(a.states + qz_start_index(a, r)) + 2);
                player0 = qz_cur(m0);
                steps = (int)qz_ply(s.meta) - (int)qz_ply(m0);
                rng = qz_rng_init(a.seed, a.rids ? __ldg(a.rids + r) : a.rid_base + (uint64_t)r);
                ctx = qz_ctx_build(s.H, s.V);
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(QZ_FULL_MASK, r < 0)) break;
        if (r >= 0) {
            const unsigned fl = qz_flags(s.meta);
            bool finished = (fl & (QZ_FLAG_DONE | QZ_FLAG_STALEMATE)) || steps >= a.limit - 1 ||
                            (qz_w1(s.meta) + qz_w2(s.meta)) != 0;   // walls left here only if phase 1 ended the rollout
            if (!finished) {
                const uint32_t pm = qz_mover_pawn_moves_ctx(ctx, s.meta);
                const int np = __popc(pm);
                if (np == 0) {
                    s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40;
                    finished = true;
                } else {
                    const uint32_t word = qz_rng_first_word(rng, (uint32_t)steps);
                    const int k = (int)__umulhi(word, (uint32_t)np);
                    const int act = (int)__fns(pm, 0, k + 1);
                    s = qz_apply(s, act);
                    steps++;
                    finished = qz_done(s.meta) || steps >= a.limit - 1;
                }
            }
            if (finished) {
                const int winner = qz_winner(s.meta);
                a.result[r] = (int8_t)(winner == 0 ? 0 : (winner == player0 ? 1 : -1));     // pure_mcts.py:104-108
                if (a.plies) a.plies[r] = steps;
                if (a.final_states) qz_store_state(a.final_states + r, s);
                my_plies += (unsigned long long)steps;
                r = -1;
            }
        }
    }
    // cumulative env-step counter (workspace word 1; never zeroed by the library)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) my_plies += __shfl_xor_sync(QZ_FULL_MASK, my_plies, off);
    if (lane == 0 && my_plies) atomicAdd(a.counter + 1, my_plies);
}

extern "C" int64_t qz_rollout_workspace_bytes(int64_t n_rollouts) {
    return 32 + (n_rollouts > 0 ? n_rollouts : 0) * (int64_t)sizeof(qz_state);
}

static int qz_persistent_blocks(const void *kernel, int64_t n_rollouts) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
    int64_t blocks = (int64_t)sms * per_sm;                 // one resident wave: a multiple of the SM count
    const int64_t needed = (n_rollouts + 127) / 128;
    return (int)(blocks < needed ? blocks : needed);
}

extern "C" int qz_rollout(const qz_state *states, int64_t n_states, const int32_t *state_index, int32_t per_state,
                          int64_t n_rollouts, uint64_t seed, uint64_t rid_base, const uint64_t *rids, int32_t limit,
                          int8_t *result, int32_t *plies, qz_state *final_states, void *workspace, void *stream) {
    QZ_REQUIRE(n_rollouts >= 0 && n_states >= 0 && limit >= 1);
    if (n_rollouts == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(result);
    QZ_REQUIRE_PTR(workspace);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(workspace, 8);
    QZ_REQUIRE_ALIGN(final_states, 8);
    if (state_index == nullptr) QZ_REQUIRE(per_state >= 1 && n_rollouts <= n_states * (int64_t)per_state);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *ctr = (unsigned long long *)workspace;
    cudaError_t e = cudaMemsetAsync(ctr, 0, 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctr + 2, 0, 8, st);
    if (e != cudaSuccess) return qz_fail((int)e, "qz_rollout: memset: %s", cudaGetErrorString(e));
    QzRolloutArgs a;
    a.states = states; a.state_index = state_index; a.rids = rids; a.rid_base = rid_base; a.seed = seed;
    a.n_rollouts = n_rollouts; a.per_state = per_state > 0 ? per_state : 1; a.limit = limit;
    a.result = result; a.plies = plies; a.final_states = final_states;
    a.counter = ctr;
    a.mid = (qz_state *)((char *)workspace + 32);
    qz_rollout_wall_kernel<<<qz_persistent_blocks((const void *)qz_rollout_wall_kernel, n_rollouts), 128, 0, st>>>(a);
    int rc = qz_check_launch("qz_rollout (wall phase)");
    if (rc) return rc;
    qz_rollout_pawn_kernel<<<qz_persistent_blocks((const void *)qz_rollout_pawn_kernel, n_rollouts), 128, 0, st>>>(a);
    return qz_check_launch("qz_rollout (pawn phase)");
}
