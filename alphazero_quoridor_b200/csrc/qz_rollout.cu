// qz_rollout.cu -- K4: random rollouts to terminal, one thread per rollout, persistent with refill.
// Replaces pure_mcts.MCTS._evaluate_rollout (pure_mcts.py:86-108) + rollout_policy_fn (:7-10); also serves
// BASELINE config 1 (uniform-random legal play from reset()).
//
// Work distribution: rollouts have very uneven lengths (tens to ~1000 plies), so a thread that finishes
// immediately claims the next unstarted rollout from a global counter (warp-aggregated atomicAdd) and the
// warp stays converged at the top of one "advance every live lane by one ply" loop.  No tensor cores, no
// shared memory: the whole game lives in registers (24 B) and HBM sees 24 B in + 5..29 B out per ROLLOUT.
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
    unsigned long long *counter;   // work counter (zeroed by the launcher)
};

__global__ void __launch_bounds__(128) qz_rollout_kernel(QzRolloutArgs a) {
    const int lane = threadIdx.x & 31;
    QzState s;
    QzRng rng;
    int64_t r = -1;        // rollout this lane is running, -1 = none
    int player0 = 0, steps = 0;
    unsigned long long my_plies = 0;
    bool exhausted = false;
    s.H = s.V = s.meta = 0;
    rng = qz_rng_init(0, 0);
    for (;;) {
        // ---- refill idle lanes (warp-aggregated claim) ----
        const bool want = (r < 0) && !exhausted;
        const unsigned wmask = __ballot_sync(QZ_FULL_MASK, want);
        if (wmask) {
            unsigned long long base = 0;
            const int leader = __ffs(wmask) - 1;
            if (lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(wmask));
            base = __shfl_sync(QZ_FULL_MASK, base, leader);
            if (want) {
                const int64_t cand = (int64_t)base + __popc(wmask & ((1u << lane) - 1u));
                if (cand < a.n_rollouts) {
                    r = cand;
                    const int64_t si = a.state_index ? (int64_t)__ldg(a.state_index + r) : r / a.per_state;
                    s = qz_load_state(a.states + si);
                    rng = qz_rng_init(a.seed, a.rids ? __ldg(a.rids + r) : a.rid_base + (uint64_t)r);
                    player0 = qz_cur(s.meta);
                    steps = 0;
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(QZ_FULL_MASK, r < 0)) break;
        // ---- advance every live lane by one ply (pure_mcts.py:88-101) ----
        if (r >= 0) {
            bool finished = qz_done(s.meta) || steps >= a.limit - 1;
            if (!finished) {
                const int act = qz_sample_action(s, rng, (uint32_t)steps);
                if (act < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; finished = true; }
                else { s = qz_apply(s, act); steps++; finished = qz_done(s.meta) || steps >= a.limit - 1; }
            }
            if (finished) {
                const int winner = qz_winner(s.meta);
                a.result[r] = (int8_t)(winner == 0 ? 0 : (winner == player0 ? 1 : -1));     // :104-108
                if (a.plies) a.plies[r] = steps;
                if (a.final_states) qz_store_state(a.final_states + r, s);
                my_plies += (unsigned long long)steps;
                r = -1;
            }
        }
    }
    // cumulative env-step counter (workspace[1]; never zeroed by the library)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) my_plies += __shfl_xor_sync(QZ_FULL_MASK, my_plies, off);
    if (lane == 0 && my_plies) atomicAdd(a.counter + 1, my_plies);
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
    cudaError_t e = cudaMemsetAsync(workspace, 0, 8, st);
    if (e != cudaSuccess) return qz_fail((int)e, "qz_rollout: memset: %s", cudaGetErrorString(e));
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    // persistent grid: enough 128-thread blocks to fill every SM at the kernel's register footprint
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qz_rollout_kernel, 128, 0);
    if (e != cudaSuccess || per_sm <= 0) per_sm = 4;
    int64_t blocks = (int64_t)sms * per_sm;
    const int64_t needed = (n_rollouts + 127) / 128;
    if (blocks > needed) blocks = needed;
    QzRolloutArgs a;
    a.states = states; a.state_index = state_index; a.rids = rids; a.rid_base = rid_base; a.seed = seed;
    a.n_rollouts = n_rollouts; a.per_state = per_state > 0 ? per_state : 1; a.limit = limit;
    a.result = result; a.plies = plies; a.final_states = final_states;
    a.counter = (unsigned long long *)workspace;
    qz_rollout_kernel<<<(unsigned)blocks, 128, 0, st>>>(a);
    return qz_check_launch("qz_rollout");
}
