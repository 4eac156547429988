// qz_warp.cuh -- warp-cooperative pieces built on qz_rules.cuh (device only).
#pragma once
#include "qz_rules.cuh"

#define QZ_FULL_MASK 0xFFFFFFFFu

__device__ __forceinline__ uint64_t qz_warp_or64(uint64_t x) {
    uint32_t lo = __reduce_or_sync(QZ_FULL_MASK, (uint32_t)x);
    uint32_t hi = __reduce_or_sync(QZ_FULL_MASK, (uint32_t)(x >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// Quoridor.actions() for one game by one warp (all 32 lanes must call, with the same state).
// The <=128 wall candidates that survive the prechecks (quoridor.py:432-461) are compacted and dealt
// round-robin to the lanes; each lane runs the two flood fills of _blocks_path (quoridor.py:463-477) for
// its candidates in registers, and the per-lane verdicts are OR-reduced.  Every lane returns the full result.
__device__ __forceinline__ void qz_warp_legal(const QzState &s, uint32_t &pawn, uint64_t &hl, uint64_t &vl) {
    pawn = 0; hl = 0; vl = 0;
    if (qz_done(s.meta) || !qz_on_board(s.meta)) return;
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    pawn = qz_mover_pawn_moves_ctx(c, s.meta);
    if (qz_mover_walls(s.meta) <= 0) return;
    const int lane = threadIdx.x & 31;
    const QzSweep w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta));
    const uint64_t hc = qz_hcand(s.H, s.V), vc = qz_vcand(s.H, s.V);
    const int nh = qz_popc64(hc), total = nh + qz_popc64(vc);
    uint64_t myh = 0, myv = 0;
    for (int k = lane; k < total; k += 32) {
        const bool vert = k >= nh;
        const int ix = qz_nth_bit64(vert ? vc : hc, vert ? k - nh : k);
        if (qz_wall_keeps_paths(w, ix, vert)) {
            if (vert) myv |= 1ull << ix; else myh |= 1ull << ix;
        }
    }
    hl = qz_warp_or64(myh);
    vl = qz_warp_or64(myv);
}
