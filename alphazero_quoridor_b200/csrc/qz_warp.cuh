// qz_warp.cuh -- warp-cooperative pieces built on qz_rules.cuh (device only).
#pragma once
#include "qz_rules.cuh"

#define QZ_FULL_MASK 0xFFFFFFFFu

__device__ __forceinline__ uint64_t qz_warp_or64(uint64_t x) {
    uint32_t lo = __reduce_or_sync(QZ_FULL_MASK, (uint32_t)x);
    uint32_t hi = __reduce_or_sync(QZ_FULL_MASK, (uint32_t)(x >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// ---- witness paths -------------------------------------------------------------------------------------------
// Adding a wall only removes edges (every rule of quoridor.py:272-353 is a "no wall here" test), so a candidate
// that cuts no edge of some path a player ALREADY has cannot block that player: its path check is skipped.
// A witness path is a shortest plain-move path that avoids the opponent's tile (so every edge of it is an edge
// of the reference's search graph whatever the jump rules say), found by a layered flood (layers kept in shared
// memory) and a walk back from the goal row.  "Cuts" is tested conservatively on the path's TILE set: an edge
// a -> b is treated as on the path whenever both tiles are.
#define QZ_WITNESS_MAX_DEPTH 48
#define QZ_WARP_SCRATCH_WORDS (2 * QZ_WITNESS_MAX_DEPTH * 3 + 128)     // two layer stacks + 256 u16 tasks

// Called by a whole half-warp with identical arguments: one lane stores the layers, all lanes read them back.
__device__ __forceinline__ bool qz_witness_path(const QzDirs &d, int start, int O, int player, uint32_t *layers, BB &path) {
    const bool writer = (threadIdx.x & 15) == 0;
    const unsigned half = (threadIdx.x & 16) ? 0xFFFF0000u : 0x0000FFFFu;
    BB reach = bb_bit(start);
    BB keep = bb_bit(O);
    keep.w0 = ~keep.w0; keep.w1 = ~keep.w1; keep.w2 = ~keep.w2;
    int depth = 0;
    path = bb_zero();
    for (;;) {
        if (writer) { layers[3 * depth] = reach.w0; layers[3 * depth + 1] = reach.w1; layers[3 * depth + 2] = reach.w2; }
        const BB a = bb_shl(bb_and(reach, d.n), 9), b = bb_shr(bb_and(reach, d.s), 9);
        const BB c = bb_shl(bb_and(reach, d.e), 1), e = bb_shr(bb_and(reach, d.w), 1);
        BB nxt;
        nxt.w0 = reach.w0 | ((a.w0 | b.w0 | c.w0 | e.w0) & keep.w0);
        nxt.w1 = reach.w1 | ((a.w1 | b.w1 | c.w1 | e.w1) & keep.w1);
        nxt.w2 = reach.w2 | ((a.w2 | b.w2 | c.w2 | e.w2) & keep.w2);
        const uint32_t hit = player == 1 ? (nxt.w2 & QZ_ROW8_W2) : (nxt.w0 & QZ_ROW0_W0);
        if (hit) {
            // walk back: u is on layer depth+1, find a tile of layer `depth` with an open move onto u
            int u = player == 1 ? 64 + (__ffs(hit) - 1) : __ffs(hit) - 1;
            path = bb_bit(u);
            __syncwarp(half);
            for (int k = depth; k >= 0; k--) {
                // tiles at distance exactly k: a tile at distance k+1 always has a predecessor among them
                BB L = bb_make(layers[3 * k], layers[3 * k + 1], layers[3 * k + 2]);
                if (k > 0) L = bb_andn(L, bb_make(layers[3 * k - 3], layers[3 * k - 2], layers[3 * k - 1]));
                int t = -1;
                if (u >= 9 && bb_test(L, u - 9) && bb_test(d.n, u - 9)) t = u - 9;
                else if (u <= 71 && bb_test(L, u + 9) && bb_test(d.s, u + 9)) t = u + 9;
                else if (u >= 1 && bb_test(L, u - 1) && bb_test(d.e, u - 1)) t = u - 1;
                else if (u <= 79 && bb_test(L, u + 1) && bb_test(d.w, u + 1)) t = u + 1;
                if (t < 0) return false;            // cannot happen; be safe
                path = bb_or(path, bb_bit(t));
                u = t;
                if (u == start) break;
            }
            return true;
        }
        if (bb_eq(nxt, reach) || depth + 1 >= QZ_WITNESS_MAX_DEPTH) return false;   // no plain path (or very deep): no witness
        reach = nxt;
        depth++;
    }
}

// does a wall at intersection ix possibly remove an edge whose two tiles both lie on `path`?
__device__ __forceinline__ bool qz_wall_touches(const BB &path, int ix, bool vertical) {
    const int r = ix >> 3, c = ix & 7, t = r * 9 + c;
    const uint32_t a = bb_at(path, t), b = bb_at(path, t + 1), cc = bb_at(path, t + 9), dd = bb_at(path, t + 10);
    if (!vertical) return (a & cc) | (b & dd);                       // N/S edges t<->t+9, t+1<->t+10 (covers the row-0 quirk)
    uint32_t hit = (a & b) | (cc & dd);                              // E/W edges t<->t+1, t+9<->t+10
    if (r == 0 && c < 7) hit |= b & bb_at(path, t + 2);              // row-0 quirk: E from (0,c+1) tests this wall (:388,:392)
    return hit;
}

// Quoridor.actions() for one game by one warp (all 32 lanes must call, with the same state).
// 1. the two witness paths are found by the two half-warps at once;  2. every candidate that passes the
// prechecks (quoridor.py:432-461) is classified -- most cut neither path and are legal at once;  3. the
// remaining (candidate, player) path checks (quoridor.py:463-477) are compacted and dealt round-robin to the
// lanes, each a flood fill in registers.  `scratch` = QZ_WARP_SCRATCH_WORDS words of shared memory per warp.
__device__ __forceinline__ void qz_warp_legal(const QzState &s, uint32_t &pawn, uint64_t &hl, uint64_t &vl, uint32_t *scratch) {
    pawn = 0; hl = 0; vl = 0;
    if (qz_done(s.meta) || !qz_on_board(s.meta)) return;
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    pawn = qz_mover_pawn_moves_ctx(c, s.meta);
    if (qz_mover_walls(s.meta) <= 0) return;
    const int lane = threadIdx.x & 31;
    const int p1 = qz_p1(s.meta), p2 = qz_p2(s.meta);
    const QzDirs dirs = c.d;
    // 1. witness paths: lanes 0-15 search for P1, lanes 16-31 for P2 (same code, different data)
    const bool second = lane >= 16;
    BB mypath;
    const bool myfound = qz_witness_path(dirs, second ? p2 : p1, second ? p1 : p2, second ? 2 : 1,
                                         scratch + (second ? QZ_WITNESS_MAX_DEPTH * 3 : 0), mypath);
    __syncwarp();
    BB path1, path2;
    path1.w0 = __shfl_sync(QZ_FULL_MASK, mypath.w0, 0); path1.w1 = __shfl_sync(QZ_FULL_MASK, mypath.w1, 0);
    path1.w2 = __shfl_sync(QZ_FULL_MASK, mypath.w2, 0);
    path2.w0 = __shfl_sync(QZ_FULL_MASK, mypath.w0, 16); path2.w1 = __shfl_sync(QZ_FULL_MASK, mypath.w1, 16);
    path2.w2 = __shfl_sync(QZ_FULL_MASK, mypath.w2, 16);
    const bool found1 = __shfl_sync(QZ_FULL_MASK, (int)myfound, 0), found2 = __shfl_sync(QZ_FULL_MASK, (int)myfound, 16);
    // 2. classify the candidates; queue the path checks that are really needed
    const uint64_t hc = qz_hcand(s.H, s.V), vc = qz_vcand(s.H, s.V);
    const int nh = qz_popc64(hc), total = nh + qz_popc64(vc);
    uint16_t *tasks = reinterpret_cast<uint16_t *>(scratch + 2 * QZ_WITNESS_MAX_DEPTH * 3);
    int ntasks = 0;
    for (int base = 0; base < total; base += 32) {
        const int k = base + lane;
        bool need1 = false, need2 = false;
        if (k < total) {
            const bool vert = k >= nh;
            const int ix = qz_nth_bit64(vert ? vc : hc, vert ? k - nh : k);
            need1 = !found1 || qz_wall_touches(path1, ix, vert);
            need2 = !found2 || qz_wall_touches(path2, ix, vert);
        }
        const unsigned b1 = __ballot_sync(QZ_FULL_MASK, need1), b2 = __ballot_sync(QZ_FULL_MASK, need2);
        const unsigned lt = (1u << lane) - 1u;
        if (need1) tasks[ntasks + __popc(b1 & lt)] = (uint16_t)(k << 1);
        if (need2) tasks[ntasks + __popc(b1) + __popc(b2 & lt)] = (uint16_t)((k << 1) | 1);
        ntasks += __popc(b1) + __popc(b2);
    }
    __syncwarp();
    // 3. the path checks that are left
    uint64_t fail_h = 0, fail_v = 0;
    for (int i = lane; i < ntasks; i += 32) {
        const int task = tasks[i], k = task >> 1, player = (task & 1) + 1;
        const bool vert = k >= nh;
        const int ix = qz_nth_bit64(vert ? vc : hc, vert ? k - nh : k);
        QzDirs d = dirs;
        uint64_t H = s.H, V = s.V;
        if (vert) { qz_dirs_place_v(d, ix); V |= 1ull << ix; } else { qz_dirs_place_h(d, ix); H |= 1ull << ix; }
        const bool ok = player == 1 ? qz_reaches_goal(d, p1, p2, 1, H, V) : qz_reaches_goal(d, p2, p1, 2, H, V);
        if (!ok) { if (vert) fail_v |= 1ull << ix; else fail_h |= 1ull << ix; }
    }
    hl = hc & ~qz_warp_or64(fail_h);
    vl = vc & ~qz_warp_or64(fail_v);
    __syncwarp();
}
