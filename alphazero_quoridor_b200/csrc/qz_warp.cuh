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
// of the reference's search graph whatever the jump rules say), found by a layered flood (cumulative layers R_k
// kept in shared memory) and a walk back from the goal row that also records every path tile's index (= its
// distance from the pawn).  "Cuts" is tested conservatively on the path's TILE set: an edge a -> b is treated as
// on the path whenever both tiles are.
//
// A candidate that does touch the path splits it into a PREFIX (tiles before the first touched one: still joined
// to the pawn, no edge between them can be cut) and a SUFFIX (tiles after the last touched one: still joined to
// the goal).  The player keeps a path iff the prefix can reach the suffix or the goal row around the wall, so the
// flood of quoridor.py:479-528 starts from the whole prefix and stops at the suffix: 2-4 iterations for a local
// detour instead of the 10-20 of a search from the pawn to the goal row.  If that closure (which equals the
// closure from the pawn) reaches neither, the jump edges decide exactly as before (qz_reach_with_jumps).
#define QZ_WITNESS_MAX_DEPTH 48
#define QZ_WITNESS_LAYER_WORDS (QZ_WITNESS_MAX_DEPTH * 3)
#define QZ_WARP_TASK_WORDS 64                                           // 256 u8 tasks
#define QZ_WARP_PIDX_WORDS 21                                           // 84 u8 path indices per player
#define QZ_WARP_SCRATCH_WORDS (2 * QZ_WITNESS_LAYER_WORDS + QZ_WARP_TASK_WORDS + 2 * QZ_WARP_PIDX_WORDS)

// Called by a whole half-warp with identical arguments: one lane stores the layers / indices, all lanes read them.
// layers[3k..3k+2] = tiles within k plain moves of `start`; pidx[t] = index of path tile t.
__device__ __forceinline__ bool qz_witness_path(const QzDirs &d, int start, int O, int player, uint32_t *layers,
                                                uint8_t *pidx, BB &path) {
    const bool writer = (threadIdx.x & 15) == 0;
    const unsigned half = (threadIdx.x & 16) ? 0xFFFF0000u : 0x0000FFFFu;
    BB reach = bb_bit(start);
    BB keep = bb_bit(O);
    keep.w0 = ~keep.w0; keep.w1 = ~keep.w1; keep.w2 = ~keep.w2;
    int depth = 0;
    path = bb_zero();
    for (;;) {
        if (writer) { layers[3 * depth] = reach.w0; layers[3 * depth + 1] = reach.w1; layers[3 * depth + 2] = reach.w2; }
        const BB nxt = qz_plain_step(d, reach, keep);
        const uint32_t hit = qz_goal_hit(nxt, player);
        if (hit) {
            // walk back: u is within depth+1 moves; any tile within `k` moves with an open move onto u continues
            // the path (its own distance is then exactly k, since u's is k+1)
            int u = player == 1 ? 64 + (__ffs(hit) - 1) : __ffs(hit) - 1;
            path = bb_bit(u);
            if (writer) pidx[u] = (uint8_t)(depth + 1);
            __syncwarp(half);
            for (int k = depth; k >= 0; k--) {
                const BB L = bb_make(layers[3 * k], layers[3 * k + 1], layers[3 * k + 2]);
                int t = -1;
                if (u >= 9 && bb_test(L, u - 9) && bb_test(d.n, u - 9)) t = u - 9;
                else if (u <= 71 && bb_test(L, u + 9) && bb_test(d.s, u + 9)) t = u + 9;
                else if (u >= 1 && bb_test(L, u - 1) && bb_test(d.e, u - 1)) t = u - 1;
                else if (u <= 79 && bb_test(L, u + 1) && bb_test(d.w, u + 1)) t = u + 1;
                if (t < 0) return false;            // cannot happen; be safe
                path = bb_or(path, bb_bit(t));
                if (writer) pidx[t] = (uint8_t)k;
                u = t;
                if (u == start) break;
            }
            __syncwarp(half);
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

// The path check of one (candidate, player) pair given the player's witness path: `d`, H, V already hold the
// candidate.  Without a witness (found == false) this is the plain search from the pawn.
__device__ __forceinline__ bool qz_detour_check(const QzDirs &d, uint64_t H, uint64_t V, int ix, bool vertical, int start, int O,
                                                int player, bool found, const BB &path, const uint32_t *layers,
                                                const uint8_t *pidx) {
    BB reach = bb_bit(start), target = bb_zero();
    if (found) {
        const int t = (ix >> 3) * 9 + (ix & 7);
        int lo = 255, hi = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int tile = t + (j == 0 ? 0 : (j == 1 ? 1 : (j == 2 ? 9 : (j == 3 ? 10 : 2))));
            if (j == 4 && !(vertical && ix < 7)) continue;           // only the row-0 quirk looks at t+2
            if (bb_at(path, tile)) {
                const int k = pidx[tile];
                lo = k < lo ? k : lo;
                hi = k > hi ? k : hi;
            }
        }
        if (lo < hi) {                                               // always: a touched edge has two path tiles
            const BB rlo = bb_make(layers[3 * lo], layers[3 * lo + 1], layers[3 * lo + 2]);
            const BB rhi = bb_make(layers[3 * hi - 3], layers[3 * hi - 2], layers[3 * hi - 1]);
            reach = bb_and(path, rlo);
            target = bb_andn(path, rhi);
        }
    }
    BB keep = bb_bit(O);
    keep.w0 = ~keep.w0; keep.w1 = ~keep.w1; keep.w2 = ~keep.w2;
    for (;;) {
        const BB nxt = qz_plain_step(d, reach, keep);
        if (bb_any(bb_and(nxt, target)) || qz_goal_hit(nxt, player)) return true;
        if (bb_eq(nxt, reach)) break;
        reach = nxt;
    }
    return qz_reach_with_jumps(d, reach, O, player, H, V);
}

// Quoridor.actions() for one game by one warp (all 32 lanes must call, with the same state).
// 1. the two witness paths are found by the two half-warps at once;  2. every candidate that passes the
// prechecks (quoridor.py:432-461) is classified -- most cut neither path and are legal at once;  3. the
// remaining (candidate, player) path checks (quoridor.py:463-477) are compacted and dealt round-robin to the
// lanes, each a short flood fill in registers.  `scratch` = QZ_WARP_SCRATCH_WORDS words of shared memory per warp.
__device__ __forceinline__ void qz_warp_legal(const QzState &s, uint32_t &pawn, uint64_t &hl, uint64_t &vl, uint32_t *scratch) {
    pawn = 0; hl = 0; vl = 0;
    if (qz_done(s.meta) || !qz_on_board(s.meta)) return;
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    pawn = qz_mover_pawn_moves_ctx(c, s.meta);
    if (qz_mover_walls(s.meta) <= 0) return;
    const int lane = threadIdx.x & 31;
    const int p1 = qz_p1(s.meta), p2 = qz_p2(s.meta);
    const QzDirs dirs = c.d;
    uint32_t *layers1 = scratch, *layers2 = scratch + QZ_WITNESS_LAYER_WORDS;
    uint8_t *tasks = reinterpret_cast<uint8_t *>(scratch + 2 * QZ_WITNESS_LAYER_WORDS);
    uint8_t *pidx1 = reinterpret_cast<uint8_t *>(scratch + 2 * QZ_WITNESS_LAYER_WORDS + QZ_WARP_TASK_WORDS);
    uint8_t *pidx2 = pidx1 + 4 * QZ_WARP_PIDX_WORDS;
    // 1. witness paths: lanes 0-15 search for P1, lanes 16-31 for P2 (same code, different data)
    const bool second = lane >= 16;
    BB mypath;
    const bool myfound = qz_witness_path(dirs, second ? p2 : p1, second ? p1 : p2, second ? 2 : 1,
                                         second ? layers2 : layers1, second ? pidx2 : pidx1, mypath);
    __syncwarp();
    BB path1, path2;
    path1.w0 = __shfl_sync(QZ_FULL_MASK, mypath.w0, 0); path1.w1 = __shfl_sync(QZ_FULL_MASK, mypath.w1, 0);
    path1.w2 = __shfl_sync(QZ_FULL_MASK, mypath.w2, 0);
    path2.w0 = __shfl_sync(QZ_FULL_MASK, mypath.w0, 16); path2.w1 = __shfl_sync(QZ_FULL_MASK, mypath.w1, 16);
    path2.w2 = __shfl_sync(QZ_FULL_MASK, mypath.w2, 16);
    const bool found1 = __shfl_sync(QZ_FULL_MASK, (int)myfound, 0), found2 = __shfl_sync(QZ_FULL_MASK, (int)myfound, 16);
    // 2. classify the candidates (lane = intersection, four passes: H 0-31, H 32-63, V 0-31, V 32-63); queue the
    //    path checks that are really needed as task bytes ix | vertical << 6 | (player - 1) << 7
    const uint64_t hc = qz_hcand(s.H, s.V), vc = qz_vcand(s.H, s.V);
    int ntasks = 0;
#pragma unroll
    for (int pass = 0; pass < 4; pass++) {
        const bool vert = pass >= 2;
        const int ix = ((pass & 1) << 5) | lane;
        const bool cand = ((vert ? vc : hc) >> ix) & 1ull;
        const bool need1 = cand && (!found1 || qz_wall_touches(path1, ix, vert));
        const bool need2 = cand && (!found2 || qz_wall_touches(path2, ix, vert));
        const unsigned b1 = __ballot_sync(QZ_FULL_MASK, need1), b2 = __ballot_sync(QZ_FULL_MASK, need2);
        const unsigned lt = (1u << lane) - 1u;
        const uint8_t code = (uint8_t)(ix | (vert ? 64 : 0));
        if (need1) tasks[ntasks + __popc(b1 & lt)] = code;
        if (need2) tasks[ntasks + __popc(b1) + __popc(b2 & lt)] = (uint8_t)(code | 128);
        ntasks += __popc(b1) + __popc(b2);
    }
    __syncwarp();
    // 3. the path checks that are left
    uint64_t fail_h = 0, fail_v = 0;
    for (int i = lane; i < ntasks; i += 32) {
        const int task = tasks[i], ix = task & 63;
        const bool vert = task & 64, for2 = task & 128;
        QzDirs d = dirs;
        uint64_t H = s.H, V = s.V;
        if (vert) { qz_dirs_place_v(d, ix); V |= 1ull << ix; } else { qz_dirs_place_h(d, ix); H |= 1ull << ix; }
        const bool ok = for2 ? qz_detour_check(d, H, V, ix, vert, p2, p1, 2, found2, path2, layers2, pidx2)
                             : qz_detour_check(d, H, V, ix, vert, p1, p2, 1, found1, path1, layers1, pidx1);
        if (!ok) { if (vert) fail_v |= 1ull << ix; else fail_h |= 1ull << ix; }
    }
    hl = hc & ~qz_warp_or64(fail_h);
    vl = vc & ~qz_warp_or64(fail_v);
    __syncwarp();
}
