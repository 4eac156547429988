// qz_sample.cuh -- one uniformly random legal action without the full 128-candidate sweep (host + device).
//
// pure_mcts.py:7-10,99 picks argmax of iid U(0,1) over actions(), i.e. a uniform legal action.  Drawing
// uniformly, WITH replacement, from the cheap SUPERSET S = {legal pawn moves} U {walls passing the prechecks
// of quoridor.py:432-461} until the drawn element is legal -- a wall is legal iff the reference's path check
// (quoridor.py:463-477) accepts it -- is the same distribution (rejection sampling) at ~1 path check per ply
// instead of 128.  The procedure is specified exactly so that every implementation of it -- the per-lane
// kernel path, the warp-per-rollout path for stuck rollouts, and the oracle (oq_sample_action, built on the
// literal actions()) -- returns the same action bit for bit:
//   S is ordered: pawn ids ascending, then H candidates by intersection, then V candidates; M = |S|.
//   attempt 0 of ply t  uses word (t & 3)       of Philox4x32-10(key = seed; ctr = (rid_lo, rid_hi, t >> 2, 0))
//   attempt j >= 1      uses word ((j - 1) & 3) of Philox4x32-10(key = seed; ctr = (rid_lo, rid_hi, t, 1 + ((j - 1) >> 2)))
//   the attempt draws S[(word * M) >> 32] and stops if it is legal.
// No legal element at all (stalemate) returns -1.  Because draws are independent of earlier outcomes, a path
// that already knows the legal set can evaluate many attempts in parallel and take the first legal one.
#pragma once
#include "qz_philox.cuh"
#include "qz_rules.cuh"

struct QzRng {
    uint64_t seed, rid;
    QzPhilox4 blk;      // cached block for plies 4q..4q+3
    uint32_t blk_q;     // which q the cache holds (0xFFFFFFFF = none)
};

QZ_HD QzRng qz_rng_init(uint64_t seed, uint64_t rid) {
    QzRng r;
    r.seed = seed; r.rid = rid; r.blk_q = 0xFFFFFFFFu;
    r.blk.x = r.blk.y = r.blk.z = r.blk.w = 0;
    return r;
}

QZ_HD uint32_t qz_rng_first_word(QzRng &r, uint32_t ply) {
    const uint32_t q = ply >> 2;
    if (q != r.blk_q) { r.blk = qz_philox(r.seed, r.rid, q, 0); r.blk_q = q; }
    return qz_philox_word(r.blk, (int)(ply & 3u));
}

QZ_HD uint32_t qz_attempt_word(QzRng &rng, uint32_t ply, uint32_t j) {
    if (j == 0) return qz_rng_first_word(rng, ply);
    const QzPhilox4 b = qz_philox(rng.seed, rng.rid, ply, 1u + ((j - 1u) >> 2));
    return qz_philox_word(b, (int)((j - 1u) & 3u));
}

// element k of the ordered superset -> action id
QZ_HD int qz_superset_action(uint32_t pmask, uint64_t hc, uint64_t vc, int npawn, int nh, int k) {
    if (k < npawn) return qz_nth_bit64((uint64_t)pmask, k);
    k -= npawn;
    if (k < nh) return 12 + qz_nth_bit64(hc, k);
    return 76 + qz_nth_bit64(vc, k - nh);
}

// Returns the action (0..139) or -1 when the mover has no legal action (stalemate); -2 when more than
// `max_rejects` DISTINCT drawn walls failed the path check (nothing is consumed: the caller redoes the ply
// another way -- see qz_rollout_stuck_kernel).
QZ_HD int qz_sample_action_capped(const QzState &s, QzRng &rng, uint32_t ply, uint32_t max_rejects) {
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    const uint32_t pmask = qz_mover_pawn_moves_ctx(c, s.meta);
    uint64_t hc = 0, vc = 0;
    if (qz_mover_walls(s.meta) > 0) { hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V); }
    const int npawn = qz_popc32(pmask), nh = qz_popc64(hc), nv = qz_popc64(vc);
    const uint32_t M = (uint32_t)(npawn + nh + nv);
    if (M == 0) return -1;
    uint64_t bad_h = 0, bad_v = 0;          // walls already found to block a path
    uint32_t n_bad = 0;
    bool prepared = false;
    QzSweep w;
    for (uint32_t j = 0;; j++) {
        const uint32_t word = qz_attempt_word(rng, ply, j);
        const int act = qz_superset_action(pmask, hc, vc, npawn, nh, (int)qz_mulhi32(word, M));
        if (act < 12) return act;
        const bool vert = act >= 76;
        const int ix = vert ? act - 76 : act - 12;
        const uint64_t bit = 1ull << ix;
        if ((vert ? bad_v : bad_h) & bit) continue;                     // drawn again: still illegal
        if (!prepared) { w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta)); prepared = true; }
        if (qz_wall_keeps_paths(w, ix, vert)) return act;
        if (vert) bad_v |= bit; else bad_h |= bit;
        n_bad++;
        if (npawn == 0 && n_bad == M) return -1;                        // every candidate is a blocking wall
        if (n_bad > max_rejects) return -2;
    }
}

QZ_HD int qz_sample_action(const QzState &s, QzRng &rng, uint32_t ply) {
    return qz_sample_action_capped(s, rng, ply, 0xFFFFFFFFu);
}

// The same attempt sequence when the legal walls (hl, vl) are already known (from a full sweep): legality is
// a table lookup.  Identical result; kept as the sequential statement of the rule (host harness cross-check) --
// qz_rollout_stuck_kernel evaluates 32 attempts at a time instead.
QZ_HD int qz_sample_action_known(const QzState &s, QzRng &rng, uint32_t ply, uint32_t pmask, uint64_t hl, uint64_t vl) {
    uint64_t hc = 0, vc = 0;
    if (qz_mover_walls(s.meta) > 0) { hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V); }
    const int npawn = qz_popc32(pmask), nh = qz_popc64(hc), nv = qz_popc64(vc);
    const uint32_t M = (uint32_t)(npawn + nh + nv);
    if (M == 0 || (npawn == 0 && (hl | vl) == 0)) return -1;
    for (uint32_t j = 0;; j++) {
        const uint32_t word = qz_attempt_word(rng, ply, j);
        const int act = qz_superset_action(pmask, hc, vc, npawn, nh, (int)qz_mulhi32(word, M));
        if (act < 12) return act;
        if (act < 76 ? (hl >> (act - 12)) & 1ull : (vl >> (act - 76)) & 1ull) return act;
    }
}

// pure_mcts.py:86-108: at most limit-1 random plies from `s`; value from the starting mover's view.
// Returns +1 / -1, or 0 when nobody has won (limit or stalemate).  `s` is advanced in place.
QZ_HD int qz_rollout(QzState &s, uint64_t seed, uint64_t rid, int limit, int &plies) {
    const int player = qz_cur(s.meta);
    QzRng rng = qz_rng_init(seed, rid);
    int steps = 0;
    for (int i = 0; i < limit; i++) {
        if (qz_done(s.meta)) break;
        if (i == limit - 1) break;
        const int a = qz_sample_action(s, rng, (uint32_t)i);
        if (a < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; break; }
        s = qz_apply(s, a);
        steps++;
    }
    plies = steps;
    const int winner = qz_winner(s.meta);
    return winner == 0 ? 0 : (winner == player ? 1 : -1);
}
