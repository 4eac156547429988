// qz_sample.cuh -- one uniformly random legal action without the full 128-candidate sweep (host + device).
//
// pure_mcts.py:7-10,99 picks argmax of iid U(0,1) over actions(), i.e. a uniform legal action.  Drawing
// uniformly from the cheap SUPERSET {legal pawn moves} U {walls passing the prechecks of quoridor.py:432-461}
// and accepting a wall only if the reference's path check (quoridor.py:463-477) accepts it is the same
// distribution (rejection sampling); rejected candidates are removed before the redraw.  The procedure
// is specified exactly so the oracle (oracle/quoridor_oracle.c: oq_sample_action) reproduces it bit for bit:
//   attempt 0 of ply t : word (t & 3) of Philox(seed; rid, t >> 2, 0)
//   attempt j >= 1     : word 0       of Philox(seed; rid, t, j)
//   index = (word * M) >> 32 over the M remaining candidates ordered pawn ids, H by ix, V by ix.
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

// Returns the action (0..139) or -1 when the mover has no legal action (stalemate); -2 when more than
// `max_rejects` drawn walls failed the path check (nothing is consumed: the caller redoes the ply another way).
QZ_HD int qz_sample_action_capped(const QzState &s, QzRng &rng, uint32_t ply, uint32_t max_rejects) {
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    uint32_t pmask = qz_mover_pawn_moves_ctx(c, s.meta);
    uint64_t hc = 0, vc = 0;
    const bool walls = qz_mover_walls(s.meta) > 0;
    if (walls) { hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V); }
    bool prepared = false;
    QzSweep w;
    for (uint32_t j = 0;; j++) {
        const int npawn = qz_popc32(pmask), nh = qz_popc64(hc), nv = qz_popc64(vc);
        const uint32_t M = (uint32_t)(npawn + nh + nv);
        if (M == 0) return -1;
        const uint32_t word = j == 0 ? qz_rng_first_word(rng, ply) : qz_philox(rng.seed, rng.rid, ply, j).x;
        int k = (int)qz_mulhi32(word, M);
        if (k < npawn) return qz_nth_bit64((uint64_t)pmask, k);
        k -= npawn;
        if (!prepared) { w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta)); prepared = true; }
        if (k < nh) {
            const int ix = qz_nth_bit64(hc, k);
            if (qz_wall_keeps_paths(w, ix, false)) return 12 + ix;
            hc &= ~(1ull << ix);
        } else {
            const int ix = qz_nth_bit64(vc, k - nh);
            if (qz_wall_keeps_paths(w, ix, true)) return 76 + ix;
            vc &= ~(1ull << ix);
        }
        if (j >= max_rejects) return -2;
    }
}

QZ_HD int qz_sample_action(const QzState &s, QzRng &rng, uint32_t ply) {
    return qz_sample_action_capped(s, rng, ply, 0xFFFFFFFFu);
}

// The same draw sequence as qz_sample_action when the legal walls (hl, vl) are already known (from a full
// sweep): rejected candidates are struck out by table lookup instead of by flood fills.  Identical result.
QZ_HD int qz_sample_action_known(const QzState &s, QzRng &rng, uint32_t ply, uint32_t pmask, uint64_t hl, uint64_t vl) {
    uint64_t hc = 0, vc = 0;
    if (qz_mover_walls(s.meta) > 0) { hc = qz_hcand(s.H, s.V); vc = qz_vcand(s.H, s.V); }
    for (uint32_t j = 0;; j++) {
        const int npawn = qz_popc32(pmask), nh = qz_popc64(hc), nv = qz_popc64(vc);
        const uint32_t M = (uint32_t)(npawn + nh + nv);
        if (M == 0) return -1;
        const uint32_t word = j == 0 ? qz_rng_first_word(rng, ply) : qz_philox(rng.seed, rng.rid, ply, j).x;
        int k = (int)qz_mulhi32(word, M);
        if (k < npawn) return qz_nth_bit64((uint64_t)pmask, k);
        k -= npawn;
        if (k < nh) {
            const int ix = qz_nth_bit64(hc, k);
            if ((hl >> ix) & 1ull) return 12 + ix;
            hc &= ~(1ull << ix);
        } else {
            const int ix = qz_nth_bit64(vc, k - nh);
            if ((vl >> ix) & 1ull) return 76 + ix;
            vc &= ~(1ull << ix);
        }
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
