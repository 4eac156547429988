// qz_rules.cuh -- Quoridor rules on bitboards, shared by every kernel in this directory.
//
// Everything here is `__host__ __device__` so the exact code the sm_100a kernels run can also be
// compiled for the host by tests/host_harness (a TEST build, never shipped or loaded by the product)
// and diffed against the oracle in a container without a GPU.
//
// Data model (DESIGN.md "Data layout"):
//   QzState  24 B/game: u64 H (horizontal walls, bit ix = r*8+c), u64 V (vertical walls), u64 meta
//            meta: byte0 P1 tile (int8, may be 81..89 after an off-board winning jump), byte1 P2 tile
//            (int8, may be -9..-1), byte2/3 walls left P1/P2, byte4 mover (1|2), byte5 flags, bytes6-7 ply
//   BB       81-bit tile set, tile t = r*9+c at bit t, as 3 x u32 (bits 81..95 always zero)
//   legal mask 140 bits as 3 x u64: bit a = action a (0..11 pawn, 12..75 H wall ix a-12, 76..139 V wall ix a-76)
//
// Reference semantics restated (file:line are into the reference repository):
//   corner values        quoridor.py:356-418  (incl. the row-0 NE/NW aliasing at :388,:392)
//   pawn moves           quoridor.py:272-353
//   wall prechecks       quoridor.py:432-461
//   path check           quoridor.py:463-528  (BFS -> bit-parallel flood fill; same reachability)
//   step / winner        quoridor.py:159-202, :217-269
//   state planes         quoridor.py:58-131
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QZ_HD __host__ __device__ __forceinline__
#define QZ_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define QZ_HD static inline
#define QZ_HD_NOINLINE static
#endif

struct QzState {
    uint64_t H, V, meta;
};

// ---- meta accessors -------------------------------------------------------------------------
#define QZ_FLAG_DONE 0x01u
#define QZ_FLAG_WINNER_SHIFT 1      // bits 1-2: winner (0 none, 1, 2)
#define QZ_FLAG_STALEMATE 0x08u     // mover has no legal action (reference: actions()==[] then crashes)
#define QZ_FLAG_TRUNCATED 0x10u     // ply cap hit (engine-side cap; reference loops forever)
#define QZ_FLAG_ILLEGAL 0x20u       // safe-mode step rejected the action (quoridor.py:167-169)
#define QZ_FLAG_PENDING 0x40u       // rollout scratch only: parked for the stuck-rollout kernel (never in user states)

QZ_HD int qz_p1(uint64_t m) { return (int)(int8_t)(m & 0xFF); }
QZ_HD int qz_p2(uint64_t m) { return (int)(int8_t)((m >> 8) & 0xFF); }
QZ_HD int qz_w1(uint64_t m) { return (int)((m >> 16) & 0xFF); }
QZ_HD int qz_w2(uint64_t m) { return (int)((m >> 24) & 0xFF); }
QZ_HD int qz_cur(uint64_t m) { return (int)((m >> 32) & 0xFF); }
QZ_HD unsigned qz_flags(uint64_t m) { return (unsigned)((m >> 40) & 0xFF); }
QZ_HD unsigned qz_ply(uint64_t m) { return (unsigned)((m >> 48) & 0xFFFF); }
QZ_HD bool qz_done(uint64_t m) { return (m >> 40) & QZ_FLAG_DONE; }
QZ_HD int qz_winner(uint64_t m) { return (int)((m >> (40 + QZ_FLAG_WINNER_SHIFT)) & 3); }

QZ_HD uint64_t qz_pack_meta(int p1, int p2, int w1, int w2, int cur, unsigned flags, unsigned ply) {
    return (uint64_t)(uint8_t)p1 | ((uint64_t)(uint8_t)p2 << 8) | ((uint64_t)(w1 & 0xFF) << 16) |
           ((uint64_t)(w2 & 0xFF) << 24) | ((uint64_t)(cur & 0xFF) << 32) | ((uint64_t)(flags & 0xFF) << 40) |
           ((uint64_t)(ply & 0xFFFF) << 48);
}

// quoridor.py:26-56
QZ_HD QzState qz_initial_state() {
    QzState s;
    s.H = 0; s.V = 0;
    s.meta = qz_pack_meta(4, 76, 10, 10, 1, 0, 0);
    return s;
}

// ---- small portable intrinsics ------------------------------------------------------------------
QZ_HD int qz_popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
QZ_HD int qz_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
QZ_HD uint32_t qz_fshl(uint32_t lo, uint32_t hi, int k) {   // high word of ((hi:lo) << k), 0<k<32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, k);
#else
    return (hi << k) | (lo >> (32 - k));
#endif
}
QZ_HD uint32_t qz_fshr(uint32_t lo, uint32_t hi, int k) {   // low word of ((hi:lo) >> k), 0<k<32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, k);
#else
    return (lo >> k) | (hi << (32 - k));
#endif
}
// index of the k-th (0-based) set bit of m; m must have more than k bits set
QZ_HD int qz_nth_bit64(uint64_t m, int k) {
#if defined(__CUDA_ARCH__)
    // binary search on popcounts of the low half / byte / nibble / pair (the __fns intrinsic is ~45 instructions and
    // sat on the critical path of every draw of the rollout kernels)
    uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    int c = __popc(lo);
    uint32_t w = lo; int r = 0;
    if (k >= c) { k -= c; w = hi; r = 32; }
    c = __popc(w & 0xFFFFu);
    if (k >= c) { k -= c; w >>= 16; r += 16; }
    c = __popc(w & 0xFFu);
    if (k >= c) { k -= c; w >>= 8; r += 8; }
    c = __popc(w & 0xFu);
    if (k >= c) { k -= c; w >>= 4; r += 4; }
    c = __popc(w & 0x3u);
    if (k >= c) { k -= c; w >>= 2; r += 2; }
    if (k >= (int)(w & 1u)) r += 1;
    return r;
#else
    for (int i = 0; i < k; i++) m &= m - 1;
    return __builtin_ctzll(m);
#endif
}

// ---- 81-bit boards --------------------------------------------------------------------------------
struct BB {
    uint32_t w0, w1, w2;
};
QZ_HD BB bb_zero() { BB b; b.w0 = b.w1 = b.w2 = 0; return b; }
QZ_HD BB bb_bit(int t) {   // 0 <= t <= 80
    BB b;
    b.w0 = t < 32 ? 1u << t : 0u;
    b.w1 = (t >= 32 && t < 64) ? 1u << (t - 32) : 0u;
    b.w2 = t >= 64 ? 1u << (t - 64) : 0u;
    return b;
}
QZ_HD bool bb_test(BB b, int t) {
    uint32_t w = t < 32 ? b.w0 : (t < 64 ? b.w1 : b.w2);
    return (w >> (t & 31)) & 1u;
}
QZ_HD BB bb_and(BB a, BB b) { BB r; r.w0 = a.w0 & b.w0; r.w1 = a.w1 & b.w1; r.w2 = a.w2 & b.w2; return r; }
QZ_HD BB bb_or(BB a, BB b) { BB r; r.w0 = a.w0 | b.w0; r.w1 = a.w1 | b.w1; r.w2 = a.w2 | b.w2; return r; }
QZ_HD BB bb_andn(BB a, BB b) { BB r; r.w0 = a.w0 & ~b.w0; r.w1 = a.w1 & ~b.w1; r.w2 = a.w2 & ~b.w2; return r; }
QZ_HD bool bb_eq(BB a, BB b) { return ((a.w0 ^ b.w0) | (a.w1 ^ b.w1) | (a.w2 ^ b.w2)) == 0; }
QZ_HD bool bb_any(BB a) { return (a.w0 | a.w1 | a.w2) != 0; }
QZ_HD BB bb_shl(BB a, int k) {   // toward higher tiles; caller guarantees nothing leaves bit 80
    BB r;
    r.w2 = qz_fshl(a.w1, a.w2, k);
    r.w1 = qz_fshl(a.w0, a.w1, k);
    r.w0 = a.w0 << k;
    return r;
}
QZ_HD BB bb_shr(BB a, int k) {
    BB r;
    r.w0 = qz_fshr(a.w0, a.w1, k);
    r.w1 = qz_fshr(a.w1, a.w2, k);
    r.w2 = a.w2 >> k;
    return r;
}

#define QZ_ROW0_W0 0x1FFu          // tiles 0..8
#define QZ_ROW8_W2 0x1FF00u        // tiles 72..80 = bits 8..16 of w2
#define QZ_BOARD_W2 0x1FFFFu       // tiles 64..80

// 8x8 intersection mask (bit r*8+c) -> 9-stride board (bit r*9+c); column 8 and row 8 stay empty
QZ_HD BB bb_spread8(uint64_t x) {
    // rows 4..7 up by 4, then rows {2,3,6,7} by 2, then odd rows by 1 (each row r moves up by r)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    // after step 1: bits 0..31 = rows 0-3, bits 36..67 = rows 4-7  -> 3 words
    uint32_t a0 = lo, a1 = hi << 4, a2 = hi >> 28;
    // step 2: rows {2,3} are bits 16..31 of a0; rows {6,7} are bits 52..67 (a1 bits 20..31, a2 bits 0..3)
    uint32_t m0 = a0 & 0xFFFF0000u, m1 = a1 & 0xFFF00000u, m2 = a2;
    uint32_t k0 = a0 & 0x0000FFFFu, k1 = a1 & 0x000FFFFFu;
    uint32_t b0 = k0 | (m0 << 2);
    uint32_t b1 = k1 | (m0 >> 30) | (m1 << 2);
    uint32_t b2 = (m1 >> 30) | (m2 << 2);
    // now row r sits at bit 8r + 2*(r>>1) + 4*(r>>2)... explicitly: r0@0 r1@8 r2@18 r3@26 r4@36 r5@44 r6@54 r7@62
    // step 3: odd rows up by 1: r1@8..15 (b0), r3@26..33 (b0 bits 26-31, b1 bits 0-1), r5@44..51 (b1 12..19),
    //         r7@62..69 (b1 bits 30-31, b2 bits 0-5)
    uint32_t o0 = b0 & 0xFC00FF00u, o1 = b1 & 0xC00FF003u, o2 = b2 & 0x3Fu;
    uint32_t e0 = b0 & ~0xFC00FF00u, e1 = b1 & ~0xC00FF003u, e2 = b2 & ~0x3Fu;
    BB r;
    r.w0 = e0 | (o0 << 1);
    r.w1 = e1 | (o0 >> 31) | (o1 << 1);
    r.w2 = e2 | (o1 >> 31) | (o2 << 1);
    return r;
}

// ---- direction masks ("a pawn on tile t may make the plain move X"), opponent ignored ----------------
// SURVEY.md Appendix A table, derived from quoridor.py:287-293 + :356-418; note the row-0 quirk
// (N and E from row 0 test the wall WEST of the tile's NE corner).
struct QzDirs {
    BB n, s, e, w;
};

#define QZ_COL0_W0 0x08040201u     // tiles 0,9,18,27
#define QZ_COL0_W1 0x80402010u     // tiles 36,45,54,63
#define QZ_COL0_W2 0x00000100u     // tile 72
#define QZ_COL8_W0 0x04020100u     // tiles 8,17,26
#define QZ_COL8_W1 0x40201008u     // tiles 35,44,53,62
#define QZ_COL8_W2 0x00010080u     // tiles 71,80

QZ_HD QzDirs qz_dirs(uint64_t H, uint64_t V) {
    BB h9 = bb_spread8(H), v9 = bb_spread8(V);
    BB hh = bb_or(h9, bb_shl(h9, 1));            // h(r,c) | h(r,c-1) at tile (r,c)
    BB vv = bb_or(v9, bb_shl(v9, 9));            // v(r,c) | v(r-1,c) at tile (r,c)
    // blocked-N: regular rows use hh; row 0 uses h(0,c-1) (c>0) or h(0,0) (c==0); row 8 never moves N
    uint32_t r0h = h9.w0 & 0xFFu;                // row-0 H walls, columns 0..7
    uint32_t r0v = v9.w0 & 0xFFu;
    BB bn = hh;
    bn.w0 = (bn.w0 & ~QZ_ROW0_W0) | ((r0h << 1) | (r0h & 1u));
    bn.w2 |= QZ_ROW8_W2;
    // blocked-S: h(r-1,c) | h(r-1,c-1); row 0 never moves S
    BB bs = bb_shl(hh, 9);
    bs.w0 |= QZ_ROW0_W0;
    // blocked-E: v(r,c) | v(r-1,c); row 0: v(0,c-1) (c>0) or v(0,0); column 8 never
    BB be = vv;
    be.w0 = (be.w0 & ~QZ_ROW0_W0) | ((r0v << 1) | (r0v & 1u));
    be.w0 |= QZ_COL8_W0; be.w1 |= QZ_COL8_W1; be.w2 |= QZ_COL8_W2;
    // blocked-W: v(r,c-1) | v(r-1,c-1); column 0 never
    BB bw = bb_shl(vv, 1);
    bw.w0 |= QZ_COL0_W0; bw.w1 |= QZ_COL0_W1; bw.w2 |= QZ_COL0_W2;
    QzDirs d;
    d.n.w0 = ~bn.w0; d.n.w1 = ~bn.w1; d.n.w2 = ~bn.w2 & QZ_BOARD_W2;
    d.s.w0 = ~bs.w0; d.s.w1 = ~bs.w1; d.s.w2 = ~bs.w2 & QZ_BOARD_W2;
    d.e.w0 = ~be.w0; d.e.w1 = ~be.w1; d.e.w2 = ~be.w2 & QZ_BOARD_W2;
    d.w.w0 = ~bw.w0; d.w.w1 = ~bw.w1; d.w.w2 = ~bw.w2 & QZ_BOARD_W2;
    return d;
}

// Tiles whose N / S (horizontal wall) or E / W (vertical wall) move a single new wall at
// intersection ix removes.  The blocked sets are unions over walls, so placing a candidate is
// `dirs.x &= ~delta` -- no re-spread per candidate.
QZ_HD void qz_dirs_place_h(QzDirs &d, int ix) {
    int r = ix >> 3, c = ix & 7, t = r * 9 + c;
    BB dn;
    if (r == 0) { dn = bb_bit(t + 1); if (c == 0) dn.w0 |= 1u; }
    else dn = bb_or(bb_bit(t), bb_bit(t + 1));
    BB ds = bb_or(bb_bit(t + 9), bb_bit(t + 10));
    d.n = bb_andn(d.n, dn);
    d.s = bb_andn(d.s, ds);
}
QZ_HD void qz_dirs_place_v(QzDirs &d, int ix) {
    int r = ix >> 3, c = ix & 7, t = r * 9 + c;
    BB de;
    if (r == 0) { de = bb_or(bb_bit(t + 1), bb_bit(t + 9)); if (c == 0) de.w0 |= 1u; }
    else de = bb_or(bb_bit(t), bb_bit(t + 9));
    BB dw = bb_or(bb_bit(t + 1), bb_bit(t + 10));
    d.e = bb_andn(d.e, de);
    d.w = bb_andn(d.w, dw);
}

// ---- corner values (quoridor.py:356-418): 0 none, 1 horizontal, 2 vertical -----------------------------
#define QZ_CH 1
#define QZ_CV 2
QZ_HD int qz_wall_at(uint64_t H, uint64_t V, int r, int c) {
    int i = r * 8 + c;
    return (int)((H >> i) & 1u) | ((int)((V >> i) & 1u) << 1);
}
struct QzCorners {
    int nw, ne, se, sw;
};
QZ_HD QzCorners qz_corners(uint64_t H, uint64_t V, int t) {   // 0 <= t <= 80
    int r = t / 9, c = t - 9 * r;
    QzCorners k;
    if (r == 8) {
        k.ne = QZ_CH;
        k.nw = (c == 0) ? QZ_CV : QZ_CH;
        k.se = (c == 8) ? QZ_CV : qz_wall_at(H, V, 7, c);
        k.sw = (c == 0) ? QZ_CV : qz_wall_at(H, V, 7, c - 1);
    } else if (r == 0) {
        k.sw = QZ_CH;
        k.se = (c == 8) ? QZ_CV : QZ_CH;
        if (c == 0) { k.nw = QZ_CV; k.ne = qz_wall_at(H, V, 0, 0); }
        else { k.nw = k.ne = qz_wall_at(H, V, 0, c - 1); }          // the :388/:392 aliasing
    } else {
        k.nw = (c == 0) ? QZ_CV : qz_wall_at(H, V, r, c - 1);
        k.sw = (c == 0) ? QZ_CV : qz_wall_at(H, V, r - 1, c - 1);
        k.ne = (c == 8) ? QZ_CV : qz_wall_at(H, V, r, c);
        k.se = (c == 8) ? QZ_CV : qz_wall_at(H, V, r - 1, c);
    }
    return k;
}

// quoridor.py:272-353 -> 12-bit mask, bit id = action id (0 N,1 S,2 E,3 W,4 NN,5 SS,6 EE,7 WW,8 NE,9 NW,10 SE,11 SW)
// L must be on the board; O is the opponent's tile (on the board for every non-terminal state).
QZ_HD uint32_t qz_pawn_moves(uint64_t H, uint64_t V, int L, int O, int player) {
    QzCorners I = qz_corners(H, V, L);
    int row = L / 9;
    bool on = (L == O - 9), os = (L == O + 9), oe = (L == O - 1), ow = (L == O + 1);
    uint32_t m = 0;
    bool n = I.nw != QZ_CH && I.ne != QZ_CH && !on;
    bool s = I.sw != QZ_CH && I.se != QZ_CH && !os;
    bool e = I.ne != QZ_CV && I.se != QZ_CV && !oe;
    bool w = I.nw != QZ_CV && I.sw != QZ_CV && !ow;
    if (n || (player == 1 && row == 8)) m |= 1u << 0;
    if (s || (player == 2 && row == 0)) m |= 1u << 1;
    if (e) m |= 1u << 2;
    if (w) m |= 1u << 3;
    if (!(on || os || oe || ow) || O < 0 || O > 80) return m;
    QzCorners P = qz_corners(H, V, O);
    if (on && I.ne != QZ_CH && I.nw != QZ_CH) {
        if ((P.nw != QZ_CH && P.ne != QZ_CH) || (row == 7 && player == 1)) m |= 1u << 4;
        if (P.ne != QZ_CV && I.ne != QZ_CV) m |= 1u << 8;
        if (P.nw != QZ_CV && I.nw != QZ_CV) m |= 1u << 9;
    } else if (os && I.se != QZ_CH && I.sw != QZ_CH) {
        if ((P.sw != QZ_CH && P.se != QZ_CH) || (row == 1 && player == 2)) m |= 1u << 5;
        if (P.se != QZ_CV && I.se != QZ_CV) m |= 1u << 10;
        if (P.sw != QZ_CV && I.sw != QZ_CV) m |= 1u << 11;
    } else if (oe && I.se != QZ_CV && I.ne != QZ_CV) {
        if (P.se != QZ_CV && P.ne != QZ_CV) m |= 1u << 6;
        if (P.ne != QZ_CH) m |= 1u << 8;
        if (P.se != QZ_CH) m |= 1u << 10;
    } else if (ow && I.sw != QZ_CV && I.nw != QZ_CV) {
        if (P.nw != QZ_CV && P.sw != QZ_CV) m |= 1u << 7;
        if (P.nw != QZ_CH) m |= 1u << 9;
        if (P.sw != QZ_CH) m |= 1u << 11;
    }
    return m;
}

// tile offset of each pawn action (quoridor.py:217-243)
QZ_HD int qz_delta(int a) {
    // {9,-9,1,-1,18,-18,2,-2,10,8,-8,-10} packed as bytes to stay in registers
    const uint64_t lo = 0xFE02EE12FF01F709ull;   // a = 0..7 : 9,-9,1,-1,18,-18,2,-2
    const uint32_t hi = 0xF6F8080Au;             // a = 8..11: 10,8,-8,-10
    return a < 8 ? (int)(int8_t)(lo >> (8 * a)) : (int)(int8_t)(hi >> (8 * (a - 8)));
}

// ---- branch-free pawn-move generation from per-corner masks ---------------------------------------------------
// The scalar qz_corners/qz_pawn_moves above mirror the reference line by line but branch on the tile, which
// serialises a warp whose lanes stand on different tiles.  The device paths therefore use twelve 81-bit masks
// built once per wall configuration: for every tile, "corner k IS a horizontal / vertical wall" (synthetic
// border walls and the row-0 NE/NW aliasing included, quoridor.py:356-418) and the four plain-move masks
// derived from them.  A pawn-move query is then ~20 bit tests and no branch (quoridor.py:272-353).
struct QzPawnCtx {
    QzDirs d;
    BB neV, nwV, seV, swV;
    BB neH, nwH, seH, swH;
};

QZ_HD BB bb_make(uint32_t w0, uint32_t w1, uint32_t w2) { BB b; b.w0 = w0; b.w1 = w1; b.w2 = w2; return b; }
QZ_HD uint32_t bb_at(const BB &b, int t) {     // bit t as 0/1, 0 <= t <= 80
    const uint32_t w = t < 32 ? b.w0 : (t < 64 ? b.w1 : b.w2);
    return (w >> (t & 31)) & 1u;
}

QZ_HD QzPawnCtx qz_ctx_build(uint64_t H, uint64_t V) {
    const BB h9 = bb_spread8(H), v9 = bb_spread8(V);
    const uint32_t r0h = h9.w0 & 0xFFu, r0v = v9.w0 & 0xFFu;
    const uint32_t fixh = (r0h << 1) | (r0h & 1u), fixv = (r0v << 1) | (r0v & 1u);   // row 0: NE = NW = ix(0,c-1)
    QzPawnCtx c;
    c.neV = bb_or(v9, bb_make(0x04020000u, QZ_COL8_W1, 0x80u));        // column 8, rows 1..7
    c.neV.w0 = (c.neV.w0 & ~QZ_ROW0_W0) | fixv;
    c.neH = h9;
    c.neH.w2 |= QZ_ROW8_W2;                                             // row 8: NE = H
    c.neH.w0 = (c.neH.w0 & ~QZ_ROW0_W0) | fixh;
    c.nwV = bb_or(bb_shl(v9, 1), bb_make(QZ_COL0_W0, QZ_COL0_W1, QZ_COL0_W2));
    c.nwH = bb_shl(h9, 1);
    c.nwH.w2 |= 0x1FE00u;                                               // row 8, columns 1..8
    c.seV = bb_or(bb_shl(v9, 9), bb_make(QZ_COL8_W0, QZ_COL8_W1, QZ_COL8_W2));
    c.seH = bb_shl(h9, 9);
    c.seH.w0 |= 0xFFu;                                                  // row 0, columns 0..7
    c.swV = bb_or(bb_shl(v9, 10), bb_make(0x08040200u, QZ_COL0_W1, QZ_COL0_W2));   // column 0, rows 1..8
    c.swH = bb_shl(h9, 10);
    c.swH.w0 |= QZ_ROW0_W0;                                             // row 0
    c.d.n.w0 = ~(c.nwH.w0 | c.neH.w0); c.d.n.w1 = ~(c.nwH.w1 | c.neH.w1); c.d.n.w2 = ~(c.nwH.w2 | c.neH.w2) & QZ_BOARD_W2;
    c.d.s.w0 = ~(c.swH.w0 | c.seH.w0); c.d.s.w1 = ~(c.swH.w1 | c.seH.w1); c.d.s.w2 = ~(c.swH.w2 | c.seH.w2) & QZ_BOARD_W2;
    c.d.e.w0 = ~(c.neV.w0 | c.seV.w0); c.d.e.w1 = ~(c.neV.w1 | c.seV.w1); c.d.e.w2 = ~(c.neV.w2 | c.seV.w2) & QZ_BOARD_W2;
    c.d.w.w0 = ~(c.nwV.w0 | c.swV.w0); c.d.w.w1 = ~(c.nwV.w1 | c.swV.w1); c.d.w.w2 = ~(c.nwV.w2 | c.swV.w2) & QZ_BOARD_W2;
    return c;
}

// the twelve masks as 36 words (for parking a ctx in shared memory while the walls do not change)
#define QZ_CTX_WORDS 36
#define QZ_CTX_EACH(F) F(d.n, 0) F(d.s, 1) F(d.e, 2) F(d.w, 3) F(neV, 4) F(nwV, 5) F(seV, 6) F(swV, 7) F(neH, 8) F(nwH, 9) F(seH, 10) F(swH, 11)
QZ_HD void qz_ctx_store(const QzPawnCtx &c, uint32_t *w) {
#define QZ_CTX_ST(m, i) w[3 * i] = c.m.w0; w[3 * i + 1] = c.m.w1; w[3 * i + 2] = c.m.w2;
    QZ_CTX_EACH(QZ_CTX_ST)
#undef QZ_CTX_ST
}
QZ_HD QzPawnCtx qz_ctx_load(const uint32_t *w) {
    QzPawnCtx c;
#define QZ_CTX_LD(m, i) c.m.w0 = w[3 * i]; c.m.w1 = w[3 * i + 1]; c.m.w2 = w[3 * i + 2];
    QZ_CTX_EACH(QZ_CTX_LD)
#undef QZ_CTX_LD
    return c;
}

// quoridor.py:272-353 without a branch.  L on the board; any O (an off-board O is never adjacent).
QZ_HD uint32_t qz_pawn_moves_ctx(const QzPawnCtx &c, int L, int O, int player) {
    const uint32_t ovalid = (unsigned)O <= 80u ? 1u : 0u;
    const int Oc = ovalid ? O : 0;
    const uint32_t p1 = player == 1 ? 1u : 0u, p2 = p1 ^ 1u;
    const uint32_t on = (L == O - 9 ? 1u : 0u) & ovalid, os = (L == O + 9 ? 1u : 0u) & ovalid;
    const uint32_t oe = (L == O - 1 ? 1u : 0u) & ovalid, ow = (L == O + 1 ? 1u : 0u) & ovalid;
    const uint32_t nL = bb_at(c.d.n, L), sL = bb_at(c.d.s, L), eL = bb_at(c.d.e, L), wL = bb_at(c.d.w, L);
    uint32_t m = ((nL & (on ^ 1u)) | (p1 & (L >= 72 ? 1u : 0u)))
               | (((sL & (os ^ 1u)) | (p2 & (L < 9 ? 1u : 0u))) << 1)
               | ((eL & (oe ^ 1u)) << 2) | ((wL & (ow ^ 1u)) << 3);
    const uint32_t gN = on & nL, gS = os & sL, gE = oe & eL, gW = ow & wL;      // wall-free contact with the opponent
    const uint32_t nn = gN & (bb_at(c.d.n, Oc) | (p1 & (L >= 63 ? 1u : 0u)));   // :305-308 (row 7, P1: off-board win)
    const uint32_t ss = gS & (bb_at(c.d.s, Oc) | (p2 & (L < 18 ? 1u : 0u)));    // :319-321 (row 1, P2)
    const uint32_t ee = gE & bb_at(c.d.e, Oc);
    const uint32_t ww = gW & bb_at(c.d.w, Oc);
    const uint32_t ne = (gN & (bb_at(c.neV, Oc) ^ 1u) & (bb_at(c.neV, L) ^ 1u)) | (gE & (bb_at(c.neH, Oc) ^ 1u));
    const uint32_t nw = (gN & (bb_at(c.nwV, Oc) ^ 1u) & (bb_at(c.nwV, L) ^ 1u)) | (gW & (bb_at(c.nwH, Oc) ^ 1u));
    const uint32_t se = (gS & (bb_at(c.seV, Oc) ^ 1u) & (bb_at(c.seV, L) ^ 1u)) | (gE & (bb_at(c.seH, Oc) ^ 1u));
    const uint32_t sw = (gS & (bb_at(c.swV, Oc) ^ 1u) & (bb_at(c.swV, L) ^ 1u)) | (gW & (bb_at(c.swH, Oc) ^ 1u));
    m |= (nn << 4) | (ss << 5) | (ee << 6) | (ww << 7) | (ne << 8) | (nw << 9) | (se << 10) | (sw << 11);
    return m;
}

// ---- per-tile info bytes (pawn-only play) --------------------------------------------------------------------
// When no wall can be placed any more the walls are constant, so everything quoridor.py:272-353 reads about a tile
// is eight bits: bit 0..3 = plain move N,S,E,W open (d.n/s/e/w), bit 4..7 = corner NE,NW,SE,SW is a VERTICAL wall.
// Byte t of an 84-byte table (21 words) belongs to tile t.  The table is the transpose of eight ctx masks, built
// four tiles at a time: nibble * 0x00204081 puts bit j of the nibble at bit 8j (terms at offsets 0,7,14,21 never
// overlap), shifted by k for mask k.
#define QZ_TILE_TABLE_WORDS 21
QZ_HD uint32_t qz_tile_group(uint32_t n, uint32_t s, uint32_t e, uint32_t w, uint32_t a, uint32_t b, uint32_t c,
                             uint32_t d, int sh) {
#define QZ_SPREAD4(x, k) (((((x) >> sh) & 0xFu) * (0x00204081u << (k))) & (0x01010101u << (k)))
    return QZ_SPREAD4(n, 0) | QZ_SPREAD4(s, 1) | QZ_SPREAD4(e, 2) | QZ_SPREAD4(w, 3) | QZ_SPREAD4(a, 4) |
           QZ_SPREAD4(b, 5) | QZ_SPREAD4(c, 6) | QZ_SPREAD4(d, 7);
#undef QZ_SPREAD4
}
// tbl may be any word-addressable memory with element stride `stride` (shared memory on the device)
QZ_HD void qz_tile_table(const QzPawnCtx &c, uint32_t *tbl, int stride) {
#pragma unroll
    for (int g = 0; g < 8; g++)
        tbl[g * stride] = qz_tile_group(c.d.n.w0, c.d.s.w0, c.d.e.w0, c.d.w.w0, c.neV.w0, c.nwV.w0, c.seV.w0, c.swV.w0, 4 * g);
#pragma unroll
    for (int g = 0; g < 8; g++)
        tbl[(8 + g) * stride] = qz_tile_group(c.d.n.w1, c.d.s.w1, c.d.e.w1, c.d.w.w1, c.neV.w1, c.nwV.w1, c.seV.w1, c.swV.w1, 4 * g);
#pragma unroll
    for (int g = 0; g < 5; g++)
        tbl[(16 + g) * stride] = qz_tile_group(c.d.n.w2, c.d.s.w2, c.d.e.w2, c.d.w.w2, c.neV.w2, c.nwV.w2, c.seV.w2, c.swV.w2, 4 * g);
}

// Which plain move of the mover on L runs into the opponent on O with no wall between (one-hot N,S,E,W or 0):
// the only case in which jump moves exist (quoridor.py:303,317,331,343).  Adjacency is by raw tile offset (:278-281).
QZ_HD uint32_t qz_pawn_contact(uint32_t iL, int L, int O) {
    const int d = O - L;
    const uint32_t adj = (d == 9 ? 1u : 0u) | (d == -9 ? 2u : 0u) | (d == 1 ? 4u : 0u) | (d == -1 ? 8u : 0u);
    return iL & adj;
}
// quoridor.py:272-353 from info bytes: iL / iO of the mover's and the opponent's tile, hO = "corner is a
// HORIZONTAL wall" bits of the opponent's tile (bit 0 NE, 1 NW, 2 SE, 3 SW; only read on east / west contact).
// Same result as qz_pawn_moves_ctx for L, O on the board.
QZ_HD uint32_t qz_pawn_moves_info(uint32_t iL, uint32_t iO, uint32_t hO, int L, int O, int player) {
    const int d = O - L;
    const uint32_t adj = (d == 9 ? 1u : 0u) | (d == -9 ? 2u : 0u) | (d == 1 ? 4u : 0u) | (d == -1 ? 8u : 0u);
    uint32_t m = iL & 0xFu & ~adj;
    if (player == 1 && L >= 72) m |= 1u;                       // :295
    if (player == 2 && L < 9) m |= 2u;                         // :296-297
    const uint32_t g = iL & adj;
    if (g) {
        const uint32_t nV = ~(iL | iO);                        // bit 4..7: corner is vertical on neither tile
        if (g & 1u) {
            const uint32_t nn = (iO & 1u) | ((player == 1 && L >= 63) ? 1u : 0u);              // :305-308
            m |= (nn << 4) | (((nV >> 4) & 1u) << 8) | (((nV >> 5) & 1u) << 9);                 // :310-314
        } else if (g & 2u) {
            const uint32_t ss = ((iO >> 1) & 1u) | ((player == 2 && L < 18) ? 1u : 0u);        // :319-321
            m |= (ss << 5) | (((nV >> 6) & 1u) << 10) | (((nV >> 7) & 1u) << 11);               // :323-327
        } else if (g & 4u) {
            m |= (((iO >> 2) & 1u) << 6) | (((~hO) & 1u) << 8) | (((~hO >> 2) & 1u) << 10);     // :333-339
        } else {
            m |= (((iO >> 3) & 1u) << 7) | (((~hO >> 1) & 1u) << 9) | (((~hO >> 3) & 1u) << 11); // :345-351
        }
    }
    return m;
}

// ---- jump edges seen by the path check ------------------------------------------------------------------
// In _bfs_to_goal (quoridor.py:479-528) the opponent pawn is a fixed obstacle: plain moves into its
// tile are dropped and up to three jump edges leave each of the four neighbouring tiles.  Encoded as one
// u32: byte i = actions 4..11 available from source tile O + QZ_SRC[i]  (i: 0 = O-9, 1 = O+9, 2 = O-1, 3 = O+1).
QZ_HD int qz_jump_src(int O, int i) { return O + (i == 0 ? -9 : (i == 1 ? 9 : (i == 2 ? -1 : 1))); }

QZ_HD uint32_t qz_jump_set_ctx(const QzPawnCtx &c, int O, int player) {
    uint32_t js = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int src = qz_jump_src(O, i);
        const uint32_t valid = (unsigned)src <= 80u ? 1u : 0u;
        const uint32_t m = qz_pawn_moves_ctx(c, valid ? src : 0, O, player) >> 4;
        js |= (valid ? (m & 0xFFu) : 0u) << (8 * i);
    }
    return js;
}

// Jump edges for one search under an arbitrary wall configuration.  Out of line: only needed when the
// plain-move closure fails to reach the goal (a blocked or nearly blocked candidate), see qz_reaches_goal.
QZ_HD_NOINLINE uint32_t qz_jump_set_rebuilt(uint64_t H, uint64_t V, int O, int player) {
    const QzPawnCtx c = qz_ctx_build(H, V);
    return qz_jump_set_ctx(c, O, player);
}

// ---- the path check: can `player` standing on `start` still reach its goal row? -------------------------
// Same reachability as quoridor.py:479-528: closure of {plain moves not entering O} U {jump edges};
// touching the goal row ends the search; off-board landings (NN from row 7 / SS from row 1) are recorded by
// the reference but never expanded and never equal the goal row, so they are dropped here.
// Jump edges only ADD reachability, so the search first closes over plain moves alone -- which settles almost
// every legal candidate -- and builds the jump set (walls H, V = the configuration being tested) only if
// that closure did not touch the goal row.
// One closure step of the plain moves (no jumps): reach | moves(reach) & keep.
QZ_HD BB qz_plain_step(const QzDirs &d, const BB &reach, const BB &keep) {
    const BB a = bb_shl(bb_and(reach, d.n), 9), b = bb_shr(bb_and(reach, d.s), 9);
    const BB c = bb_shl(bb_and(reach, d.e), 1), e = bb_shr(bb_and(reach, d.w), 1);
    BB nxt;
    nxt.w0 = reach.w0 | ((a.w0 | b.w0 | c.w0 | e.w0) & keep.w0);
    nxt.w1 = reach.w1 | ((a.w1 | b.w1 | c.w1 | e.w1) & keep.w1);
    nxt.w2 = reach.w2 | ((a.w2 | b.w2 | c.w2 | e.w2) & keep.w2);
    return nxt;
}
QZ_HD uint32_t qz_goal_hit(const BB &r, int player) { return player == 1 ? (r.w2 & QZ_ROW8_W2) : (r.w0 & QZ_ROW0_W0); }

// Continue a search whose plain-move closure `reach` (a fixpoint that misses the goal row) is known: apply the
// jump edges, re-close, repeat.  The jump set is built here, lazily (see qz_reaches_goal).
QZ_HD bool qz_reach_with_jumps(const QzDirs &d, BB reach, int O, int player, uint64_t H, uint64_t V) {
    // A jump edge leaves one of the four tiles next to the opponent (raw offsets, quoridor.py:278-281).  `reach` is closed
    // under plain moves and plain moves add nothing next to the opponent later, so if none of those tiles is in it no
    // jump can ever be taken: the answer is "no" without building the jump set (a corner-mask rebuild, ~600
    // instructions -- and in blocked positions nearly every check ends here).
    {
        BB src = bb_zero();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int t = qz_jump_src(O, i);
            if ((unsigned)t <= 80u) src = bb_or(src, bb_bit(t));
        }
        if (!bb_any(bb_and(reach, src))) return false;
    }
    BB keep = bb_bit(O);
    keep.w0 = ~keep.w0; keep.w1 = ~keep.w1; keep.w2 = ~keep.w2;
    const uint32_t jumps = qz_jump_set_rebuilt(H, V, O, player);
    if (jumps == 0) return false;
    for (;;) {
        BB add = bb_zero();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t byte = (jumps >> (8 * i)) & 0xFFu;
            if (byte == 0) continue;
            int src = qz_jump_src(O, i);
            if (!bb_test(reach, src)) continue;
            for (int b = 0; b < 8; b++) {
                if (!((byte >> b) & 1u)) continue;
                int land = src + qz_delta(4 + b);
                if (land >= 0 && land <= 80) add = bb_or(add, bb_bit(land));
            }
        }
        if (qz_goal_hit(add, player)) return true;
        BB nxt = bb_or(reach, add);
        if (bb_eq(nxt, reach)) return false;
        reach = nxt;
        for (;;) {
            nxt = qz_plain_step(d, reach, keep);
            if (qz_goal_hit(nxt, player)) return true;
            if (bb_eq(nxt, reach)) break;
            reach = nxt;
        }
    }
}

QZ_HD bool qz_reaches_goal(const QzDirs &d, int start, int O, int player, uint64_t H, uint64_t V) {
    BB reach = bb_bit(start);
    BB keep = bb_bit(O);
    keep.w0 = ~keep.w0; keep.w1 = ~keep.w1; keep.w2 = ~keep.w2;
    for (;;) {
        const BB nxt = qz_plain_step(d, reach, keep);
        if (qz_goal_hit(nxt, player)) return true;
        if (bb_eq(nxt, reach)) break;
        reach = nxt;
    }
    return qz_reach_with_jumps(d, reach, O, player, H, V);
}

// Both searches of _blocks_path (quoridor.py:474-475) advanced in lockstep by one thread: the two closures are
// independent dependency chains, so interleaving them costs little more time than one.  Returns true iff the wall
// configuration (d, H, V) leaves both players a path.
QZ_HD bool qz_both_reach_goal(const QzDirs &d, int p1, int p2, uint64_t H, uint64_t V) {
    BB r1 = bb_bit(p1), r2 = bb_bit(p2);
    BB k1 = bb_bit(p2), k2 = bb_bit(p1);
    k1.w0 = ~k1.w0; k1.w1 = ~k1.w1; k1.w2 = ~k1.w2;
    k2.w0 = ~k2.w0; k2.w1 = ~k2.w1; k2.w2 = ~k2.w2;
    bool hit1 = false, hit2 = false, fix1 = false, fix2 = false;
    while (!((hit1 | fix1) & (hit2 | fix2))) {
        const BB n1 = qz_plain_step(d, r1, k1), n2 = qz_plain_step(d, r2, k2);     // steps after a verdict are harmless
        hit1 = hit1 | (qz_goal_hit(n1, 1) != 0);
        hit2 = hit2 | (qz_goal_hit(n2, 2) != 0);
        fix1 = bb_eq(n1, r1);
        fix2 = bb_eq(n2, r2);
        r1 = n1; r2 = n2;
    }
    if (!hit1 && !qz_reach_with_jumps(d, r1, p2, 1, H, V)) return false;
    if (!hit2 && !qz_reach_with_jumps(d, r2, p1, 2, H, V)) return false;
    return true;
}

// ---- wall candidates ------------------------------------------------------------------------------------
// quoridor.py:432-461 without the path check: the empty intersections where an H (resp. V) wall does not
// overlap a collinear neighbour.
QZ_HD uint64_t qz_hcand(uint64_t H, uint64_t V) {
    const uint64_t COL0 = 0x0101010101010101ull, COL7 = 0x8080808080808080ull;
    return ~(H | V) & ~((H << 1) & ~COL0) & ~((H >> 1) & ~COL7);
}
QZ_HD uint64_t qz_vcand(uint64_t H, uint64_t V) { return ~(H | V) & ~(V << 8) & ~(V >> 8); }

// Everything one position's 128-candidate sweep shares (registers of every lane of the warp).
struct QzSweep {
    uint64_t H, V;
    QzDirs dirs;
    int p1, p2;
};

QZ_HD QzSweep qz_sweep_prepare_ctx(const QzPawnCtx &c, uint64_t H, uint64_t V, int p1, int p2) {
    QzSweep w;
    w.H = H; w.V = V; w.p1 = p1; w.p2 = p2;
    w.dirs = c.d;
    return w;
}
QZ_HD QzSweep qz_sweep_prepare(uint64_t H, uint64_t V, int p1, int p2) {
    QzSweep w;
    w.H = H; w.V = V; w.p1 = p1; w.p2 = p2;
    w.dirs = qz_dirs(H, V);
    return w;
}

// quoridor.py:463-477 (_blocks_path negated): true iff placing the wall leaves both players a path.
QZ_HD bool qz_wall_keeps_paths(const QzSweep &w, int ix, bool vertical) {
    QzDirs d = w.dirs;
    uint64_t H = w.H, V = w.V;
    uint64_t bit = 1ull << ix;
    if (vertical) { qz_dirs_place_v(d, ix); V |= bit; }
    else { qz_dirs_place_h(d, ix); H |= bit; }
    return qz_both_reach_goal(d, w.p1, w.p2, H, V);
}

// ---- step (quoridor.py:159-186, :193-202, :217-269) ---------------------------------------------------------
// Applies `action` for the mover WITHOUT checking legality (the reference's safe=False behaviour); a
// finished game is left untouched.  Returns the new state.
QZ_HD QzState qz_apply(QzState s, int action) {
    uint64_t m = s.meta;
    if (qz_done(m)) return s;
    int p1 = qz_p1(m), p2 = qz_p2(m), w1 = qz_w1(m), w2 = qz_w2(m), cur = qz_cur(m);
    unsigned flags = qz_flags(m), ply = qz_ply(m);
    if (action < 12) {
        int dlt = qz_delta(action);
        if (cur == 1) p1 += dlt; else p2 += dlt;
    } else {
        int a = action - 12;
        // quoridor.py:246-257 assigns the cell (+1 / -1): an unchecked placement on an occupied intersection
        // (safe=False) REPLACES the wall that was there
        if (a < 64) { s.H |= 1ull << a; s.V &= ~(1ull << a); }
        else { s.V |= 1ull << (a - 64); s.H &= ~(1ull << (a - 64)); }
        if (cur == 1) w1 -= 1; else w2 -= 1;
    }
    int winner = p2 < 9 ? 2 : (p1 > 71 ? 1 : 0);       // P2 tested first (:196-201)
    if (winner) flags |= QZ_FLAG_DONE | ((unsigned)winner << QZ_FLAG_WINNER_SHIFT);
    else cur = 3 - cur;                                  // mover is NOT rotated on a winning move (:176-181)
    ply = ply < 0xFFFFu ? ply + 1 : ply;
    s.meta = qz_pack_meta(p1, p2, w1, w2, cur, flags, ply);
    return s;
}

// true iff both pawns stand on the board (every non-terminal state; terminal ones may not)
QZ_HD bool qz_on_board(uint64_t m) {
    int p1 = qz_p1(m), p2 = qz_p2(m);
    return p1 >= 0 && p1 <= 80 && p2 >= 0 && p2 <= 80;
}

// pawn part of actions() for the mover (quoridor.py:146)
QZ_HD uint32_t qz_mover_pawn_moves_ctx(const QzPawnCtx &c, uint64_t meta) {
    const int cur = qz_cur(meta);
    const int L = cur == 1 ? qz_p1(meta) : qz_p2(meta);
    const int O = cur == 1 ? qz_p2(meta) : qz_p1(meta);
    return qz_pawn_moves_ctx(c, L, O, cur);
}
QZ_HD uint32_t qz_mover_pawn_moves(const QzState &s) {
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    return qz_mover_pawn_moves_ctx(c, s.meta);
}
QZ_HD int qz_mover_walls(uint64_t m) { return qz_cur(m) == 1 ? qz_w1(m) : qz_w2(m); }

// 140-bit mask assembly: pawn (12 bits) | Hlegal << 12 | Vlegal << 76
QZ_HD void qz_pack_mask(uint32_t pawn, uint64_t hl, uint64_t vl, uint64_t out[3]) {
    out[0] = (uint64_t)pawn | (hl << 12);
    out[1] = (hl >> 52) | (vl << 12);
    out[2] = vl >> 52;
}

// Sequential full sweep (one thread).  The warp kernel in qz_env.cu distributes the same
// qz_wall_keeps_paths calls over 32 lanes; this form serves single-thread users and the host harness.
QZ_HD void qz_legal_mask_seq(const QzState &s, uint64_t out[3]) {
    out[0] = out[1] = out[2] = 0;
    if (qz_done(s.meta) || !qz_on_board(s.meta)) return;
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    uint32_t pawn = qz_mover_pawn_moves_ctx(c, s.meta);
    uint64_t hl = 0, vl = 0;
    if (qz_mover_walls(s.meta) > 0) {
        QzSweep w = qz_sweep_prepare_ctx(c, s.H, s.V, qz_p1(s.meta), qz_p2(s.meta));
        uint64_t hc = qz_hcand(s.H, s.V), vc = qz_vcand(s.H, s.V);
        for (int ix = 0; ix < 64; ix++) {
            if (((hc >> ix) & 1) && qz_wall_keeps_paths(w, ix, false)) hl |= 1ull << ix;
            if (((vc >> ix) & 1) && qz_wall_keeps_paths(w, ix, true)) vl |= 1ull << ix;
        }
    }
    qz_pack_mask(pawn, hl, vl, out);
}

// Position of action `a` in the reference's actions() ordering given the legal mask parts:
// pawn ids ascending, then H(ix0),V(ix0),H(ix1),V(ix1),... (quoridor.py:157,420-430)
QZ_HD int qz_action_rank(uint32_t pawn, uint64_t hl, uint64_t vl, int a) {
    if (a < 12) return qz_popc32(pawn & ((1u << a) - 1u));
    int np = qz_popc32(pawn);
    if (a < 76) {
        int ix = a - 12;
        uint64_t below = (1ull << ix) - 1ull;
        return np + qz_popc64(hl & below) + qz_popc64(vl & below);
    }
    int ix = a - 76;
    uint64_t below = (1ull << ix) - 1ull;
    return np + qz_popc64(hl & below) + qz_popc64(vl & below) + (int)((hl >> ix) & 1ull);
}
QZ_HD void qz_unpack_mask(const uint64_t m[3], uint32_t &pawn, uint64_t &hl, uint64_t &vl) {
    pawn = (uint32_t)(m[0] & 0xFFFu);
    hl = (m[0] >> 12) | (m[1] << 52);
    vl = (m[1] >> 12) | (m[2] << 52);
}

// ---- state planes (quoridor.py:58-131): value of plane p at cell (r,c), as 0/1 -----------------------------
// planes: 0 no-wall, 1 vertical, 2 horizontal (8x8 padded to 9x9 at the bottom/right), 3 mover pawn,
// 4 opponent pawn, 5-14 mover walls-left one-hot at index w-1 (w=0 -> index 9), 15-24 opponent, 25 mover==P2.
// Off-board pawn tiles follow numpy's negative wrap for -81..-1; tiles > 80 (reference: IndexError) light nothing.
QZ_HD int qz_plane_value(const QzState &s, int p, int r, int c) {
    uint64_t m = s.meta;
    int cur = qz_cur(m);
    if (p <= 2) {
        if (r > 7 || c > 7) return 0;
        int i = r * 8 + c;
        int h = (int)((s.H >> i) & 1u), v = (int)((s.V >> i) & 1u);
        return p == 0 ? (1 - (h | v)) : (p == 1 ? v : h);
    }
    if (p <= 4) {
        int mine = cur == 1 ? qz_p1(m) : qz_p2(m), theirs = cur == 1 ? qz_p2(m) : qz_p1(m);
        int t = p == 3 ? mine : theirs;
        if (t < 0) t += 81;
        return t == r * 9 + c;
    }
    if (p <= 24) {
        int wm = cur == 1 ? qz_w1(m) : qz_w2(m), wo = cur == 1 ? qz_w2(m) : qz_w1(m);
        int w = p <= 14 ? wm : wo;
        int idx = w - 1 < 0 ? 9 : w - 1;
        return (p - (p <= 14 ? 5 : 15)) == idx;
    }
    return cur == 2;
}
