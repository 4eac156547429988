// TEST BUILD ONLY.  Compiles the product's __host__ __device__ rules header (the exact code the sm_100a
// kernels run) for the host so it can be diffed against the oracle in a container without a GPU.
// Nothing in the product loads this library; it is built and used by tests/test_rules_host.py alone.
#include "../../alphazero_quoridor_b200/csrc/qz_rules.cuh"

extern "C" {

void qh_initial(uint64_t *s3) { QzState s = qz_initial_state(); s3[0] = s.H; s3[1] = s.V; s3[2] = s.meta; }

uint64_t qh_pack_meta(int p1, int p2, int w1, int w2, int cur, unsigned flags, unsigned ply) {
    return qz_pack_meta(p1, p2, w1, w2, cur, flags, ply);
}

void qh_apply(uint64_t *s3, int action) {
    QzState s{s3[0], s3[1], s3[2]};
    s = qz_apply(s, action);
    s3[0] = s.H; s3[1] = s.V; s3[2] = s.meta;
}

void qh_legal_mask(const uint64_t *s3, uint64_t *mask3) {
    QzState s{s3[0], s3[1], s3[2]};
    qz_legal_mask_seq(s, mask3);
}

unsigned qh_pawn_moves(uint64_t H, uint64_t V, int L, int O, int player) { return qz_pawn_moves(H, V, L, O, player); }

// plain-move bits (N,S,E,W) of every tile from the direction masks, opponent ignored
void qh_dirs(uint64_t H, uint64_t V, unsigned char *out81) {
    QzDirs d = qz_dirs(H, V);
    for (int t = 0; t < 81; t++)
        out81[t] = (unsigned char)((bb_test(d.n, t) ? 1 : 0) | (bb_test(d.s, t) ? 2 : 0) | (bb_test(d.e, t) ? 4 : 0) |
                                   (bb_test(d.w, t) ? 8 : 0));
}

// same, but built incrementally: masks of (H0,V0) then one wall placed through qz_dirs_place_*
void qh_dirs_incremental(uint64_t H0, uint64_t V0, int ix, int vertical, unsigned char *out81) {
    QzDirs d = qz_dirs(H0, V0);
    if (vertical) qz_dirs_place_v(d, ix); else qz_dirs_place_h(d, ix);
    for (int t = 0; t < 81; t++)
        out81[t] = (unsigned char)((bb_test(d.n, t) ? 1 : 0) | (bb_test(d.s, t) ? 2 : 0) | (bb_test(d.e, t) ? 4 : 0) |
                                   (bb_test(d.w, t) ? 8 : 0));
}

unsigned qh_pawn_moves_ctx(uint64_t H, uint64_t V, int L, int O, int player) {
    QzPawnCtx c = qz_ctx_build(H, V);
    return qz_pawn_moves_ctx(c, L, O, player);
}

// corner codes of every tile read back from the masks: nw | ne<<2 | se<<4 | sw<<6 (0 none, 1 H, 2 V)
void qh_ctx_corners(uint64_t H, uint64_t V, unsigned char *out81, unsigned char *scalar81) {
    QzPawnCtx c = qz_ctx_build(H, V);
    for (int t = 0; t < 81; t++) {
        unsigned nw = bb_at(c.nwH, t) | (bb_at(c.nwV, t) << 1), ne = bb_at(c.neH, t) | (bb_at(c.neV, t) << 1);
        unsigned se = bb_at(c.seH, t) | (bb_at(c.seV, t) << 1), sw = bb_at(c.swH, t) | (bb_at(c.swV, t) << 1);
        out81[t] = (unsigned char)(nw | (ne << 2) | (se << 4) | (sw << 6));
        QzCorners k = qz_corners(H, V, t);
        scalar81[t] = (unsigned char)(k.nw | (k.ne << 2) | (k.se << 4) | (k.sw << 6));
    }
}

// pawn moves through the per-tile info table (the pawn-phase rollout kernel's path)
unsigned qh_pawn_moves_info(uint64_t H, uint64_t V, int L, int O, int player) {
    QzPawnCtx c = qz_ctx_build(H, V);
    uint32_t tbl[QZ_TILE_TABLE_WORDS];
    qz_tile_table(c, tbl, 1);
    const unsigned char *b = reinterpret_cast<const unsigned char *>(tbl);
    const uint32_t hO = bb_at(c.neH, O) | (bb_at(c.nwH, O) << 1) | (bb_at(c.seH, O) << 2) | (bb_at(c.swH, O) << 3);
    return qz_pawn_moves_info(b[L], b[O], hO, L, O, player);
}

// the byte-per-tile table of the pawn-phase kernel next to the same eight bits read straight from the ctx masks
void qh_tile_table(uint64_t H, uint64_t V, unsigned char *table84, unsigned char *direct81) {
    QzPawnCtx c = qz_ctx_build(H, V);
    uint32_t tbl[QZ_TILE_TABLE_WORDS];
    qz_tile_table(c, tbl, 1);
    for (int i = 0; i < 84; i++) table84[i] = reinterpret_cast<const unsigned char *>(tbl)[i];
    for (int t = 0; t < 81; t++)
        direct81[t] = (unsigned char)(bb_at(c.d.n, t) | (bb_at(c.d.s, t) << 1) | (bb_at(c.d.e, t) << 2) | (bb_at(c.d.w, t) << 3) |
                                      (bb_at(c.neV, t) << 4) | (bb_at(c.nwV, t) << 5) | (bb_at(c.seV, t) << 6) | (bb_at(c.swV, t) << 7));
}
// qz_ctx_store / qz_ctx_load round trip (the stuck kernel parks the masks in shared memory)
int qh_ctx_roundtrip(uint64_t H, uint64_t V) {
    QzPawnCtx c = qz_ctx_build(H, V);
    uint32_t w[QZ_CTX_WORDS];
    qz_ctx_store(c, w);
    QzPawnCtx d = qz_ctx_load(w);
    int ok = 1;
    for (int L = 0; L < 81; L++)
        for (int k = 0; k < 4; k++) {
            const int O = L + (k == 0 ? 9 : (k == 1 ? -9 : (k == 2 ? 1 : -1)));
            if (O < 0 || O > 80) continue;
            ok &= qz_pawn_moves_ctx(c, L, O, 1) == qz_pawn_moves_ctx(d, L, O, 1);
            ok &= qz_pawn_moves_ctx(c, L, O, 2) == qz_pawn_moves_ctx(d, L, O, 2);
        }
    return ok;
}

void qh_dirs_ctx(uint64_t H, uint64_t V, unsigned char *out81) {
    QzPawnCtx c = qz_ctx_build(H, V);
    for (int t = 0; t < 81; t++)
        out81[t] = (unsigned char)(bb_at(c.d.n, t) | (bb_at(c.d.s, t) << 1) | (bb_at(c.d.e, t) << 2) | (bb_at(c.d.w, t) << 3));
}

void qh_spread8(uint64_t x, uint32_t *w3) { BB b = bb_spread8(x); w3[0] = b.w0; w3[1] = b.w1; w3[2] = b.w2; }

void qh_encode(const uint64_t *s3, float *out) {
    QzState s{s3[0], s3[1], s3[2]};
    for (int p = 0; p < 26; p++)
        for (int r = 0; r < 9; r++)
            for (int c = 0; c < 9; c++) out[p * 81 + r * 9 + c] = (float)qz_plane_value(s, p, r, c);
}

int qh_action_rank(const uint64_t *mask3, int a) {
    uint32_t pawn; uint64_t hl, vl;
    qz_unpack_mask(mask3, pawn, hl, vl);
    return qz_action_rank(pawn, hl, vl, a);
}

int qh_nth_bit64(uint64_t m, int k) { return qz_nth_bit64(m, k); }
int qh_delta(int a) { return qz_delta(a); }
}

#include "../../alphazero_quoridor_b200/csrc/qz_sample.cuh"
extern "C" {
void qh_philox(uint64_t seed, uint64_t rid, uint32_t c2, uint32_t c3, uint32_t *out4) {
    QzPhilox4 b = qz_philox(seed, rid, c2, c3);
    out4[0] = b.x; out4[1] = b.y; out4[2] = b.z; out4[3] = b.w;
}
int qh_sample_action(const uint64_t *s3, uint64_t seed, uint64_t rid, uint32_t ply) {
    QzState s{s3[0], s3[1], s3[2]};
    QzRng rng = qz_rng_init(seed, rid);
    return qz_sample_action(s, rng, ply);
}
// capped + table-driven sampling must agree with the plain one (-2 = cap hit)
int qh_sample_action_known(const uint64_t *s3, uint64_t seed, uint64_t rid, uint32_t ply) {
    QzState s{s3[0], s3[1], s3[2]};
    uint64_t m[3];
    qz_legal_mask_seq(s, m);
    uint32_t pawn; uint64_t hl, vl;
    qz_unpack_mask(m, pawn, hl, vl);
    QzRng rng = qz_rng_init(seed, rid);
    return qz_sample_action_known(s, rng, ply, pawn, hl, vl);
}
int qh_sample_action_capped(const uint64_t *s3, uint64_t seed, uint64_t rid, uint32_t ply, uint32_t cap) {
    QzState s{s3[0], s3[1], s3[2]};
    QzRng rng = qz_rng_init(seed, rid);
    return qz_sample_action_capped(s, rng, ply, cap);
}
int qh_rollout(uint64_t *s3, uint64_t seed, uint64_t rid, int limit, int *plies) {
    QzState s{s3[0], s3[1], s3[2]};
    int v = qz_rollout(s, seed, rid, limit, *plies);
    s3[0] = s.H; s3[1] = s.V; s3[2] = s.meta;
    return v;
}
}
