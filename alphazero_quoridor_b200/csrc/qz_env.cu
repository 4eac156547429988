// qz_env.cu -- batched Quoridor environment kernels (K1 step, K2 legal mask, K3 state encoder).
// Reference methods replaced: Quoridor.reset/step/actions/state (quoridor.py:26-186); see include/qzb200.h.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "qz_common.cuh"
#include "qz_warp.cuh"

// ------------------------------------------------------------------------------------------ reset
__global__ void qz_reset_kernel(qz_state *__restrict__ states, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    qz_store_state(states + i, qz_initial_state());
}

extern "C" int qz_env_reset(qz_state *states, int64_t n, void *stream) {
    QZ_REQUIRE(n >= 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_ALIGN(states, 8);
    qz_reset_kernel<<<qz_blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(states, n);
    return qz_check_launch("qz_env_reset");
}

// ------------------------------------------------------------------------------------------ step
// One game per thread: 24 B read + 4 B action (+24 B mask) -> 24 B write + 1 B done.
__global__ void qz_step_kernel(qz_state *__restrict__ states, const int32_t *__restrict__ actions,
                               const uint64_t *__restrict__ mask3, uint8_t *__restrict__ done, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    QzState s = qz_load_state(states + i);
    const int a = __ldg(actions + i);
    if (a >= 0 && !qz_done(s.meta)) {
        bool ok = a < QZ_N_ACTIONS;
        if (ok && mask3 != nullptr) ok = (__ldg(mask3 + 3 * i + (a >> 6)) >> (a & 63)) & 1ull;
        if (ok) {
            s = qz_apply(s, a);
            s.meta &= ~((uint64_t)QZ_FLAG_ILLEGAL << 40);
        } else {
            s.meta |= (uint64_t)QZ_FLAG_ILLEGAL << 40;
        }
        qz_store_state(states + i, s);
    }
    if (done != nullptr) done[i] = qz_done(s.meta) ? 1 : 0;
}

extern "C" int qz_env_step(qz_state *states, const int32_t *actions, const uint64_t *legal_mask3, uint8_t *done,
                           int64_t n, void *stream) {
    QZ_REQUIRE(n >= 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(actions);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(actions, 4);
    QZ_REQUIRE_ALIGN(legal_mask3, 8);
    qz_step_kernel<<<qz_blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(states, actions, legal_mask3, done, n);
    return qz_check_launch("qz_env_step");
}

// ------------------------------------------------------------------------------------------ legal mask
// One warp per game (4 games per 128-thread block).
__global__ void __launch_bounds__(128, 4) qz_legal_mask_kernel(const qz_state *__restrict__ states,
                                                            uint64_t *__restrict__ mask3, int64_t n) {
    __shared__ uint32_t scratch[4][QZ_WARP_SCRATCH_WORDS];
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    const QzState s = qz_load_state(states + g);
    uint32_t pawn; uint64_t hl, vl;
    qz_warp_legal(s, pawn, hl, vl, scratch[threadIdx.x >> 5]);
    const int lane = threadIdx.x & 31;
    if (lane < 3) {
        uint64_t out[3];
        qz_pack_mask(pawn, hl, vl, out);
        mask3[3 * g + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
    }
}

// The same sweep for the flagged positions only (lazy expansion, qz_mcts_extend: the leaves whose descent stopped at a
// node that needs its legal set now); the masks of the others are left untouched.
__global__ void __launch_bounds__(128, 4) qz_legal_mask_flagged_kernel(const qz_state *__restrict__ states,
                                                                    uint8_t *__restrict__ flags, int flag_bits,
                                                                    uint64_t *__restrict__ mask3, int64_t n,
                                                                    const int32_t *__restrict__ key, int group, int dup_bits) {
    __shared__ uint32_t scratch[4][QZ_WARP_SCRATCH_WORDS];
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    if (!(flags[g] & flag_bits)) return;
    if (key != nullptr && group > 1) {
        // an EARLIER flagged position of the same group with the same key (the leaves of one game that stopped at the
        // same tree node) gets the sweep; this one is marked with dup_bits and skipped
        const int64_t first = g - g % group;
        const int32_t mine = key[g];
        bool dup = false;
        for (int64_t j = first + (threadIdx.x & 31); j < g; j += 32) dup |= (flags[j] & flag_bits) && key[j] == mine;
        if (__any_sync(0xFFFFFFFFu, dup)) {
            if ((threadIdx.x & 31) == 0) flags[g] |= (uint8_t)dup_bits;
            return;
        }
    }
    const QzState s = qz_load_state(states + g);
    uint32_t pawn; uint64_t hl, vl;
    qz_warp_legal(s, pawn, hl, vl, scratch[threadIdx.x >> 5]);
    const int lane = threadIdx.x & 31;
    if (lane < 3) {
        uint64_t out[3];
        qz_pack_mask(pawn, hl, vl, out);
        mask3[3 * g + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
    }
}

extern "C" int qz_env_legal_mask_flagged(const qz_state *states, uint8_t *flags, int flag_bits, uint64_t *mask3,
                                         int64_t n, const int32_t *key, int32_t group, int dup_bits, void *stream) {
    QZ_REQUIRE(n >= 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(flags);
    QZ_REQUIRE_PTR(mask3);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(mask3, 8);
    QZ_REQUIRE(group >= 0);
    QZ_REQUIRE((dup_bits & flag_bits) == 0);
    qz_legal_mask_flagged_kernel<<<qz_blocks_for(n, 4), 128, 0, (cudaStream_t)stream>>>(states, flags, flag_bits, mask3, n, key,
                                                                                       group, dup_bits);
    return qz_check_launch("qz_env_legal_mask_flagged");
}

extern "C" int qz_env_legal_mask(const qz_state *states, uint64_t *mask3, int64_t n, void *stream) {
    QZ_REQUIRE(n >= 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(mask3);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(mask3, 8);
    qz_legal_mask_kernel<<<qz_blocks_for(n, 4), 128, 0, (cudaStream_t)stream>>>(states, mask3, n);
    return qz_check_launch("qz_env_legal_mask");
}

// ------------------------------------------------------------------------------------------ uniform legal pick
// The reference's random policy (pure_mcts.py:7-10: argmax of iid U(0,1) over actions() == a uniform legal action)
// given the FULL legal mask: action = k-th set bit of the mask, k = (word * count) >> 32, word = Philox4x32-10(key =
// seed; counter = (game id, ply >> 2, 0x7000))[ply & 3] with ply read from the state.  One game per thread.
// Finished games and games without a legal action get -1.
#include "qz_philox.cuh"

__global__ void qz_sample_legal_kernel(const qz_state *__restrict__ states, const uint64_t *__restrict__ mask3, uint64_t seed,
                                       const int64_t *__restrict__ game_id, int32_t *__restrict__ actions, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t meta = __ldg(reinterpret_cast<const uint64_t *>(states + i) + 2);
    const uint64_t m0 = __ldg(mask3 + 3 * i), m1 = __ldg(mask3 + 3 * i + 1), m2 = __ldg(mask3 + 3 * i + 2);
    const int c0 = qz_popc64(m0), c1 = qz_popc64(m1), cnt = c0 + c1 + qz_popc64(m2);
    int act = -1;
    if (!qz_done(meta) && cnt > 0) {
        const uint32_t ply = qz_ply(meta);
        const QzPhilox4 b = qz_philox(seed, game_id ? (uint64_t)game_id[i] : (uint64_t)i, ply >> 2, 0x7000u);
        int k = (int)qz_mulhi32(qz_philox_word(b, (int)(ply & 3u)), (uint32_t)cnt);
        if (k < c0) act = qz_nth_bit64(m0, k);
        else if (k < c0 + c1) act = 64 + qz_nth_bit64(m1, k - c0);
        else act = 128 + qz_nth_bit64(m2, k - c0 - c1);
    }
    actions[i] = act;
}

extern "C" int qz_env_sample_legal(const qz_state *states, const uint64_t *mask3, uint64_t seed, const int64_t *game_id,
                                   int32_t *actions, int64_t n, void *stream) {
    QZ_REQUIRE(n >= 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(mask3);
    QZ_REQUIRE_PTR(actions);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(mask3, 8);
    qz_sample_legal_kernel<<<qz_blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(states, mask3, seed, game_id, actions, n);
    return qz_check_launch("qz_env_sample_legal");
}

// ------------------------------------------------------------------------------------------ whole random games
// BASELINE config 1 in two launches: uniform-random legal play with the FULL legal set computed on every ply and the
// pick rule of qz_env_sample_legal, played to the end (or max_plies) inside the kernels -- the same games, ply for ply,
// as the legal_mask / sample_legal / step loop.  While a wall is still in hand a game is driven by one warp (the
// 128-candidate sweep of qz_warp_legal per ply); afterwards by one thread (twelve corner masks built once, a pawn-move
// query per ply).
__device__ __forceinline__ int qz_pick_from_mask(uint32_t pawn, uint64_t hl, uint64_t vl, uint64_t seed, uint64_t gid, uint32_t ply) {
    uint64_t m[3];
    qz_pack_mask(pawn, hl, vl, m);
    const int c0 = qz_popc64(m[0]), c1 = qz_popc64(m[1]), cnt = c0 + c1 + qz_popc64(m[2]);
    if (cnt == 0) return -1;
    const QzPhilox4 b = qz_philox(seed, gid, ply >> 2, 0x7000u);
    const int k = (int)qz_mulhi32(qz_philox_word(b, (int)(ply & 3u)), (uint32_t)cnt);
    if (k < c0) return qz_nth_bit64(m[0], k);
    if (k < c0 + c1) return 64 + qz_nth_bit64(m[1], k - c0);
    return 128 + qz_nth_bit64(m[2], k - c0 - c1);
}

__global__ void __launch_bounds__(128, 4) qz_random_play_wall_kernel(qz_state *__restrict__ states, uint64_t seed,
                                                                     const int64_t *__restrict__ game_id, int max_plies, int64_t n) {
    __shared__ uint32_t scratch[4][QZ_WARP_SCRATCH_WORDS];
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    QzState s = qz_load_state(states + g);
    const uint64_t gid = game_id ? (uint64_t)game_id[g] : (uint64_t)g;
    while (!qz_done(s.meta) && !(qz_flags(s.meta) & QZ_FLAG_STALEMATE) && (int)qz_ply(s.meta) < max_plies &&
           qz_w1(s.meta) + qz_w2(s.meta) > 0) {
        uint32_t pawn; uint64_t hl, vl;
        qz_warp_legal(s, pawn, hl, vl, scratch[threadIdx.x >> 5]);
        const int act = qz_pick_from_mask(pawn, hl, vl, seed, gid, qz_ply(s.meta));
        if (act < 0) s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40;
        else s = qz_apply(s, act);
    }
    if ((threadIdx.x & 31) == 0) qz_store_state(states + g, s);
}

__global__ void __launch_bounds__(128) qz_random_play_pawn_kernel(qz_state *__restrict__ states, uint64_t seed,
                                                                  const int64_t *__restrict__ game_id, int max_plies, int64_t n) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    QzState s = qz_load_state(states + g);
    if (qz_done(s.meta) || (qz_flags(s.meta) & QZ_FLAG_STALEMATE) || qz_w1(s.meta) + qz_w2(s.meta) > 0) return;
    const uint64_t gid = game_id ? (uint64_t)game_id[g] : (uint64_t)g;
    const QzPawnCtx c = qz_ctx_build(s.H, s.V);
    while (!qz_done(s.meta) && (int)qz_ply(s.meta) < max_plies) {
        const int act = qz_pick_from_mask(qz_mover_pawn_moves_ctx(c, s.meta), 0, 0, seed, gid, qz_ply(s.meta));
        if (act < 0) { s.meta |= (uint64_t)QZ_FLAG_STALEMATE << 40; break; }
        s = qz_apply(s, act);
    }
    qz_store_state(states + g, s);
}

extern "C" int qz_env_random_play(qz_state *states, uint64_t seed, const int64_t *game_id, int32_t max_plies, int64_t n,
                                  void *stream) {
    QZ_REQUIRE(n >= 0 && max_plies >= 0 && max_plies <= 65535);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_ALIGN(states, 8);
    cudaStream_t st = (cudaStream_t)stream;
    qz_random_play_wall_kernel<<<qz_blocks_for(n, 4), 128, 0, st>>>(states, seed, game_id, max_plies, n);
    int rc = qz_check_launch("qz_env_random_play (wall phase)");
    if (rc) return rc;
    qz_random_play_pawn_kernel<<<qz_blocks_for(n, 128), 128, 0, st>>>(states, seed, game_id, max_plies, n);
    return qz_check_launch("qz_env_random_play (pawn phase)");
}

// ------------------------------------------------------------------------------------------ encode
// HBM-bound by design: 24 B read and 2106 (NCHW) / 81*c_stride (NHWC) elements written per game.
// A block of 8 warps encodes 8 games.  Phase 1 builds the block's output as a packed BIT stream in shared
// memory -- NCHW: 26 plane masks of 81 bits OR-ed in at bit g*2106 + p*81; NHWC with 32 channels: one 32-bit
// "planes hot on this tile" word per tile.  Phase 2 is then layout-agnostic: output chunk q (8 two-byte or
// 4 four-byte elements = 16 B) is byte / nibble q of the stream, expanded with integer multiplies and
// written with one fully coalesced 16-byte store per lane (~1.4 instructions per byte written).
template <typename T> struct QzOne;
template <> struct QzOne<float> { static __device__ __forceinline__ uint32_t bits() { return 0x3F800000u; } };
template <> struct QzOne<__nv_bfloat16> { static __device__ __forceinline__ uint32_t bits() { return 0x3F80u; } };
template <> struct QzOne<__half> { static __device__ __forceinline__ uint32_t bits() { return 0x3C00u; } };

#define QZ_ENC_GAMES 8
#define QZ_ENC_WORDS_NCHW ((QZ_ENC_GAMES * QZ_STATE_ELEMS + 31) / 32 + 1)      // 527 + pad
#define QZ_ENC_WORDS_NHWC (QZ_ENC_GAMES * 81)

struct QzEncodeView {
    int mine, theirs;       // tiles of the mover's / opponent's pawn (numpy-wrapped; > 80 lights nothing)
    uint32_t cbits;         // planes 5..25 that are constant over the board
};

__device__ __forceinline__ QzEncodeView qz_encode_view(const QzState &s) {
    const uint64_t m = s.meta;
    const int cur = qz_cur(m);
    QzEncodeView v;
    v.mine = cur == 1 ? qz_p1(m) : qz_p2(m);
    v.theirs = cur == 1 ? qz_p2(m) : qz_p1(m);
    if (v.mine < 0) v.mine += 81;                  // numpy negative index wrap (quoridor.py:69,73)
    if (v.theirs < 0) v.theirs += 81;
    const int wm = cur == 1 ? qz_w1(m) : qz_w2(m), wo = cur == 1 ? qz_w2(m) : qz_w1(m);
    const int im = wm - 1 < 0 ? 9 : (wm - 1 > 9 ? 9 : wm - 1);      // index -1 wraps to plane 9 (:79-80)
    const int io = wo - 1 < 0 ? 9 : (wo - 1 > 9 ? 9 : wo - 1);
    v.cbits = (1u << (5 + im)) | (1u << (15 + io)) | (cur == 2 ? 1u << 25 : 0u);
    return v;
}

// 81-bit mask of plane p (quoridor.py:58-131)
__device__ __forceinline__ BB qz_plane_mask(const QzState &s, const QzEncodeView &v, int p) {
    if (p == 0) return bb_spread8(~(s.H | s.V));
    if (p == 1) return bb_spread8(s.V);
    if (p == 2) return bb_spread8(s.H);
    if (p == 3) return (unsigned)v.mine <= 80u ? bb_bit(v.mine) : bb_zero();
    if (p == 4) return (unsigned)v.theirs <= 80u ? bb_bit(v.theirs) : bb_zero();
    return ((v.cbits >> p) & 1u) ? bb_make(0xFFFFFFFFu, 0xFFFFFFFFu, QZ_BOARD_W2) : bb_zero();
}

__device__ __forceinline__ uint32_t qz_tile_word(const QzState &s, int t, const QzEncodeView &v) {
    const int r = t / 9, c = t - 9 * r;
    uint32_t w = v.cbits;
    if (r < 8 && c < 8) {
        const int i = r * 8 + c;
        const uint32_t h = (uint32_t)(s.H >> i) & 1u, vv = (uint32_t)(s.V >> i) & 1u;
        w |= h ? 4u : (vv ? 2u : 1u);
    }
    if (t == v.mine) w |= 8u;
    if (t == v.theirs) w |= 16u;
    return w;
}

// OR an 81-bit mask into a shared bit stream at bit offset `off`
__device__ __forceinline__ void qz_or_bits(uint32_t *stream, int off, const BB &b) {
    const int w = off >> 5, sh = off & 31;
    const uint64_t lo = (uint64_t)b.w0 << sh, mid = (uint64_t)b.w1 << sh, hi = (uint64_t)b.w2 << sh;
    const uint32_t x0 = (uint32_t)lo, x1 = (uint32_t)(lo >> 32) | (uint32_t)mid;
    const uint32_t x2 = (uint32_t)(mid >> 32) | (uint32_t)hi, x3 = (uint32_t)(hi >> 32);
    if (x0) atomicOr(stream + w, x0);
    if (x1) atomicOr(stream + w + 1, x1);
    if (x2) atomicOr(stream + w + 2, x2);
    if (x3) atomicOr(stream + w + 3, x3);
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) qz_encode_kernel(const qz_state *__restrict__ states, T *__restrict__ out, int64_t n) {
    constexpr int WORDS = LAYOUT == QZ_LAYOUT_NCHW ? QZ_ENC_WORDS_NCHW : QZ_ENC_WORDS_NHWC;
    constexpr int ELEMS_PER_GAME = LAYOUT == QZ_LAYOUT_NCHW ? QZ_STATE_ELEMS : 81 * 32;
    constexpr int PER_CHUNK = 16 / (int)sizeof(T);                      // elements per 16-byte store
    __shared__ __align__(16) uint32_t stream[WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = (int64_t)blockIdx.x * QZ_ENC_GAMES;
    const int games = (int)min((int64_t)QZ_ENC_GAMES, n - g0);
    if (LAYOUT == QZ_LAYOUT_NCHW) {
        for (int i = threadIdx.x; i < WORDS; i += 256) stream[i] = 0;
        __syncthreads();
    }
    if (warp < games) {
        const QzState s = qz_load_state(states + g0 + warp);
        const QzEncodeView v = qz_encode_view(s);
        if (LAYOUT == QZ_LAYOUT_NCHW) {
            if (lane < QZ_N_PLANES) qz_or_bits(stream, warp * QZ_STATE_ELEMS + lane * 81, qz_plane_mask(s, v, lane));
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int t = lane + 32 * k;
                if (t < 81) stream[warp * 81 + t] = qz_tile_word(s, t, v);
            }
        }
    }
    __syncthreads();
    const uint32_t one = QzOne<T>::bits();
    T *base = out + g0 * (int64_t)ELEMS_PER_GAME;
    const int total = games * ELEMS_PER_GAME;
    const int full_chunks = total / PER_CHUNK;
    const uint8_t *bytes = reinterpret_cast<const uint8_t *>(stream);
    for (int q = threadIdx.x; q < full_chunks; q += 256) {
        uint4 o;
        if (sizeof(T) == 2) {
            const uint32_t b = bytes[q];
            o.x = (b & 1u) * one + ((b >> 1) & 1u) * (one << 16);
            o.y = ((b >> 2) & 1u) * one + ((b >> 3) & 1u) * (one << 16);
            o.z = ((b >> 4) & 1u) * one + ((b >> 5) & 1u) * (one << 16);
            o.w = ((b >> 6) & 1u) * one + ((b >> 7) & 1u) * (one << 16);
        } else {
            const uint32_t b = (bytes[q >> 1] >> ((q & 1) * 4)) & 0xFu;
            o.x = (b & 1u) * one; o.y = ((b >> 1) & 1u) * one; o.z = ((b >> 2) & 1u) * one; o.w = ((b >> 3) & 1u) * one;
        }
        *reinterpret_cast<uint4 *>(base + (int64_t)q * PER_CHUNK) = o;
    }
    // tail of a partial last block (total not a multiple of the chunk size)
    for (int e = full_chunks * PER_CHUNK + threadIdx.x; e < total; e += 256) {
        const uint32_t bit = (stream[e >> 5] >> (e & 31)) & 1u;
        if (sizeof(T) == 2) reinterpret_cast<uint16_t *>(base)[e] = (uint16_t)(bit * one);
        else reinterpret_cast<uint32_t *>(base)[e] = bit * one;
    }
}

// generic channel stride for NHWC (c_stride != 32): one warp per game, 4-byte stores
template <typename T>
__global__ void __launch_bounds__(256) qz_encode_nhwc_generic_kernel(const qz_state *__restrict__ states, T *__restrict__ out,
                                                                     int c_stride, int64_t n) {
    __shared__ uint32_t words[8][84];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * 8 + warp;
    if (g >= n) return;
    const QzState s = qz_load_state(states + g);
    const QzEncodeView v = qz_encode_view(s);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int t = lane + 32 * k;
        if (t < 81) words[warp][t] = qz_tile_word(s, t, v);
    }
    __syncwarp();
    const uint32_t one = QzOne<T>::bits();
    T *base = out + g * (int64_t)81 * c_stride;          // c_stride is even (checked by the host)
    const int total = 81 * c_stride / 2;
    for (int q = lane; q < total; q += 32) {
        const int e0 = 2 * q;
        const int t = e0 / c_stride, c0 = e0 - t * c_stride;
        const uint32_t wv = words[warp][t];
        const uint32_t b0 = c0 < 26 ? (wv >> c0) & 1u : 0u, b1 = c0 + 1 < 26 ? (wv >> (c0 + 1)) & 1u : 0u;
        if (sizeof(T) == 4) {
            uint2 o; o.x = b0 * one; o.y = b1 * one;
            *reinterpret_cast<uint2 *>(base + e0) = o;
        } else {
            *reinterpret_cast<uint32_t *>(base + e0) = b0 * one | (b1 * one) << 16;
        }
    }
}

template <typename T>
static int qz_encode_launch(const qz_state *states, void *out, int layout, int c_stride, int64_t n, cudaStream_t st) {
    const unsigned blocks = qz_blocks_for(n, QZ_ENC_GAMES);
    if (layout == QZ_LAYOUT_NCHW)
        qz_encode_kernel<T, QZ_LAYOUT_NCHW><<<blocks, 256, 0, st>>>(states, (T *)out, n);
    else if (c_stride == 32)
        qz_encode_kernel<T, QZ_LAYOUT_NHWC><<<blocks, 256, 0, st>>>(states, (T *)out, n);
    else
        qz_encode_nhwc_generic_kernel<T><<<blocks, 256, 0, st>>>(states, (T *)out, c_stride, n);
    return qz_check_launch("qz_env_encode");
}

extern "C" int qz_env_encode(const qz_state *states, void *out, int dtype, int layout, int c_stride, int64_t n,
                             void *stream) {
    QZ_REQUIRE(n >= 0);
    QZ_REQUIRE(dtype == QZ_DTYPE_F32 || dtype == QZ_DTYPE_BF16 || dtype == QZ_DTYPE_F16);
    QZ_REQUIRE(layout == QZ_LAYOUT_NCHW || layout == QZ_LAYOUT_NHWC);
    if (layout == QZ_LAYOUT_NHWC) QZ_REQUIRE(c_stride >= 26 && c_stride <= 64 && (c_stride % 2) == 0);
    if (n == 0) return 0;
    QZ_REQUIRE_PTR(states);
    QZ_REQUIRE_PTR(out);
    QZ_REQUIRE_ALIGN(states, 8);
    QZ_REQUIRE_ALIGN(out, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == QZ_DTYPE_F32) return qz_encode_launch<float>(states, out, layout, c_stride, n, st);
    if (dtype == QZ_DTYPE_BF16) return qz_encode_launch<__nv_bfloat16>(states, out, layout, c_stride, n, st);
    return qz_encode_launch<__half>(states, out, layout, c_stride, n, st);
}
