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
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    const QzState s = qz_load_state(states + g);
    uint32_t pawn; uint64_t hl, vl;
    qz_warp_legal(s, pawn, hl, vl);
    const int lane = threadIdx.x & 31;
    if (lane < 3) {
        uint64_t out[3];
        qz_pack_mask(pawn, hl, vl, out);
        mask3[3 * g + lane] = lane == 0 ? out[0] : (lane == 1 ? out[1] : out[2]);
    }
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

// ------------------------------------------------------------------------------------------ encode
// One warp per game.  The warp first builds, per tile, the 26-bit "which planes are hot here" word in
// shared memory (81 words), then streams the tensor out with fully coalesced 4- or 8-byte stores.
// HBM-bound by design: 2106 elements written per game, 24 B read.
template <typename T> struct QzOne;
template <> struct QzOne<float> { static __device__ __forceinline__ uint32_t bits() { return 0x3F800000u; } };
template <> struct QzOne<__nv_bfloat16> { static __device__ __forceinline__ uint32_t bits() { return 0x3F80u; } };
template <> struct QzOne<__half> { static __device__ __forceinline__ uint32_t bits() { return 0x3C00u; } };

__device__ __forceinline__ uint32_t qz_tile_word(const QzState &s, int t, int mine, int theirs, uint32_t cbits) {
    const int r = t / 9, c = t - 9 * r;
    uint32_t w = cbits;
    if (r < 8 && c < 8) {
        const int i = r * 8 + c;
        const uint32_t h = (uint32_t)(s.H >> i) & 1u, v = (uint32_t)(s.V >> i) & 1u;
        w |= h ? 4u : (v ? 2u : 1u);
    }
    if (t == mine) w |= 8u;
    if (t == theirs) w |= 16u;
    return w;
}

template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) qz_encode_kernel(const qz_state *__restrict__ states, T *__restrict__ out,
                                                        int c_stride, int64_t n) {
    __shared__ uint32_t words[8][84];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * 8 + warp;
    if (g >= n) return;
    const QzState s = qz_load_state(states + g);
    const uint64_t m = s.meta;
    const int cur = qz_cur(m);
    int mine = cur == 1 ? qz_p1(m) : qz_p2(m), theirs = cur == 1 ? qz_p2(m) : qz_p1(m);
    if (mine < 0) mine += 81;                 // numpy negative index wrap (quoridor.py:69,73)
    if (theirs < 0) theirs += 81;
    const int wm = cur == 1 ? qz_w1(m) : qz_w2(m), wo = cur == 1 ? qz_w2(m) : qz_w1(m);
    const int im = wm - 1 < 0 ? 9 : (wm - 1 > 9 ? 9 : wm - 1);      // index -1 wraps to plane 9 (:79-80)
    const int io = wo - 1 < 0 ? 9 : (wo - 1 > 9 ? 9 : wo - 1);
    const uint32_t cbits = (1u << (5 + im)) | (1u << (15 + io)) | (cur == 2 ? 1u << 25 : 0u);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int t = lane + 32 * k;
        if (t < 81) words[warp][t] = qz_tile_word(s, t, mine, theirs, cbits);
    }
    __syncwarp();
    const uint32_t one = QzOne<T>::bits();
    if (LAYOUT == QZ_LAYOUT_NCHW) {
        T *base = out + g * (int64_t)QZ_STATE_ELEMS;
        // element pairs (e, e+1), e even; 2106 is even so pairs never straddle games
        for (int q = lane; q < QZ_STATE_ELEMS / 2; q += 32) {
            const int e0 = 2 * q, e1 = e0 + 1;
            const int p0 = e0 / 81, t0 = e0 - 81 * p0;
            const int p1 = e1 / 81, t1 = e1 - 81 * p1;
            const uint32_t b0 = (words[warp][t0] >> p0) & 1u, b1 = (words[warp][t1] >> p1) & 1u;
            if (sizeof(T) == 4) {
                uint2 v; v.x = b0 ? one : 0u; v.y = b1 ? one : 0u;
                *reinterpret_cast<uint2 *>(base + e0) = v;
            } else {
                *reinterpret_cast<uint32_t *>(base + e0) = (b0 ? one : 0u) | ((b1 ? one : 0u) << 16);
            }
        }
    } else {
        T *base = out + g * (int64_t)81 * c_stride;      // c_stride is even (checked by the host)
        const int total = 81 * c_stride / 2;
        for (int q = lane; q < total; q += 32) {
            const int e0 = 2 * q;
            const int t = e0 / c_stride, c0 = e0 - t * c_stride;
            const uint32_t wv = words[warp][t];
            const uint32_t b0 = c0 < 26 ? (wv >> c0) & 1u : 0u, b1 = c0 + 1 < 26 ? (wv >> (c0 + 1)) & 1u : 0u;
            if (sizeof(T) == 4) {
                uint2 v; v.x = b0 ? one : 0u; v.y = b1 ? one : 0u;
                *reinterpret_cast<uint2 *>(base + e0) = v;
            } else {
                *reinterpret_cast<uint32_t *>(base + e0) = (b0 ? one : 0u) | ((b1 ? one : 0u) << 16);
            }
        }
    }
}

template <typename T>
static int qz_encode_launch(const qz_state *states, void *out, int layout, int c_stride, int64_t n, cudaStream_t st) {
    const unsigned blocks = qz_blocks_for(n, 8);
    if (layout == QZ_LAYOUT_NCHW)
        qz_encode_kernel<T, QZ_LAYOUT_NCHW><<<blocks, 256, 0, st>>>(states, (T *)out, 26, n);
    else
        qz_encode_kernel<T, QZ_LAYOUT_NHWC><<<blocks, 256, 0, st>>>(states, (T *)out, c_stride, n);
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
    QZ_REQUIRE_ALIGN(out, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == QZ_DTYPE_F32) return qz_encode_launch<float>(states, out, layout, c_stride, n, st);
    if (dtype == QZ_DTYPE_BF16) return qz_encode_launch<__nv_bfloat16>(states, out, layout, c_stride, n, st);
    return qz_encode_launch<__half>(states, out, layout, c_stride, n, st);
}
