// qz_common.cuh -- launch / error plumbing shared by the translation units of libqzb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qzb200.h"
#include "qz_rules.cuh"

#define QZ_WARP 32

// thread-local error text (qz_last_error_string)
char *qz_err_buf();
int qz_fail(int code, const char *fmt, ...);
int qz_check_launch(const char *what);

#define QZ_REQUIRE_PTR(p)                                                  \
    do {                                                                   \
        if ((p) == nullptr) return qz_fail(QZ_E_NULL, "%s: %s is NULL", __func__, #p); \
    } while (0)
#define QZ_REQUIRE_ALIGN(p, a)                                             \
    do {                                                                   \
        if (((uintptr_t)(p)) % (a)) return qz_fail(QZ_E_ALIGN, "%s: %s not %d-byte aligned", __func__, #p, (int)(a)); \
    } while (0)
#define QZ_REQUIRE(cond)                                                   \
    do {                                                                   \
        if (!(cond)) return qz_fail(QZ_E_RANGE, "%s: requirement failed: %s", __func__, #cond); \
    } while (0)

static inline unsigned qz_blocks_for(int64_t work_items, int per_block) {
    return (unsigned)((work_items + per_block - 1) / per_block);
}

// 24-byte state load/store through 64-bit accesses
__device__ __forceinline__ QzState qz_load_state(const qz_state *p) {
    const uint64_t *q = reinterpret_cast<const uint64_t *>(p);
    QzState s;
    s.H = __ldg(q); s.V = __ldg(q + 1); s.meta = __ldg(q + 2);
    return s;
}
__device__ __forceinline__ void qz_store_state(qz_state *p, const QzState &s) {
    uint64_t *q = reinterpret_cast<uint64_t *>(p);
    q[0] = s.H; q[1] = s.V; q[2] = s.meta;
}
