// Shared device/host helpers for the rerevst_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/rerevst_b200.h"

namespace rrv {

// ---- error / launch bookkeeping (capi.cu) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> message; counts the launch
extern int g_lo_fp16;                  // rrv_set_lo_format

#define RRV_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            rrv::set_error(__VA_ARGS__);       \
            return 1;                          \
        }                                      \
    } while (0)

// ---- hi/lo split of an fp32 value into two 16-bit pieces ----
// hi = bf16_rn(v); lo = T_rn(v - float(hi)), T = bf16 or fp16.
__device__ __forceinline__ uint16_t bf16_bits(float v) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float bf16_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ uint16_t lo_bits(float r, int lo_fp16) {
    return lo_fp16 ? __half_as_ushort(__float2half_rn(r)) : bf16_bits(r);
}
__device__ __forceinline__ float lo_to_f32(uint16_t b, int lo_fp16) {
    return lo_fp16 ? __half2float(__ushort_as_half(b)) : bf16_to_f32(b);
}
__device__ __forceinline__ void split_hi_lo(float v, int lo_fp16, uint16_t& hi, uint16_t& lo) {
    hi = bf16_bits(v);
    lo = lo_bits(v - bf16_to_f32(hi), lo_fp16);
}

// 8 consecutive channels of a planes tensor -> 8 floats (hi + lo).
__device__ __forceinline__ void load8(const uint16_t* hi, const uint16_t* lo, int lo_fp16, float* v) {
    uint4 h = *reinterpret_cast<const uint4*>(hi);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = bf16_to_f32((uint16_t)(hw[i] & 0xffffu));
        v[2 * i + 1] = bf16_to_f32((uint16_t)(hw[i] >> 16));
    }
    if (lo != nullptr) {
        uint4 l = *reinterpret_cast<const uint4*>(lo);
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] += lo_to_f32((uint16_t)(lw[i] & 0xffffu), lo_fp16);
            v[2 * i + 1] += lo_to_f32((uint16_t)(lw[i] >> 16), lo_fp16);
        }
    }
}

__device__ __forceinline__ void store8(uint16_t* hi, uint16_t* lo, int lo_fp16, const float* v) {
    uint32_t hw[4], lw[4];
    if (!lo_fp16) {                      // bf16 lo (the default): two values per conversion instruction
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
            const float r0 = v[2 * i] - __uint_as_float(hw[i] << 16);
            const float r1 = v[2 * i + 1] - __uint_as_float(hw[i] & 0xffff0000u);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
            lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (lo != nullptr) *reinterpret_cast<uint4*>(lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint16_t h0, l0, h1, l1;
        split_hi_lo(v[2 * i], lo_fp16, h0, l0);
        split_hi_lo(v[2 * i + 1], lo_fp16, h1, l1);
        hw[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        lw[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (lo != nullptr) *reinterpret_cast<uint4*>(lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// ---- fused epilogue (see rrv_epilogue in the public header) ----
struct EpiDev {
    const float* bias;
    const float* norm1;
    const uint16_t* res_hi;
    const uint16_t* res_lo;
    const float* norm2;
    const float* affine;
    long long res_batch_stride;
    int act, res_shift, res_H, res_W, C, lo_fp16, res_f32;
};

inline EpiDev make_epi(const rrv_epilogue& e, int C) {
    EpiDev d;
    d.bias = e.bias;
    d.norm1 = e.norm1;
    d.res_hi = (const uint16_t*)e.res_hi;
    d.res_lo = (const uint16_t*)e.res_lo;
    d.norm2 = e.norm2;
    d.affine = e.affine;
    d.res_batch_stride = e.res_batch_stride;
    d.act = e.act;
    d.res_shift = e.res_shift;
    d.res_H = e.res_H;
    d.res_W = e.res_W;
    d.C = C;
    d.lo_fp16 = g_lo_fp16;
    d.res_f32 = e.res_f32;
    return d;
}

__device__ __forceinline__ float saved_norm(float v, const float* tab, int C, int c) {
    // InstanceNorm.forward: (x - mean) * rstd, max(lo, .), min(hi, .)
    v = (v - __ldg(tab + c)) * __ldg(tab + C + c);
    v = fmaxf(__ldg(tab + 2 * C + c), v);
    v = fminf(__ldg(tab + 3 * C + c), v);
    return v;
}

// NV channels starting at c0 (NV = 8 normally), pixel (n, y, x).
template <int NV>
__device__ __forceinline__ void apply_epilogue(const EpiDev& e, float* v, int n, int y, int x, int c0) {
    if (e.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __ldg(e.bias + c0 + i);
    }
    if (e.act == 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (e.act == 2) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = v[i] > 0.0f ? v[i] : 0.2f * v[i];
    }
    if (e.norm1 != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = saved_norm(v[i], e.norm1, e.C, c0 + i);
    }
    if (e.res_hi != nullptr) {
        const long long off = (long long)n * e.res_batch_stride +
                              ((long long)(y >> e.res_shift) * e.res_W + (x >> e.res_shift)) * e.C + c0;
        if (e.res_f32) {
            const float* rf = reinterpret_cast<const float*>(e.res_hi) + off;
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] += __ldg(rf + i);
        } else if (NV == 8) {
            float r[8];
            load8(e.res_hi + off, e.res_lo ? e.res_lo + off : nullptr, e.lo_fp16, r);
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] += r[i];
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float r = bf16_to_f32(e.res_hi[off + i]);
                if (e.res_lo) r += lo_to_f32(e.res_lo[off + i], e.lo_fp16);
                v[i] += r;
            }
        }
    }
    if (e.norm2 != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = saved_norm(v[i], e.norm2, e.C, c0 + i);
    }
    if (e.affine != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = v[i] * __ldg(e.affine + c0 + i) + __ldg(e.affine + e.C + c0 + i);
    }
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- double min / max atomics (statistics partials) ----
__device__ __forceinline__ void atomic_min_double(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_double(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

}  // namespace rrv
