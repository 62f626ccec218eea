// Per-channel statistics of the pre-pass ("Sequence-Level Global Feature Sharing").
//
// InstanceNorm.compute (style_network_global.py:59-77) takes mean / biased variance / min / max
// of the normalised tensor over batch and space in two passes; EncoderStyle.cal_mean_std
// (:304-315) takes mean and unbiased std; FilterPredictor.compute (:161-172) takes a spatial
// and batch mean followed by a 64 -> 1024 linear layer.  Here one two-pass reduction produces
// mergeable partials {count, sum, M2, min, max} in double precision (per-thread accumulation over a
// strip of pixels, shared memory across the threads of a block, one atomic per channel per block), so that ranks of the
// frame-parallel driver can all-gather and merge them (Chan et al.) in a fixed order.
#include <math_constants.h>

#include "rrv_common.cuh"

namespace rrv {

constexpr int ST_PIX = 2048;   // pixels per block

__global__ void stats_init_kernel(double* part, int C, double count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    part[c] = count;
    part[C + c] = 0.0;
    part[2 * C + c] = 0.0;
    part[3 * C + c] = CUDART_INF;
    part[4 * C + c] = -CUDART_INF;
}

// PASS 0: sum.  PASS 1: M2 about sum/count, min, max.
template <int PASS>
__global__ void __launch_bounds__(256) stats_kernel(const float* __restrict__ x, long long npix, int C, double* part) {
    __shared__ double s_a[256];
    __shared__ float s_mn[256], s_mx[256];
    const int cpb = C < 256 ? C : 256;          // channels covered by one block
    const int ppb = 256 / cpb;                  // pixel lanes per block
    const int t = threadIdx.x;
    const int cl = t % cpb, pl = t / cpb;
    const int c = blockIdx.y * 256 + cl;
    const bool active = pl < ppb && c < C;
    const long long p0 = (long long)blockIdx.x * ST_PIX;
    const long long p1 = p0 + ST_PIX < npix ? p0 + ST_PIX : npix;
    double acc = 0.0;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    if (active) {
        if (PASS == 0) {
            for (long long p = p0 + pl; p < p1; p += ppb) acc += (double)x[p * C + c];
        } else {
            const float mean = (float)(part[C + c] / part[c]);
            for (long long p = p0 + pl; p < p1; p += ppb) {
                const float v = x[p * C + c];
                const float d = v - mean;          // fp32 like the reference's x - saved_mean
                acc += (double)d * (double)d;
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
        }
    }
    s_a[t] = acc; s_mn[t] = mn; s_mx[t] = mx;
    __syncthreads();
    if (pl == 0 && c < C) {
        for (int k = 1; k < ppb; ++k) {
            acc += s_a[k * cpb + cl];
            mn = fminf(mn, s_mn[k * cpb + cl]);
            mx = fmaxf(mx, s_mx[k * cpb + cl]);
        }
        if (PASS == 0) {
            atomicAdd(part + C + c, acc);
        } else {
            atomicAdd(part + 2 * C + c, acc);
            atomic_min_double(part + 3 * C + c, (double)mn);
            atomic_max_double(part + 4 * C + c, (double)mx);
        }
    }
}

int channel_stats(const float* x, long long npix, int C, double* part, cudaStream_t st) {
    RRV_REQUIRE(x && part, "rrv_channel_stats: NULL tensor");
    RRV_REQUIRE(npix > 0 && C > 0, "rrv_channel_stats: empty tensor (npix=%lld C=%d)", npix, C);
    RRV_REQUIRE(C <= 256 ? (256 % C == 0) : (C % 256 == 0), "rrv_channel_stats: C=%d must divide or be a multiple of 256", C);
    stats_init_kernel<<<ceil_div(C, 256), 256, 0, st>>>(part, C, (double)npix);
    if (check_launch("stats_init_kernel")) return 1;
    dim3 grid(ceil_div(npix, ST_PIX), ceil_div(C, 256));
    stats_kernel<0><<<grid, 256, 0, st>>>(x, npix, C, part);
    if (check_launch("stats_kernel<0>")) return 1;
    stats_kernel<1><<<grid, 256, 0, st>>>(x, npix, C, part);
    return check_launch("stats_kernel<1>");
}

int stats_init(double* part, int C, double count, cudaStream_t st) {
    RRV_REQUIRE(part && C > 0, "rrv_stats_init: bad arguments");
    stats_init_kernel<<<ceil_div(C, 256), 256, 0, st>>>(part, C, count);
    return check_launch("stats_init_kernel");
}

// Partials accumulated in ONE pass by a producing kernel (rrv_conv.stats, rrv_pointwise's stats) hold sum(x^2) in row 2;
// M2 about the mean = sum(x^2) - sum(x)^2 / n, in double (the inputs are fp32 values: ~2^-29 relative on M2 for data whose
// mean and spread are comparable, as every tensor on this path is).
__global__ void stats_sums_to_m2_kernel(double* part, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double n = part[c], s = part[C + c], q = part[2 * C + c];
    const double m2 = q - s * s / n;
    part[2 * C + c] = m2 > 0.0 ? m2 : 0.0;
}

int stats_sums_to_m2(double* part, int C, cudaStream_t st) {
    RRV_REQUIRE(part && C > 0, "rrv_stats_sums_to_m2: bad arguments");
    stats_sums_to_m2_kernel<<<ceil_div(C, 128), 128, 0, st>>>(part, C);
    return check_launch("stats_sums_to_m2_kernel");
}

// Chan/Golub/LeVeque pairwise merge, applied left to right over the parts (deterministic).
__global__ void stats_merge_kernel(const double* __restrict__ parts, int nparts, int C, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double n = 0.0, sum = 0.0, m2 = 0.0, mn = CUDART_INF, mx = -CUDART_INF;
    for (int k = 0; k < nparts; ++k) {
        const double* p = parts + (long long)k * 5 * C;
        const double nb = p[c];
        if (nb <= 0.0) continue;
        const double sb = p[C + c];
        if (n == 0.0) {
            n = nb; sum = sb; m2 = p[2 * C + c];
        } else {
            const double delta = sb / nb - sum / n;
            m2 = m2 + p[2 * C + c] + delta * delta * n * nb / (n + nb);
            n += nb; sum += sb;
        }
        mn = fmin(mn, p[3 * C + c]);
        mx = fmax(mx, p[4 * C + c]);
    }
    out[c] = n; out[C + c] = sum; out[2 * C + c] = m2; out[3 * C + c] = mn; out[4 * C + c] = mx;
}

int stats_merge(const double* parts, int nparts, int C, double* merged, cudaStream_t st) {
    RRV_REQUIRE(parts && merged && nparts > 0, "rrv_stats_merge: bad arguments");
    stats_merge_kernel<<<ceil_div(C, 128), 128, 0, st>>>(parts, nparts, C, merged);
    return check_launch("stats_merge_kernel");
}

// from_sums: row 2 of the partial still holds sum(x^2) (a one-pass producer, see rrv_stats_sums_to_m2): M2 is formed here.
__global__ void stats_finalize_kernel(const double* __restrict__ part, int C, int kind, float eps, float* __restrict__ out, int from_sums) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double n = part[c];
    const float mean = (float)(part[C + c] / n);
    double m2 = part[2 * C + c];
    if (from_sums) {
        m2 -= part[C + c] * part[C + c] / n;
        m2 = m2 > 0.0 ? m2 : 0.0;
    }
    if (kind == 2) { out[c] = mean; return; }
    if (kind == 1) {
        // EncoderStyle.cal_mean_std: sqrt(var_unbiased + eps), mean  -> AdaIN {scale, shift}
        const float var = (float)(m2 / (n > 1.0 ? n - 1.0 : 1.0));
        out[c] = sqrtf(var + eps);
        out[C + c] = mean;
        return;
    }
    // InstanceNorm.compute: rsqrt(mean((x-mean)^2) + eps); x_min / x_max of the normalised tensor.
    // fp32 subtraction and multiplication are monotone, so max((x-m)*r) == (max(x)-m)*r bit for bit.
    const float var = (float)(m2 / n);
    const float rstd = 1.0f / sqrtf(var + eps);
    out[c] = mean;
    out[C + c] = rstd;
    if (kind == 0) {
        out[2 * C + c] = ((float)part[3 * C + c] - mean) * rstd;
        out[3 * C + c] = ((float)part[4 * C + c] - mean) * rstd;
    } else {
        out[2 * C + c] = -CUDART_INF_F;
        out[3 * C + c] = CUDART_INF_F;
    }
}

int stats_finalize(const double* part, int C, int kind, float eps, float* out, cudaStream_t st) {
    RRV_REQUIRE(part && out, "rrv_stats_finalize: NULL tensor");
    const int from_sums = (kind & 16) ? 1 : 0;
    kind &= 15;
    RRV_REQUIRE(kind >= 0 && kind <= 3, "rrv_stats_finalize: bad kind %d", kind);
    stats_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(part, C, kind, eps, out, from_sums);
    return check_launch("stats_finalize_kernel");
}

// FilterPredictor FC: out[j] = b[j] + sum_k W[j][k] * cat(c, s)[k]   (style_network_global.py:169)
__global__ void filter_fc_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ c_mean,
                                 const float* __restrict__ s_mean, float* __restrict__ out) {
    __shared__ float s_in[64];
    if (threadIdx.x < 32) s_in[threadIdx.x] = c_mean[threadIdx.x];
    else if (threadIdx.x < 64) s_in[threadIdx.x] = s_mean[threadIdx.x - 32];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 1024) return;
    // the row's 64 weights as 16 independent 16-byte loads (all in flight at once), then the same sequential FMA order as before
    float4 wr[16];
    const float4* wj = reinterpret_cast<const float4*>(w + (size_t)j * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) wr[q] = __ldg(wj + q);
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        acc = fmaf(wr[q].x, s_in[4 * q], acc);
        acc = fmaf(wr[q].y, s_in[4 * q + 1], acc);
        acc = fmaf(wr[q].z, s_in[4 * q + 2], acc);
        acc = fmaf(wr[q].w, s_in[4 * q + 3], acc);
    }
    out[j] = acc + b[j];
}

int filter_fc(const float* w, const float* b, const float* c_mean, const float* s_mean, float* out, cudaStream_t st) {
    RRV_REQUIRE(w && b && c_mean && s_mean && out, "rrv_filter_fc: NULL tensor");
    filter_fc_kernel<<<16, 64, 0, st>>>(w, b, c_mean, s_mean, out);
    return check_launch("filter_fc_kernel");
}

}  // namespace rrv

// ---- spatial mean of a 3x3 convolution without computing the convolution ------------------------------------------------
// FilterPredictor (style_network_global.py:150-172, style_network_frame.py:53-62) only ever uses
// mean_{n,y,x}(conv3x3(content) + b).  The convolution is linear, so the sum over all output pixels of tap (dy, dx) is the sum
// of the input over the pixels that tap reads: everything except the last / first row (dy = 0 / 2) and the last / first column
// (dx = 0 / 2) -- zero padding contributes nothing.  Nine per-channel sums of the input (total, first / last row, first / last
// column, four corners) give all nine tap sums by inclusion-exclusion, and a 64 x 4608 dot product finishes: one read of the
// input instead of a 512 -> 64 convolution per predictor pair (0.07 ms each at 152 x 256, three per frame in frame mode).
namespace rrv {

// out: double[9][C] = {total, row 0, row H-1, col 0, col W-1, (0,0), (0,W-1), (H-1,0), (H-1,W-1)}, zero-initialised by the caller
__global__ void __launch_bounds__(256, 3) border_sums_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int N, int H,
                                                          int W, int C, double* __restrict__ out) {
    const int C8 = C >> 3;
    const unsigned gtid = blockIdx.x * 256u + threadIdx.x, gthreads = gridDim.x * 256u;
    const int c0 = (int)(gtid % (unsigned)C8) * 8;
    const unsigned pstride = gthreads / (unsigned)C8;
    const unsigned HW = (unsigned)H * (unsigned)W, npix = (unsigned)N * HW;
    float acc[5][8];
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[q][k] = 0.0f;
    for (unsigned pp = gtid / (unsigned)C8; pp < npix; pp += 2 * pstride) {
      float vv[2][8];
      const bool two = pp + pstride < npix;
      load8(hi + (size_t)pp * C + c0, lo ? lo + (size_t)pp * C + c0 : nullptr, 0, vv[0]);       // both pixels' loads first
      if (two) load8(hi + (size_t)(pp + pstride) * C + c0, lo ? lo + (size_t)(pp + pstride) * C + c0 : nullptr, 0, vv[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const unsigned p = pp + (unsigned)u * pstride;
        const float* v = vv[u];
        const unsigned rem = p % HW;
        const unsigned y = rem / (unsigned)W, x = rem - y * (unsigned)W;
        const bool r0 = y == 0, rl = y == (unsigned)H - 1, q0 = x == 0, ql = x == (unsigned)W - 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            acc[0][k] += v[k];
            if (r0) acc[1][k] += v[k];
            if (rl) acc[2][k] += v[k];
            if (q0) acc[3][k] += v[k];
            if (ql) acc[4][k] += v[k];
        }
        if ((r0 || rl) && (q0 || ql)) {          // a corner pixel (four per image): straight to memory
            const int corner = (rl ? 2 : 0) + (ql ? 1 : 0);
            if (r0 && rl) {                      // H == 1: the pixel is both a first-row and a last-row corner
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(out + (size_t)(5 + (ql ? 1 : 0)) * C + c0 + k, (double)v[k]);
            }
            if (q0 && ql) {                      // W == 1
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(out + (size_t)(5 + (rl ? 2 : 0)) * C + c0 + k, (double)v[k]);
            }
            if (r0 && rl && q0 && ql) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(out + (size_t)5 * C + c0 + k, (double)v[k]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(out + (size_t)(5 + corner) * C + c0 + k, (double)v[k]);
        }
      }
    }
    __shared__ float s_red[256][8 + 1];
    const int t = threadIdx.x;
    const int groups = C8 < 256 ? C8 : 256;
    for (int q = 0; q < 5; ++q) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[t][k] = acc[q][k];
        __syncthreads();
        for (int j = t; j < groups * 8; j += 256) {
            const int g = j >> 3, k = j & 7;
            double a = 0.0;
            for (int u = g; u < 256; u += groups) a += (double)s_red[u][k];
            const int c = (int)((blockIdx.x * 256u + (unsigned)g) % (unsigned)C8) * 8 + k;
            atomicAdd(out + (size_t)q * C + c, a);
        }
    }
}

// part = double[5][Cout] = {count, sum over all output pixels of conv(x) + b, 0, 0, 0}: a partial like rrv_channel_stats'
// (count and sum only), so that the ranks of a sharded pre-pass can merge it.
__global__ void __launch_bounds__(256) conv_mean_finish_kernel(const double* __restrict__ sums, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int Cin, int Cout, double count,
                                                               double* __restrict__ part) {
    // one block per output channel: every thread forms the tap sums of its own elements of the 9 Cin-long dot product (no shared
    // table, no serial per-warp loop: 8 blocks x 8 warps x 144 double FMAs per lane took 39 us, three times per frame in frame mode)
    const int o = blockIdx.x;
    const float* wo = w + (size_t)o * Cin * 9;
    double acc = 0.0;
    for (int i = threadIdx.x; i < Cin * 9; i += 256) {
        const int c = i / 9, t = i - c * 9, dy = t / 3, dx = t - dy * 3;
        // pixels NOT read by tap (dy, dx): last row for dy = 0, first row for dy = 2; last / first column for dx = 0 / 2
        const int row = dy == 0 ? 2 : (dy == 2 ? 1 : -1);          // index into sums of the excluded row (row H-1 / row 0)
        const int col = dx == 0 ? 4 : (dx == 2 ? 3 : -1);          // excluded column (col W-1 / col 0)
        double s = sums[c];
        if (row >= 0) s -= sums[(size_t)row * Cin + c];
        if (col >= 0) s -= sums[(size_t)col * Cin + c];
        if (row >= 0 && col >= 0) s += sums[(size_t)(5 + (row == 2 ? 2 : 0) + (col == 4 ? 1 : 0)) * Cin + c];
        acc += (double)__ldg(wo + i) * s;
    }
    __shared__ double s_red[8];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += s_red[k];
        part[o] = count;
        part[Cout + o] = a + (bias ? (double)bias[o] * count : 0.0);
        part[2 * Cout + o] = 0.0;
        part[3 * Cout + o] = 0.0;
        part[4 * Cout + o] = 0.0;
    }
}

int conv3x3_output_sum(const void* in_hi, const void* in_lo, int N, int H, int W, int Cin, const float* w_oihw, const float* bias,
                       int Cout, double* scratch, double* part, cudaStream_t st) {
    RRV_REQUIRE(in_hi && w_oihw && scratch && part, "rrv_conv3x3_output_sum: NULL tensor");
    RRV_REQUIRE(N > 0 && H > 0 && W > 0 && Cin % 8 == 0 && Cin / 8 <= 256 && 256 % (Cin / 8) == 0 && Cout > 0,
                "rrv_conv3x3_output_sum: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d", N, H, W, Cin, Cout);
    RRV_REQUIRE((long long)N * H * W < (1LL << 31), "rrv_conv3x3_output_sum: more than 2^31 pixels");
    cudaMemsetAsync(scratch, 0, sizeof(double) * 9 * Cin, st);
    const long long total = (long long)N * H * W * (Cin / 8);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 6);
    border_sums_kernel<<<grid, 256, 0, st>>>((const uint16_t*)in_hi, (const uint16_t*)in_lo, N, H, W, Cin, scratch);
    if (check_launch("border_sums_kernel")) return 1;
    conv_mean_finish_kernel<<<Cout, 256, 0, st>>>(scratch, w_oihw, bias, Cin, Cout, (double)N * H * W, part);
    return check_launch("conv_mean_finish_kernel");
}

}  // namespace rrv
