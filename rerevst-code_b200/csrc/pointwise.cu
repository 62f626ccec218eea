// HBM-bound pointwise kernels: 2x2 max-pool on planes, the epilogue chain alone (pre-pass
// normalisation), layout conversions, the de-normalise/clamp/BGR postprocess and the fp32
// weight repack.  All vectorised to 16-byte accesses along the channel dimension.
#include <math_constants.h>

#include "rrv_common.cuh"

namespace rrv {

// ---- nn.MaxPool2d(2,2): vgg19.features[4], [9], [18] (style_network_global.py:275-278) ----
__global__ void __launch_bounds__(256) maxpool_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo,
                                                      uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                                      int N, int H, int W, int C, int lo_fp16) {
    const int Ho = H >> 1, Wo = W >> 1, C8 = C >> 3;
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c8 = (int)(i % C8);
        long long t = i / C8;
        const int xo = (int)(t % Wo); t /= Wo;
        const int yo = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float best[8];
        uint32_t bh[8], bl[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long off = (((long long)n * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C + c8 * 8;
            const uint4 h = *reinterpret_cast<const uint4*>(in_hi + off);
            uint4 l = make_uint4(0, 0, 0, 0);
            if (in_lo) l = *reinterpret_cast<const uint4*>(in_lo + off);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint16_t hb = (uint16_t)(hw[j >> 1] >> ((j & 1) * 16));
                const uint16_t lb = (uint16_t)(lw[j >> 1] >> ((j & 1) * 16));
                const float v = bf16_to_f32(hb) + (in_lo ? lo_to_f32(lb, lo_fp16) : 0.0f);
                if (k == 0 || v > best[j]) { best[j] = v; bh[j] = hb; bl[j] = lb; }
            }
        }
        const long long o = (((long long)n * Ho + yo) * Wo + xo) * C + c8 * 8;
        *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(bh[0] | (bh[1] << 16), bh[2] | (bh[3] << 16),
                                                           bh[4] | (bh[5] << 16), bh[6] | (bh[7] << 16));
        if (out_lo)
            *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(bl[0] | (bl[1] << 16), bl[2] | (bl[3] << 16),
                                                               bl[4] | (bl[5] << 16), bl[6] | (bl[7] << 16));
    }
}

int maxpool2x2(const void* in_hi, const void* in_lo, int N, int H, int W, int C, void* out_hi, void* out_lo,
               cudaStream_t st) {
    RRV_REQUIRE(in_hi && out_hi, "rrv_maxpool2x2: NULL tensor");
    RRV_REQUIRE(C % 8 == 0, "rrv_maxpool2x2: C must be a multiple of 8 (got %d)", C);
    RRV_REQUIRE((in_lo == nullptr) == (out_lo == nullptr), "rrv_maxpool2x2: lo planes must both be set or both NULL");
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    maxpool_kernel<<<grid, 256, 0, st>>>((const uint16_t*)in_hi, (const uint16_t*)in_lo, (uint16_t*)out_hi,
                                         (uint16_t*)out_lo, N, H, W, C, g_lo_fp16);
    return check_launch("maxpool_kernel");
}

// ---- epilogue chain alone over fp32 NHWC ----
// A thread keeps ONE 8-channel group for the whole kernel (the grid stride is a multiple of C / 8) and walks over pixels: the
// per-channel constants of the chain are loaded into registers once, the index math is 32-bit, and per pixel there are two 16-byte
// loads (+ the residual), ~10 FP32 operations per value, the hi / lo split and two 16-byte stores -- an HBM-bound pass.
// STATS: the per-channel {sum, sum of squares[, min, max]} of the values written also go to `stats` (double[5][C]): fp32 per thread
// (a few dozen pixels), then double across the block's threads of equal group through shared memory, one atomic per channel.
// KIND fixes the set of stages at compile time so that only their constants occupy registers (occupancy is what an HBM-bound
// pass lives on): 0 = norm1; 1 = norm1 + affine (AdaIN); 2 = norm1 + residual; 3 = residual only; 4 = whatever `ep` says.
template <bool STATS, int KIND>
__global__ void __launch_bounds__(256, (KIND == 4 || STATS) ? 2 : 3) pointwise_kernel(const float* __restrict__ in, long long in_bs, int N, int H, int W,
                                                        int C, EpiDev ep, int out_mode, uint16_t* out_hi,
                                                        uint16_t* out_lo, float* out_f32, double* stats, int stats_minmax) {
    constexpr int V = 4;                      // channels per thread: 16-byte loads, and the constants below fit ~90 registers
    const int CV = C / V;
    const unsigned gtid = blockIdx.x * 256u + threadIdx.x, gthreads = gridDim.x * 256u;
    const int c0 = (int)(gtid % (unsigned)CV) * V;
    const unsigned pstride = gthreads / (unsigned)CV;
    const unsigned HW = (unsigned)H * (unsigned)W, npix = (unsigned)N * HW;
    float ssum[V], ssq[V], smn[V], smx[V];
    if (STATS) {
#pragma unroll
        for (int k = 0; k < V; ++k) { ssum[k] = 0.0f; ssq[k] = 0.0f; smn[k] = CUDART_INF_F; smx[k] = -CUDART_INF_F; }
    }
    // which stages exist: compile-time for KIND < 4 (the host guarantees the match), from `ep` otherwise
    const bool has_bias = KIND == 4 && ep.bias != nullptr, has_act = KIND == 4 && ep.act != 0;
    const bool has_n1 = KIND <= 2 || (KIND == 4 && ep.norm1 != nullptr);
    const bool has_res = KIND == 2 || KIND == 3 || (KIND == 4 && ep.res_hi != nullptr);
    const bool has_n2 = KIND == 4 && ep.norm2 != nullptr;
    const bool has_aff = KIND == 1 || (KIND == 4 && ep.affine != nullptr);
    // per-channel constants, once per thread
    float bias[V], m1[V], r1[V], lo1[V], hi1[V], m2[V], r2[V], lo2[V], hi2[V], sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = c0 + k;
        bias[k] = has_bias ? __ldg(ep.bias + c) : 0.0f;
        m1[k] = has_n1 ? __ldg(ep.norm1 + c) : 0.0f;
        r1[k] = has_n1 ? __ldg(ep.norm1 + C + c) : 1.0f;
        lo1[k] = has_n1 ? __ldg(ep.norm1 + 2 * C + c) : -CUDART_INF_F;
        hi1[k] = has_n1 ? __ldg(ep.norm1 + 3 * C + c) : CUDART_INF_F;
        m2[k] = has_n2 ? __ldg(ep.norm2 + c) : 0.0f;
        r2[k] = has_n2 ? __ldg(ep.norm2 + C + c) : 1.0f;
        lo2[k] = has_n2 ? __ldg(ep.norm2 + 2 * C + c) : -CUDART_INF_F;
        hi2[k] = has_n2 ? __ldg(ep.norm2 + 3 * C + c) : CUDART_INF_F;
        sc[k] = has_aff ? __ldg(ep.affine + c) : 1.0f;
        sh[k] = has_aff ? __ldg(ep.affine + C + c) : 0.0f;
    }
    // U pixels per trip: all their loads are issued before the first value is used (one 16-byte load per pixel and tensor would
    // leave too few bytes in flight per SM for an HBM-bound pass)
    constexpr int U = 4;
    for (unsigned p0 = gtid / (unsigned)CV; p0 < npix; p0 += U * pstride) {
        float4 a[U], q[U];
        uint2 qh[U], ql[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned p = p0 + (unsigned)u * pstride;
            ok[u] = p < npix;
            if (!ok[u]) continue;
            unsigned n = 0, rem = p;
            if (N > 1) { n = p / HW; rem = p - n * HW; }
            a[u] = __ldg(reinterpret_cast<const float4*>(in + (long long)n * in_bs + (long long)rem * C + c0));
            if (has_res) {
                const unsigned y = rem / (unsigned)W, x = rem - y * (unsigned)W;
                const long long off = (long long)n * ep.res_batch_stride + ((long long)(y >> ep.res_shift) * ep.res_W + (x >> ep.res_shift)) * C + c0;
                if (ep.res_f32) {
                    q[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.res_hi) + off));
                } else {
                    qh[u] = __ldg(reinterpret_cast<const uint2*>(ep.res_hi + off));
                    if (ep.res_lo != nullptr) ql[u] = __ldg(reinterpret_cast<const uint2*>(ep.res_lo + off));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const unsigned p = p0 + (unsigned)u * pstride;
            float v[V] = {a[u].x, a[u].y, a[u].z, a[u].w};
            float rr[V];
            if (has_res) {
                if (ep.res_f32) {
                    rr[0] = q[u].x; rr[1] = q[u].y; rr[2] = q[u].z; rr[3] = q[u].w;
                } else {
                    rr[0] = __uint_as_float(qh[u].x << 16); rr[1] = __uint_as_float(qh[u].x & 0xffff0000u);
                    rr[2] = __uint_as_float(qh[u].y << 16); rr[3] = __uint_as_float(qh[u].y & 0xffff0000u);
                    if (ep.res_lo != nullptr) {
                        rr[0] += lo_to_f32((uint16_t)(ql[u].x & 0xffffu), ep.lo_fp16); rr[1] += lo_to_f32((uint16_t)(ql[u].x >> 16), ep.lo_fp16);
                        rr[2] += lo_to_f32((uint16_t)(ql[u].y & 0xffffu), ep.lo_fp16); rr[3] += lo_to_f32((uint16_t)(ql[u].y >> 16), ep.lo_fp16);
                    }
                }
            }
            // the chain of rrv_epilogue, in the reference's operation order (rrv_common.cuh: apply_epilogue)
            if (has_bias) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] += bias[k];
            }
            if (!has_act) {
            } else if (ep.act == 1) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = fmaxf(v[k], 0.0f);
            } else if (ep.act == 2) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = v[k] > 0.0f ? v[k] : 0.2f * v[k];
            }
            if (has_n1) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = fminf(hi1[k], fmaxf(lo1[k], (v[k] - m1[k]) * r1[k]));
            }
            if (has_res) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] += rr[k];
            }
            if (has_n2) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = fminf(hi2[k], fmaxf(lo2[k], (v[k] - m2[k]) * r2[k]));
            }
            if (has_aff) {
#pragma unroll
                for (int k = 0; k < V; ++k) v[k] = v[k] * sc[k] + sh[k];
            }
            if (STATS) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    ssum[k] += v[k];
                    ssq[k] = fmaf(v[k], v[k], ssq[k]);
                    smn[k] = fminf(smn[k], v[k]);
                    smx[k] = fmaxf(smx[k], v[k]);
                }
            }
            const long long o = (long long)p * C + c0;
            if (out_mode == RRV_OUT_PLANES) {
                uint32_t hw[2], lw[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    uint16_t h0, l0, h1, l1;
                    split_hi_lo(v[2 * k], ep.lo_fp16, h0, l0);
                    split_hi_lo(v[2 * k + 1], ep.lo_fp16, h1, l1);
                    hw[k] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    lw[k] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(hw[0], hw[1]);
                if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(lw[0], lw[1]);
            } else {
                *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    if (STATS) {
        // threads t, t + CV, t + 2 CV, ... of the block own the same channel group (256 % CV == 0)
        __shared__ float s_red[4][256][V + 1];
        const int t = threadIdx.x;
#pragma unroll
        for (int k = 0; k < V; ++k) { s_red[0][t][k] = ssum[k]; s_red[1][t][k] = ssq[k]; s_red[2][t][k] = smn[k]; s_red[3][t][k] = smx[k]; }
        __syncthreads();
        const int groups = CV < 256 ? CV : 256;
        for (int j = t; j < groups * V; j += 256) {
            const int g = j / V, k = j % V;
            double sa = 0.0, sb = 0.0;
            float mn = CUDART_INF_F, mx = -CUDART_INF_F;
            for (int u = g; u < 256; u += groups) {
                sa += (double)s_red[0][u][k];
                sb += (double)s_red[1][u][k];
                mn = fminf(mn, s_red[2][u][k]);
                mx = fmaxf(mx, s_red[3][u][k]);
            }
            const int c = (int)((blockIdx.x * 256u + (unsigned)g) % (unsigned)CV) * V + k;      // the group thread g of this block owns
            atomicAdd(stats + C + c, sa);
            atomicAdd(stats + 2 * C + c, sb);
            if (stats_minmax) {
                atomic_min_double(stats + 3 * C + c, (double)mn);
                atomic_max_double(stats + 4 * C + c, (double)mx);
            }
        }
    }
}

int pointwise(const float* in, long long in_bs, int N, int H, int W, int C, const rrv_epilogue* ep, int out_mode,
              void* out_hi, void* out_lo, float* out_f32, double* stats, int stats_minmax, cudaStream_t st) {
    RRV_REQUIRE(in && ep, "rrv_pointwise: NULL input");
    RRV_REQUIRE(C % 8 == 0, "rrv_pointwise: C must be a multiple of 8 (got %d)", C);
    RRV_REQUIRE(out_mode == RRV_OUT_PLANES || out_mode == RRV_OUT_F32_NHWC, "rrv_pointwise: bad out_mode %d", out_mode);
    RRV_REQUIRE(out_mode == RRV_OUT_PLANES ? out_hi != nullptr : out_f32 != nullptr, "rrv_pointwise: NULL output");
    const long long total = (long long)N * H * W * (C / 8);
    if (total == 0) return 0;
    RRV_REQUIRE((long long)N * H * W < (1LL << 31), "rrv_pointwise: more than 2^31 pixels");
    RRV_REQUIRE(C / 4 <= 256 && 256 % (C / 4) == 0, "rrv_pointwise: C / 4 must divide 256 (C=%d)", C);     // a thread keeps its channel group
    const int grid = (int)std::min<long long>((2 * total + 255) / 256, 148LL * 16);
    const bool plain = !ep->bias && ep->act == 0 && !ep->norm2;
    int kind = 4;
    if (plain && ep->norm1 && !ep->res_hi && !ep->affine) kind = 0;
    else if (plain && ep->norm1 && !ep->res_hi && ep->affine) kind = 1;
    else if (plain && ep->norm1 && ep->res_hi && !ep->affine) kind = 2;
    else if (plain && !ep->norm1 && ep->res_hi && !ep->affine) kind = 3;
    const EpiDev e = make_epi(*ep, C);
#define RRV_PW(S, K)                                                                                                              \
    pointwise_kernel<S, K><<<grid, 256, 0, st>>>(in, in_bs, N, H, W, C, e, out_mode, (uint16_t*)out_hi, (uint16_t*)out_lo, out_f32, \
                                                 stats, stats_minmax)
    if (stats != nullptr) {
        switch (kind) {
            case 2: RRV_PW(true, 2); break;
            case 3: RRV_PW(true, 3); break;
            default: RRV_PW(true, 4); break;
        }
        return check_launch("pointwise_kernel<stats>");
    }
    switch (kind) {
        case 0: RRV_PW(false, 0); break;
        case 1: RRV_PW(false, 1); break;
        case 2: RRV_PW(false, 2); break;
        case 3: RRV_PW(false, 3); break;
        default: RRV_PW(false, 4); break;
    }
#undef RRV_PW
    return check_launch("pointwise_kernel");
}

// ---- layout conversions ----
__global__ void __launch_bounds__(256) planes_to_nchw_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                                                             int N, int H, int W, int C, float* __restrict__ out, int lo_fp16) {
    const long long total = (long long)N * C * H * W;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % W);
        long long t = i / W;
        const int y = (int)(t % H); t /= H;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        const long long s = (((long long)n * H + y) * W + x) * C + c;
        float v = bf16_to_f32(hi[s]);
        if (lo) v += lo_to_f32(lo[s], lo_fp16);
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256) nchw_to_planes_kernel(const float* __restrict__ in, int N, int H, int W, int C,
                                                             uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int lo_fp16) {
    const long long total = (long long)N * H * W * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        uint16_t h, l;
        split_hi_lo(in[(((long long)n * C + c) * H + y) * W + x], lo_fp16, h, l);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}

int planes_to_nchw(const void* hi, const void* lo, int N, int H, int W, int C, float* out, cudaStream_t st) {
    RRV_REQUIRE(hi && out, "rrv_planes_to_nchw: NULL tensor");
    const long long total = (long long)N * C * H * W;
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    planes_to_nchw_kernel<<<grid, 256, 0, st>>>((const uint16_t*)hi, (const uint16_t*)lo, N, H, W, C, out, g_lo_fp16);
    return check_launch("planes_to_nchw_kernel");
}

int nchw_to_planes(const float* in, int N, int H, int W, int C, void* hi, void* lo, cudaStream_t st) {
    RRV_REQUIRE(in && hi, "rrv_nchw_to_planes: NULL tensor");
    const long long total = (long long)N * C * H * W;
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    nchw_to_planes_kernel<<<grid, 256, 0, st>>>(in, N, H, W, C, (uint16_t*)hi, (uint16_t*)lo, g_lo_fp16);
    return check_launch("nchw_to_planes_kernel");
}

// ---- transform_back_image + tensor2numpy (test/framework.py:39-49) + crop (generate_real_video.py:167) ----
template <typename T>
__global__ void __launch_bounds__(256) postprocess_kernel(const float* __restrict__ in, int N, int H, int W, int y0, int x0,
                                                          int h, int w, T* __restrict__ out) {
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};
    const long long total = (long long)N * h * w;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % w);
        long long t = i / w;
        const int y = (int)(t % h);
        const int n = (int)(t / h);
        float bgr[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = in[(((long long)n * 3 + c) * H + y0 + y) * W + x0 + x];
            v = __fadd_rn(__fmul_rn(v, sd[c]), mean[c]);        // img * std + mean
            v = fminf(fmaxf(v, 0.0f), 1.0f);                    // clamp(0, 1)
            bgr[2 - c] = __fmul_rn(v, 255.0f);                  // * 255, RGB -> BGR
        }
        T* o = out + i * 3;
        if (sizeof(T) == 1) {            // cv2.imwrite's float32 -> uint8: saturate_cast<uchar>(cvRound(v)), ties to even
            o[0] = (T)__float2int_rn(bgr[0]); o[1] = (T)__float2int_rn(bgr[1]); o[2] = (T)__float2int_rn(bgr[2]);
        } else {
            o[0] = (T)bgr[0]; o[1] = (T)bgr[1]; o[2] = (T)bgr[2];
        }
    }
}

// cv2.copyMakeBorder(img, top, bottom, left, right, BORDER_REFLECT) of generate_real_video.py:80-82 on the device: the edge pixel
// is repeated (fedcba|abcdefgh|hgfedcb), borders wider than the image keep reflecting.
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

__global__ void __launch_bounds__(256) reflect_pad_kernel(const uint8_t* __restrict__ src, int N, int H, int W, int top, int left,
                                                          int PH, int PW, uint8_t* __restrict__ dst) {
    const long long total = (long long)N * PH * PW;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % PW);
        long long t = i / PW;
        const int y = (int)(t % PH);
        const int n = (int)(t / PH);
        const uint8_t* s = src + (((long long)n * H + reflect_index(y - top, H)) * W + reflect_index(x - left, W)) * 3;
        uint8_t* d = dst + i * 3;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
}

int reflect_pad_u8(const void* src, int N, int H, int W, int top, int left, int PH, int PW, void* dst, cudaStream_t st) {
    RRV_REQUIRE(src && dst, "rrv_reflect_pad_u8: NULL tensor");
    RRV_REQUIRE(N > 0 && H > 0 && W > 0 && top >= 0 && left >= 0 && PH >= top + H && PW >= left + W,
                "rrv_reflect_pad_u8: the padded size %dx%d does not contain the %dx%d image at (%d, %d)", PH, PW, H, W, top, left);
    const long long total = (long long)N * PH * PW;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    reflect_pad_kernel<<<grid, 256, 0, st>>>((const uint8_t*)src, N, H, W, top, left, PH, PW, (uint8_t*)dst);
    return check_launch("reflect_pad_kernel");
}

int postprocess_bgr(const float* in, int N, int H, int W, int y0, int x0, int h, int w, float* out, cudaStream_t st) {
    RRV_REQUIRE(in && out, "rrv_postprocess_bgr: NULL tensor");
    RRV_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, "rrv_postprocess_bgr: crop outside the image");
    const long long total = (long long)N * h * w;
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    postprocess_kernel<float><<<grid, 256, 0, st>>>(in, N, H, W, y0, x0, h, w, out);
    return check_launch("postprocess_kernel");
}

int postprocess_bgr_u8(const float* in, int N, int H, int W, int y0, int x0, int h, int w, uint8_t* out, cudaStream_t st) {
    RRV_REQUIRE(in && out, "rrv_postprocess_bgr_u8: NULL tensor");
    RRV_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, "rrv_postprocess_bgr_u8: crop outside the image");
    const long long total = (long long)N * h * w;
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    postprocess_kernel<uint8_t><<<grid, 256, 0, st>>>(in, N, H, W, y0, x0, h, w, out);
    return check_launch("postprocess_kernel<u8>");
}

// ---- fp32 weight repack: OIHW -> [k*k][Cin_pad][Cout_pad], zero padded ----
__global__ void __launch_bounds__(256) pack_w_f32_kernel(const float* __restrict__ w, int Cin, int Cout, int kk, int Cin_pad,
                                                         int Cout_pad, float* __restrict__ out) {
    const long long total = (long long)kk * Cin_pad * Cout_pad;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int co = (int)(i % Cout_pad);
        const int ci = (int)((i / Cout_pad) % Cin_pad);
        const int tap = (int)(i / ((long long)Cout_pad * Cin_pad));
        out[i] = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * kk + tap] : 0.0f;
    }
}

int pack_weights_f32(const float* w, int Cin, int Cout, int ksize, int Cin_pad, int Cout_pad, float* out, cudaStream_t st) {
    RRV_REQUIRE(w && out, "rrv_pack_weights_f32: NULL tensor");
    RRV_REQUIRE(Cin_pad >= Cin && Cout_pad >= Cout, "rrv_pack_weights_f32: padded sizes smaller than the tensor");
    const long long total = (long long)ksize * ksize * Cin_pad * Cout_pad;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
    pack_w_f32_kernel<<<grid, 256, 0, st>>>(w, Cin, Cout, ksize * ksize, Cin_pad, Cout_pad, out);
    return check_launch("pack_w_f32_kernel");
}

}  // namespace rrv

// ---- KernelFilter fold (engine.StyleEngine._fold_filter on the device) -----------------------------------------------
// The two predicted 32x32 matrices of a KernelFilter (apply_filter, style_network_global.py:194-217) are absorbed by the
// neighbouring convolutions:  Wf1 . conv_down(x) = conv_{Wf1 . Wdown}(x),  conv_up(Wf2 . t) = conv_{Wup . Wf2}(t).
// This kernel computes both products in fp32 and writes them straight into the tensor-core weight blobs
// ([tap][Cout_pad][Cin] bf16 hi, then lo; 32 inner channels), plus the folded down bias:
// frame mode predicts new filters for every frame, so the fold must not cost library GEMMs, allocations and repack launches.
namespace rrv {

constexpr int KF_IN = 32, KF_PAD = 32, KF_C = 512;      // (KF_PAD: the inner channels as the blobs carry them -- no padding since the
                                                          //  tensor-core path reads 32-channel operands as 64-byte rows)

__global__ void __launch_bounds__(256) fold_filter_kernel(const float* __restrict__ wf1, const float* __restrict__ wf2,
                                                          const float* __restrict__ down_w, const float* __restrict__ down_b,
                                                          const float* __restrict__ up_w, uint16_t* __restrict__ down_blob,
                                                          float* __restrict__ down_bias, uint16_t* __restrict__ up_blob) {
    __shared__ float s_f1[KF_IN * KF_IN], s_f2[KF_IN * KF_IN];
    for (int i = threadIdx.x; i < KF_IN * KF_IN; i += 256) { s_f1[i] = wf1[i]; s_f2[i] = wf2[i]; }
    __syncthreads();
    const long long n_down = 9LL * KF_PAD * KF_C, n_up = 9LL * KF_C * KF_PAD;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_down + n_up + KF_PAD; i += (long long)gridDim.x * 256) {
        if (i < n_down) {                       // down blob: [t][co (32)][ci (512)]
            const int ci = (int)(i % KF_C), co = (int)((i / KF_C) % KF_PAD), t = (int)(i / ((long long)KF_C * KF_PAD));
            float v = 0.0f;
            if (co < KF_IN) {
#pragma unroll 8
                for (int j = 0; j < KF_IN; ++j) v = fmaf(s_f1[co * KF_IN + j], __ldg(down_w + ((long long)j * KF_C + ci) * 9 + t), v);
            }
            uint16_t h, l;
            split_hi_lo(v, 0, h, l);
            down_blob[i] = h;
            down_blob[n_down + i] = l;
        } else if (i < n_down + n_up) {         // up blob: [t][o (512)][ci (32)]
            const long long k = i - n_down;
            const int ci = (int)(k % KF_PAD), o = (int)((k / KF_PAD) % KF_C), t = (int)(k / ((long long)KF_PAD * KF_C));
            float v = 0.0f;
            if (ci < KF_IN) {
#pragma unroll 8
                for (int j = 0; j < KF_IN; ++j) v = fmaf(__ldg(up_w + ((long long)o * KF_IN + j) * 9 + t), s_f2[j * KF_IN + ci], v);
            }
            uint16_t h, l;
            split_hi_lo(v, 0, h, l);
            up_blob[k] = h;
            up_blob[n_up + k] = l;
        } else {                                // folded down bias
            const int co = (int)(i - n_down - n_up);
            float v = 0.0f;
            if (co < KF_IN)
                for (int j = 0; j < KF_IN; ++j) v = fmaf(s_f1[co * KF_IN + j], __ldg(down_b + j), v);
            down_bias[co] = v;
        }
    }
}

int fold_filter(const float* wf1, const float* wf2, const float* down_w, const float* down_b, const float* up_w, void* down_blob,
                float* down_bias, void* up_blob, cudaStream_t st) {
    RRV_REQUIRE(wf1 && wf2 && down_w && down_b && up_w && down_blob && down_bias && up_blob, "rrv_fold_filter: NULL tensor");
    fold_filter_kernel<<<148 * 4, 256, 0, st>>>(wf1, wf2, down_w, down_b, up_w, (uint16_t*)down_blob, down_bias, (uint16_t*)up_blob);
    return check_launch("fold_filter_kernel");
}

}  // namespace rrv

// ---- backward pieces of the frozen Vgg19 loss network (train/style_networks.py:284-314; train.py:376-414) ---------------
// Loss.backward() first runs through Vgg19 (requires_grad = False: only data gradients).  The data gradient of a 3x3
// convolution is the same implicit GEMM with transposed, 180-degree rotated weights (packed once on the host side, then
// rrv_conv2d); what remains are these two memory-bound passes.
namespace rrv {

// ReLU backward fused with the conversion to operand planes: out = (y > 0) ? g (+ g2) : 0.
// g: gradient w.r.t. the ReLU output (fp32 NHWC); g2: optional second gradient arriving at the same tensor (a loss tap);
// y: the ReLU output saved by the forward pass.
__global__ void __launch_bounds__(256) relu_backward_kernel(const float* __restrict__ g, const float* __restrict__ g2,
                                                            const float* __restrict__ y, long long n8, uint16_t* __restrict__ hi,
                                                            uint16_t* __restrict__ lo) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
        const float4 a0 = *reinterpret_cast<const float4*>(g + i * 8), a1 = *reinterpret_cast<const float4*>(g + i * 8 + 4);
        const float4 y0 = *reinterpret_cast<const float4*>(y + i * 8), y1 = *reinterpret_cast<const float4*>(y + i * 8 + 4);
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float m[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
        if (g2 != nullptr) {
            const float4 b0 = *reinterpret_cast<const float4*>(g2 + i * 8), b1 = *reinterpret_cast<const float4*>(g2 + i * 8 + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = m[k] > 0.0f ? v[k] : 0.0f;
        store8(hi + i * 8, lo ? lo + i * 8 : nullptr, 0, v);
    }
}

int relu_backward(const float* g, const float* g2, const float* y, long long n, void* hi, void* lo, cudaStream_t st) {
    RRV_REQUIRE(g && y && hi, "rrv_relu_backward: NULL tensor");
    RRV_REQUIRE(n % 8 == 0, "rrv_relu_backward: element count must be a multiple of 8");
    if (n == 0) return 0;
    const int grid = (int)std::min<long long>((n / 8 + 255) / 256, 148LL * 32);
    relu_backward_kernel<<<grid, 256, 0, st>>>(g, g2, y, n / 8, (uint16_t*)hi, (uint16_t*)lo);
    return check_launch("relu_backward_kernel");
}

// nn.MaxPool2d(2, 2) backward: the gradient of a pooled pixel goes to the first maximum of its 2x2 window in row-major order
// (ATen's max_pool2d_with_indices picks the same one); a dropped odd last row / column gets zero.
__global__ void __launch_bounds__(256) maxpool_backward_kernel(const float* __restrict__ g, const float* __restrict__ y, int N, int H,
                                                               int W, int C, float* __restrict__ gx) {
    const int C4 = C >> 2;
    const long long total = (long long)N * H * W * C4;
    const int Ho = H >> 1, Wo = W >> 1;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int x = (int)(t % W); t /= W;
        const int yy = (int)(t % H);
        const int n = (int)(t / H);
        float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const int yo = yy >> 1, xo = x >> 1;
        if (yo < Ho && xo < Wo) {
            const float4 go = *reinterpret_cast<const float4*>(g + (((long long)n * Ho + yo) * Wo + xo) * C + c);
            float4 w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = *reinterpret_cast<const float4*>(y + (((long long)n * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C + c);
            const int me = ((yy & 1) << 1) | (x & 1);
            const float* wf = reinterpret_cast<const float*>(w);
            const float gof[4] = {go.x, go.y, go.z, go.w};
            float o4[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                int best = 0;
                float bv = wf[ch];
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (wf[k * 4 + ch] > bv) { bv = wf[k * 4 + ch]; best = k; }
                o4[ch] = best == me ? gof[ch] : 0.0f;
            }
            out = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        *reinterpret_cast<float4*>(gx + i * 4) = out;
    }
}

int maxpool2x2_backward(const float* g, const float* y, int N, int H, int W, int C, float* gx, cudaStream_t st) {
    RRV_REQUIRE(g && y && gx, "rrv_maxpool2x2_backward: NULL tensor");
    RRV_REQUIRE(C % 4 == 0, "rrv_maxpool2x2_backward: C must be a multiple of 4");
    const long long total = (long long)N * H * W * (C / 4);
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    maxpool_backward_kernel<<<grid, 256, 0, st>>>(g, y, N, H, W, C, gx);
    return check_launch("maxpool_backward_kernel");
}

}  // namespace rrv
