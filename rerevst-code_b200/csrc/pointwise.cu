// HBM-bound pointwise kernels: 2x2 max-pool on planes, the epilogue chain alone (pre-pass
// normalisation), layout conversions, the de-normalise/clamp/BGR postprocess and the fp32
// weight repack.  All vectorised to 16-byte accesses along the channel dimension.
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
__global__ void __launch_bounds__(256) pointwise_kernel(const float* __restrict__ in, long long in_bs, int N, int H, int W,
                                                        int C, EpiDev ep, int out_mode, uint16_t* out_hi,
                                                        uint16_t* out_lo, float* out_f32) {
    const int C8 = C >> 3;
    const long long total = (long long)N * H * W * C8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c0 = (int)(i % C8) * 8;
        long long t = i / C8;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        const float* src = in + (long long)n * in_bs + ((long long)y * W + x) * C + c0;
        float v[8];
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        apply_epilogue<8>(ep, v, n, y, x, c0);
        const long long o = (((long long)n * H + y) * W + x) * C + c0;
        if (out_mode == RRV_OUT_PLANES) {
            store8(out_hi + o, out_lo ? out_lo + o : nullptr, ep.lo_fp16, v);
        } else {
            *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(out_f32 + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

int pointwise(const float* in, long long in_bs, int N, int H, int W, int C, const rrv_epilogue* ep, int out_mode,
              void* out_hi, void* out_lo, float* out_f32, cudaStream_t st) {
    RRV_REQUIRE(in && ep, "rrv_pointwise: NULL input");
    RRV_REQUIRE(C % 8 == 0, "rrv_pointwise: C must be a multiple of 8 (got %d)", C);
    RRV_REQUIRE(out_mode == RRV_OUT_PLANES || out_mode == RRV_OUT_F32_NHWC, "rrv_pointwise: bad out_mode %d", out_mode);
    RRV_REQUIRE(out_mode == RRV_OUT_PLANES ? out_hi != nullptr : out_f32 != nullptr, "rrv_pointwise: NULL output");
    const long long total = (long long)N * H * W * (C / 8);
    if (total == 0) return 0;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    pointwise_kernel<<<grid, 256, 0, st>>>(in, in_bs, N, H, W, C, make_epi(*ep, C), out_mode, (uint16_t*)out_hi,
                                           (uint16_t*)out_lo, out_f32);
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
