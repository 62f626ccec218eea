// fp32 CUDA-core convolution over hi/lo planes (bring-up path and on-GPU cross-check of the
// tcgen05 kernel; also runs the once-per-clip pre-pass and the style encoder).
//
// Replaces nn.Conv2d / F.conv2d of style_network_global.py:103-105, 181-187, 205, 275-281, 341
// with the nearest x2 upsample of :113 folded into the gather and the pointwise chain of
// :52-55, :116-122, :364 fused into the epilogue (rrv_common.cuh: apply_epilogue).
#include "rrv_common.cuh"

namespace rrv {

struct ConvDev {
    const uint16_t* in_hi;
    const uint16_t* in_lo;
    const float* w;          // [KS*KS][Cin][Cout_pad]
    uint16_t* out_hi;
    uint16_t* out_lo;
    float* out_f32;
    int N, H, W, Cin, Cout, Cout_pad, ups, in_H, in_W, out_mode, out_C, lo_fp16;
    EpiDev ep;
};

constexpr int TH = 8, TW = 16, CK = 16;

template <int KS, int CPT>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvDev p) {
    constexpr int PH = TH + KS - 1, PW = TW + KS - 1, PWP = PW + 1;
    constexpr int TN = 8 * CPT;
    constexpr int PAD = KS / 2;
    extern __shared__ float smem[];
    float* s_in = smem;                       // [CK][PH][PWP]
    float* s_w = smem + CK * PH * PWP;        // [KS*KS][CK][TN]

    const int tid = threadIdx.x;
    const int tx = tid & 7, pg = tid >> 3;
    const int prow = pg >> 2, pcol0 = (pg & 3) * 4;
    const int tiles_x = (p.W + TW - 1) / TW;
    const int oy0 = (blockIdx.x / tiles_x) * TH, ox0 = (blockIdx.x % tiles_x) * TW;
    const int co0 = blockIdx.y * TN;
    const int n = blockIdx.z;

    float acc[4][CPT];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < CPT; ++k) acc[j][k] = 0.0f;

    const long long in_img = (long long)n * p.in_H * p.in_W * p.Cin;

    for (int c0 = 0; c0 < p.Cin; c0 += CK) {
        // ---- stage the input patch (output coordinates, nearest-upsample gather) ----
        for (int it = tid; it < PH * PW * 2; it += 256) {
            const int half = it & 1, pix = it >> 1;
            const int r = pix / PW, c = pix % PW;
            const int oy = oy0 - PAD + r, ox = ox0 - PAD + c;
            float v[8];
            if (oy >= 0 && oy < p.H && ox >= 0 && ox < p.W) {
                const long long off = in_img + ((long long)(oy >> p.ups) * p.in_W + (ox >> p.ups)) * p.Cin + c0 + half * 8;
                load8(p.in_hi + off, p.in_lo ? p.in_lo + off : nullptr, p.lo_fp16, v);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) s_in[((half * 8 + i) * PH + r) * PWP + c] = v[i];
        }
        // ---- stage the weights of this channel chunk ----
        for (int it = tid; it < KS * KS * CK * TN / 4; it += 256) {
            const int e = it * 4;
            const int co = e % TN, ci = (e / TN) % CK, tap = e / (TN * CK);
            const float4 w4 = *reinterpret_cast<const float4*>(
                p.w + ((long long)tap * p.Cin + c0 + ci) * p.Cout_pad + co0 + co);
            *reinterpret_cast<float4*>(s_w + e) = w4;
        }
        __syncthreads();
#pragma unroll 4
        for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                float a[4 + KS - 1];
                const float* row = s_in + (ci * PH + prow + dy) * PWP + pcol0;
#pragma unroll
                for (int i = 0; i < 4 + KS - 1; ++i) a[i] = row[i];
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    float w[CPT];
                    const float* wp = s_w + ((dy * KS + dx) * CK + ci) * TN + tx * CPT;
                    if (CPT == 8) {
                        const float4 w0 = *reinterpret_cast<const float4*>(wp);
                        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
                        w[0] = w0.x; w[1 % CPT] = w0.y; w[2 % CPT] = w0.z; w[3 % CPT] = w0.w;
                        w[4 % CPT] = w1.x; w[5 % CPT] = w1.y; w[6 % CPT] = w1.z; w[7 % CPT] = w1.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < CPT; ++k) w[k] = wp[k];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int k = 0; k < CPT; ++k) acc[j][k] = fmaf(a[j + dx], w[k], acc[j][k]);
                }
            }
        }
        __syncthreads();
    }

    // ---- fused epilogue + store ----
    const int oy = oy0 + prow;
    if (oy >= p.H) return;
    const int c_out = co0 + tx * CPT;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + pcol0 + j;
        if (ox >= p.W) continue;
        if (c_out >= p.Cout) continue;
        float v[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) v[k] = acc[j][k];
        apply_epilogue<CPT>(p.ep, v, n, oy, ox, c_out);
        const long long pix = ((long long)n * p.H + oy) * p.W + ox;
        if (p.out_mode == RRV_OUT_PLANES) {
            if (CPT == 8) store8(p.out_hi + pix * p.Cout + c_out, p.out_lo ? p.out_lo + pix * p.Cout + c_out : nullptr,
                                 p.lo_fp16, v);
        } else if (p.out_mode == RRV_OUT_F32_NHWC) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) p.out_f32[pix * p.Cout + c_out + k] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < CPT; ++k)
                if (c_out + k < p.out_C)
                    p.out_f32[(((long long)n * p.out_C + c_out + k) * p.H + oy) * p.W + ox] = v[k];
        }
    }
}

template <int KS, int CPT>
static int launch(const ConvDev& d, cudaStream_t st) {
    constexpr int PH = TH + KS - 1, PWP = TW + KS;
    constexpr int TN = 8 * CPT;
    const size_t smem = sizeof(float) * (CK * PH * PWP + KS * KS * CK * TN);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_ffma_kernel<KS, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid(ceil_div(d.W, TW) * ceil_div(d.H, TH), ceil_div(d.Cout_pad, TN), d.N);
    conv_ffma_kernel<KS, CPT><<<grid, 256, smem, st>>>(d);
    return check_launch("conv_ffma_kernel");
}

int conv2d_ffma(const rrv_conv* p, cudaStream_t st) {
    RRV_REQUIRE(p->ksize == 1 || p->ksize == 3, "rrv_conv2d: ksize must be 1 or 3 (got %d)", p->ksize);
    RRV_REQUIRE(p->Cin % CK == 0, "rrv_conv2d: Cin must be a multiple of %d (got %d)", CK, p->Cin);
    RRV_REQUIRE(p->w_f32 != nullptr, "rrv_conv2d(FFMA): w_f32 is NULL");
    RRV_REQUIRE(!p->pool, "rrv_conv2d(FFMA): the fused max-pool exists on the tcgen05 path only (use rrv_maxpool2x2)");
    RRV_REQUIRE(p->in_hi != nullptr, "rrv_conv2d: in_hi is NULL");
    RRV_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "rrv_conv2d: empty output %dx%dx%d", p->N, p->H, p->W);
    RRV_REQUIRE(!p->ups || (p->H % 2 == 0 && p->W % 2 == 0), "rrv_conv2d: ups needs even output size");
    ConvDev d;
    d.in_hi = (const uint16_t*)p->in_hi;
    d.in_lo = (const uint16_t*)p->in_lo;
    d.w = p->w_f32;
    d.out_hi = (uint16_t*)p->out_hi;
    d.out_lo = (uint16_t*)p->out_lo;
    d.out_f32 = p->out_f32;
    d.N = p->N; d.H = p->H; d.W = p->W; d.Cin = p->Cin; d.Cout = p->Cout;
    d.ups = p->ups ? 1 : 0;
    d.in_H = p->H >> d.ups; d.in_W = p->W >> d.ups;
    d.out_mode = p->out_mode; d.out_C = p->out_C;
    d.lo_fp16 = g_lo_fp16;
    d.ep = make_epi(p->ep, p->Cout);
    const bool small = p->Cout < 8;
    if (small) {
        RRV_REQUIRE(p->out_mode != RRV_OUT_PLANES, "rrv_conv2d: planes output needs Cout %% 8 == 0");
        d.Cout_pad = 8;
        RRV_REQUIRE(p->out_f32 != nullptr, "rrv_conv2d: out_f32 is NULL");
        return p->ksize == 3 ? launch<3, 1>(d, st) : launch<1, 1>(d, st);
    }
    RRV_REQUIRE(p->Cout % 8 == 0, "rrv_conv2d: Cout must be < 8 or a multiple of 8 (got %d)", p->Cout);
    d.Cout_pad = (p->Cout + 63) / 64 * 64;
    RRV_REQUIRE(p->stats == nullptr, "rrv_conv2d(ffma): fused statistics are a tensor-core path feature (use rrv_channel_stats)");
    RRV_REQUIRE(p->out_mode == RRV_OUT_PLANES || p->out_mode == RRV_OUT_F32_NHWC || p->out_mode == RRV_OUT_F32_NCHW,
                "rrv_conv2d(ffma): out_mode %d is a tensor-core path feature (use rrv_postprocess_bgr after an NCHW output)", p->out_mode);
    if (p->out_mode == RRV_OUT_PLANES) RRV_REQUIRE(p->out_hi != nullptr, "rrv_conv2d: out_hi is NULL");
    else RRV_REQUIRE(p->out_f32 != nullptr, "rrv_conv2d: out_f32 is NULL");
    return p->ksize == 3 ? launch<3, 8>(d, st) : launch<1, 8>(d, st);
}

}  // namespace rrv
