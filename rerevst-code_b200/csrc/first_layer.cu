// conv1_1 (3 -> 64, 3x3, pad 1, ReLU) fused with the input conversions that precede it in the
// reference: numpy2tensor + transform_image (test/framework.py:26-35) for uint8 BGR frames and
// TransformerNet.RGB2Gray (test/style_network_global.py:487-497).  K = 27 is far too small for
// the tensor cores; this layer is bound by writing its 64-channel output (HBM roofline).
#include "rrv_common.cuh"

namespace rrv {

constexpr int FL_TH = 8, FL_TW = 32;

struct FirstDev {
    const void* src;
    const float* w;      // [64][3][3][3]
    const float* bias;   // [64]
    uint16_t* out_hi;
    uint16_t* out_lo;
    float* out_f32;
    int N, H, W, src_kind, gray, lo_fp16;
};

__device__ __forceinline__ void normalised_rgb(const FirstDev& p, int n, int y, int x, float* v) {
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};
    if (p.src_kind == 1) {
        // uint8 HWC BGR -> RGB float -> /255 -> (x - mean) / std
        const uint8_t* s = (const uint8_t*)p.src + (((long long)n * p.H + y) * p.W + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = __fdiv_rn((float)s[2 - c], 255.0f);
            v[c] = __fdiv_rn(__fsub_rn(f, mean[c]), sd[c]);
        }
    } else {
        const float* s = (const float*)p.src;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = s[(((long long)n * 3 + c) * p.H + y) * p.W + x];
    }
    if (p.gray) {
        // RGB2Gray: de-normalise, gray = ch2*0.299 + ch1*0.587 + ch0*0.114, re-normalise per channel
        float im[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) im[c] = __fadd_rn(__fmul_rn(v[c], sd[c]), mean[c]);
        const float g = __fadd_rn(__fadd_rn(__fmul_rn(im[2], 0.299f), __fmul_rn(im[1], 0.587f)),
                                  __fmul_rn(im[0], 0.114f));
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(g, mean[c]), sd[c]);
    }
}

// Tile = 8 rows x 32 columns; thread = 4 adjacent pixels x 16 output channels, so every weight
// float4 read from shared memory feeds 16 FMAs and the kernel is FFMA- rather than LDS-bound.
__global__ void __launch_bounds__(256) first_layer_kernel(const FirstDev p) {
    constexpr int PH = FL_TH + 2, PW = FL_TW + 2;
    __shared__ float s_in[3][PH][PW + 2];
    __shared__ __align__(16) float s_w[27][64];
    __shared__ float s_b[64];
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + FL_TW - 1) / FL_TW;
    const int oy0 = (blockIdx.x / tiles_x) * FL_TH, ox0 = (blockIdx.x % tiles_x) * FL_TW;
    const int n = blockIdx.y;

    for (int i = tid; i < 27 * 64; i += 256) {
        const int co = i & 63, k = i >> 6;      // k = c*9 + dy*3 + dx
        s_w[k][co] = p.w[co * 27 + k];
    }
    if (tid < 64) s_b[tid] = p.bias[tid];
    for (int i = tid; i < PH * PW; i += 256) {
        const int r = i / PW, c = i % PW;
        const int y = oy0 - 1 + r, x = ox0 - 1 + c;
        float v[3] = {0.f, 0.f, 0.f};
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) normalised_rgb(p, n, y, x, v);
        s_in[0][r][c] = v[0];
        s_in[1][r][c] = v[1];
        s_in[2][r][c] = v[2];
    }
    __syncthreads();

    const int q = tid & 3, pg = tid >> 2;
    const int r = pg >> 3, c0 = (pg & 7) * 4;
    float acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[j][k] = s_b[q * 16 + k];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            float a[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) a[i] = s_in[ch][r + dy][c0 + i];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4* wp = reinterpret_cast<const float4*>(&s_w[ch * 9 + dy * 3 + dx][q * 16]);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 w = wp[k4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][4 * k4 + 0] = fmaf(a[j + dx], w.x, acc[j][4 * k4 + 0]);
                        acc[j][4 * k4 + 1] = fmaf(a[j + dx], w.y, acc[j][4 * k4 + 1]);
                        acc[j][4 * k4 + 2] = fmaf(a[j + dx], w.z, acc[j][4 * k4 + 2]);
                        acc[j][4 * k4 + 3] = fmaf(a[j + dx], w.w, acc[j][4 * k4 + 3]);
                    }
                }
            }
        }
    const int oy = oy0 + r;
    if (oy >= p.H) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + c0 + j;
        if (ox >= p.W) continue;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[j][k] = fmaxf(acc[j][k], 0.0f);
        const long long o = (((long long)n * p.H + oy) * p.W + ox) * 64 + q * 16;
        if (p.out_hi != nullptr) {
            store8(p.out_hi + o, p.out_lo ? p.out_lo + o : nullptr, p.lo_fp16, acc[j]);
            store8(p.out_hi + o + 8, p.out_lo ? p.out_lo + o + 8 : nullptr, p.lo_fp16, acc[j] + 8);
        }
        if (p.out_f32 != nullptr) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<float4*>(p.out_f32 + o + 4 * k4) =
                    make_float4(acc[j][4 * k4], acc[j][4 * k4 + 1], acc[j][4 * k4 + 2], acc[j][4 * k4 + 3]);
        }
    }
}

int first_layer(const void* src, int src_kind, int gray, int N, int H, int W, const float* w, const float* bias,
                void* out_hi, void* out_lo, float* out_f32, cudaStream_t st) {
    RRV_REQUIRE(src && w && bias, "rrv_first_layer: NULL input");
    RRV_REQUIRE(src_kind == 0 || src_kind == 1, "rrv_first_layer: src_kind must be 0 (fp32 NCHW) or 1 (u8 HWC BGR)");
    RRV_REQUIRE(N > 0 && H > 0 && W > 0, "rrv_first_layer: empty input");
    RRV_REQUIRE(out_hi || out_f32, "rrv_first_layer: no output requested");
    FirstDev d{src, w, bias, (uint16_t*)out_hi, (uint16_t*)out_lo, out_f32, N, H, W, src_kind, gray, g_lo_fp16};
    dim3 grid(ceil_div(W, FL_TW) * ceil_div(H, FL_TH), N);
    first_layer_kernel<<<grid, 256, 0, st>>>(d);
    return check_launch("first_layer_kernel");
}

}  // namespace rrv
