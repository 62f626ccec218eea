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

// Normalised RGB of one pixel as the reference's first convolution would see it WITHOUT RGB2Gray.
__device__ __forceinline__ void load_normalised(const FirstDev& p, int n, int y, int x, float* v) {
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
}

// RGB2Gray (:487-497) up to its last step: de-normalise, gray = ch2*0.299 + ch1*0.587 + ch0*0.114 (the BGR weights
// applied to the RGB tensor, as the reference does).  The re-normalisation per channel follows in the callers.
__device__ __forceinline__ float gray_value(const float* v) {
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};
    float im[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) im[c] = __fadd_rn(__fmul_rn(v[c], sd[c]), mean[c]);
    return __fadd_rn(__fadd_rn(__fmul_rn(im[2], 0.299f), __fmul_rn(im[1], 0.587f)), __fmul_rn(im[0], 0.114f));
}

__device__ __forceinline__ void normalised_rgb(const FirstDev& p, int n, int y, int x, float* v) {
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};
    load_normalised(p, n, y, x, v);
    if (p.gray) {
        const float g = gray_value(v);
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

// Gray frames (every content frame: TransformerNet.forward and .add run RGB2Gray first): the three input
// channels are affine functions of ONE gray value, v_c = (g - mean_c) / sd_c = a_c u + b_c with the centred
// variable u = (g - GM) / GS, so the 27-tap convolution collapses to 9 taps on u:
//     out[co] = bias[co] + sum_{t in bounds} (W1[t][co] u_t + W0[t][co]),
//     W1[t][co] = sum_c w[co][c][t] a_c,   W0[t][co] = sum_c w[co][c][t] b_c   (|b_c| < 0.25: no cancellation).
// Zero padding pads v (not g) with zeros: out-of-bounds taps contribute neither term, which the border
// pixels fix up by subtracting the W0 of their missing taps.  3x fewer FMAs: the kernel becomes store-bound.
constexpr float GRAY_GM = 0.449f, GRAY_GS = 0.226f;

constexpr int FL_TPB = 4;      // consecutive tiles (along x) per block: the weight tables are built once per block

__device__ __forceinline__ float gray_u(const FirstDev& p, int n, int y, int x) {
    if (y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.0f;
    if (p.src_kind == 1) {
        // uint8 BGR: normalise -> de-normalise -> gray collapses to gray = (0.299 B + 0.587 G + 0.114 R) / 255 (the reference
        // applies the BGR weights to its RGB tensor: ch2 = B); the reference's six divisions only add rounding noise of 1e-7
        const uint8_t* s = (const uint8_t*)p.src + (((long long)n * p.H + y) * p.W + x) * 3;
        const float g = fmaf(0.114f, (float)s[2], fmaf(0.587f, (float)s[1], 0.299f * (float)s[0])) * (1.0f / 255.0f);
        return (g - GRAY_GM) * (1.0f / GRAY_GS);
    }
    float v[3];
    load_normalised(p, n, y, x, v);
    return (gray_value(v) - GRAY_GM) * (1.0f / GRAY_GS);
}

__global__ void __launch_bounds__(256) first_layer_gray_kernel(const FirstDev p) {
    constexpr int PH = FL_TH + 2, PW = FL_TW + 2;
    __shared__ float s_in[2][PH][PW + 2];
    __shared__ __align__(16) float s_w1[9][64];
    __shared__ __align__(16) float s_w0[9][64];
    __shared__ float s_b[64];
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + FL_TW - 1) / FL_TW;
    const int groups_x = (tiles_x + FL_TPB - 1) / FL_TPB;
    const int oy0 = (blockIdx.x / groups_x) * FL_TH;
    const int tx0 = (blockIdx.x % groups_x) * FL_TPB;
    const int ntile = min(FL_TPB, tiles_x - tx0);
    const int n = blockIdx.y;
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};

    for (int i = tid; i < 9 * 64; i += 256) {
        const int co = i & 63, t = i >> 6;
        float w1 = 0.0f, w0 = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float w = p.w[co * 27 + c * 9 + t];
            w1 = fmaf(w, GRAY_GS / sd[c], w1);
            w0 = fmaf(w, (GRAY_GM - mean[c]) / sd[c], w0);
        }
        s_w1[t][co] = w1;
        s_w0[t][co] = w0;
    }
    // halo pixels of a tile handled by this thread: tid and tid + 256 (PH * PW = 340)
    const int h0r = tid / PW, h0c = tid % PW;
    const int h1 = tid + 256, h1r = h1 / PW, h1c = h1 % PW;
    const bool has1 = h1 < PH * PW;
    {
        const int ox0 = tx0 * FL_TW;
        s_in[0][h0r][h0c] = gray_u(p, n, oy0 - 1 + h0r, ox0 - 1 + h0c);
        if (has1) s_in[0][h1r][h1c] = gray_u(p, n, oy0 - 1 + h1r, ox0 - 1 + h1c);
    }
    __syncthreads();
    if (tid < 64) {
        float b = p.bias[tid];
#pragma unroll
        for (int t = 0; t < 9; ++t) b += s_w0[t][tid];
        s_b[tid] = b;
    }
    __syncthreads();

    // thread = 8 adjacent pixels x 8 output channels: the 8 lanes of a pixel store one full 128-byte line per plane
    const int q = tid & 7, pg = tid >> 3;
    const int r = pg >> 2, c0 = (pg & 3) * 8;
    const int oy = oy0 + r;
    const bool edge_row = oy == 0 || oy == p.H - 1;
    for (int t = 0; t < ntile; ++t) {
        const int ox0 = (tx0 + t) * FL_TW;
        const int buf = t & 1;
        // the next tile's halo goes in flight before this tile's arithmetic
        float nx0 = 0.0f, nx1 = 0.0f;
        const bool more = t + 1 < ntile;
        if (more) {
            nx0 = gray_u(p, n, oy0 - 1 + h0r, ox0 + FL_TW - 1 + h0c);
            if (has1) nx1 = gray_u(p, n, oy0 - 1 + h1r, ox0 + FL_TW - 1 + h1c);
        }
        float acc[8][8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[j][k] = s_b[q * 8 + k];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            float a[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) a[i] = s_in[buf][r + dy][c0 + i];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4* wp = reinterpret_cast<const float4*>(&s_w1[dy * 3 + dx][q * 8]);
                const float4 wa = wp[0], wb = wp[1];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[j][0] = fmaf(a[j + dx], wa.x, acc[j][0]);
                    acc[j][1] = fmaf(a[j + dx], wa.y, acc[j][1]);
                    acc[j][2] = fmaf(a[j + dx], wa.z, acc[j][2]);
                    acc[j][3] = fmaf(a[j + dx], wa.w, acc[j][3]);
                    acc[j][4] = fmaf(a[j + dx], wb.x, acc[j][4]);
                    acc[j][5] = fmaf(a[j + dx], wb.y, acc[j][5]);
                    acc[j][6] = fmaf(a[j + dx], wb.z, acc[j][6]);
                    acc[j][7] = fmaf(a[j + dx], wb.w, acc[j][7]);
                }
            }
        }
        if (oy < p.H) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ox = ox0 + c0 + j;
                if (ox >= p.W) continue;
                if (edge_row || ox == 0 || ox == p.W - 1) {       // taps that fall outside the image carry no W0 term
                    for (int tp = 0; tp < 9; ++tp) {
                        const int y = oy + tp / 3 - 1, x = ox + tp % 3 - 1;
                        if (y >= 0 && y < p.H && x >= 0 && x < p.W) continue;
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[j][k] -= s_w0[tp][q * 8 + k];
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[j][k] = fmaxf(acc[j][k], 0.0f);
                const long long o = (((long long)n * p.H + oy) * p.W + ox) * 64 + q * 8;
                if (p.out_hi != nullptr) store8(p.out_hi + o, p.out_lo ? p.out_lo + o : nullptr, p.lo_fp16, acc[j]);
                if (p.out_f32 != nullptr) {
                    *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    *reinterpret_cast<float4*>(p.out_f32 + o + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
                }
            }
        }
        if (more) {                       // the other buffer was last read two iterations ago (one barrier in between)
            s_in[buf ^ 1][h0r][h0c] = nx0;
            if (has1) s_in[buf ^ 1][h1r][h1c] = nx1;
        }
        __syncthreads();
    }
}

// The same 9-tap gray convolution for the per-frame path (planes output, bf16 lo): persistent blocks whose WARPS work
// independently.  Round-2 profile of the kernel above at 1216 x 2048 (profiles/r2_first_layer_before.txt): 183 us against an
// 88 us floor (578 MB written at the 7.2 TB/s a memset reaches), 20.8 instructions per output of which 9.1 are the FFMAs --
// per-pixel bound checks, 64-bit address arithmetic per store, block barriers per tile and the per-block table build made up
// most of the rest.  Here:
//   * unit of work = one image row x 32 columns, taken by ONE warp (lane = 8 pixels x 8 channels as above); the warp keeps its
//     own 3 x 34 halo in shared memory (double-buffered, __syncwarp only): no block barrier after the prologue, so the store
//     bursts of one warp overlap the arithmetic of the others;
//   * grid = 2 blocks per SM, units dealt round-robin to the warps: the weight tables are built once per block;
//   * interior units (every unit of a frame padded to a multiple of 32 except the border rows and the two border column
//     groups) run without bound checks; one pointer per unit and plane, immediate offsets per pixel;
//   * the next unit's uint8 pixels are fetched with cp.async into shared memory and only converted after this unit's stores
//     (loaded into registers, the compiler converted them at once and a third of the issue slots waited for L2).
// uint8 frames with W % 4 == 0 only (word-aligned rows); everything else takes the kernel above.
template <bool LO>
__global__ void __launch_bounds__(256, 2) first_layer_gray_rows_kernel(const FirstDev p, int units, int tiles_x) {
    __shared__ __align__(16) float s_w1[9][64];
    __shared__ __align__(16) float s_w0[9][64];
    __shared__ float s_b[64];
    __shared__ __align__(16) float s_h[8][2][3][40];      // [warp][buffer][row][column]: column 0 = x0 - 1, 34 used
    __shared__ __align__(16) uint32_t s_raw[8][3][28];    // [warp][row]: the 27 aligned words that hold the row's 34 BGR pixels
    const int tid = threadIdx.x;
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float sd[3] = {0.229f, 0.224f, 0.225f};
    for (int i = tid; i < 9 * 64; i += 256) {
        const int co = i & 63, t = i >> 6;
        float w1 = 0.0f, w0 = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float w = p.w[co * 27 + c * 9 + t];
            w1 = fmaf(w, GRAY_GS / sd[c], w1);
            w0 = fmaf(w, (GRAY_GM - mean[c]) / sd[c], w0);
        }
        s_w1[t][co] = w1;
        s_w0[t][co] = w0;
    }
    __syncthreads();
    if (tid < 64) {
        float b = p.bias[tid];
#pragma unroll
        for (int t = 0; t < 9; ++t) b += s_w0[t][tid];
        s_b[tid] = b;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int q = lane & 7, c0 = (lane >> 3) * 8;
    const int nw = gridDim.x * 8;
    int u = blockIdx.x * 8 + warp;
    if (u >= units) return;
    float bq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bq[k] = s_b[q * 8 + k];
    const uint8_t* src = (const uint8_t*)p.src;
    const long long total = (long long)p.N * p.H * p.W * 3;

    auto coords = [&](int unit, int& n_, int& oy_, int& ox0_) {
        const int r = unit / tiles_x;
        ox0_ = (unit - r * tiles_x) * FL_TW;
        n_ = r / p.H;
        oy_ = r - n_ * p.H;
    };
    // The uint8 pixels of a unit's halo travel global -> shared as aligned 32-bit words with cp.async (no registers, nothing
    // waits for them until the unit before has been computed and stored); W % 4 == 0 keeps every row's first byte word-aligned.
    auto fetch_raw = [&](int n_, int oy_, int ox0_) {
        if (lane < 27) {
            const int d = (3 * ox0_ + 1) & 3;                           // = 3 (x0 - 1) mod 4
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int y = oy_ - 1 + rr;
                const long long g = ((long long)n_ * p.H + y) * p.W * 3 + (3 * (ox0_ - 1) - d) + 4 * lane;
                if (y >= 0 && y < p.H && g >= 0 && g + 4 <= total) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_raw[warp][rr][lane]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + g) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto convert_raw = [&](int oy_, int ox0_, int buf) {              // raw words -> centred gray values of the halo
        const int d = (3 * ox0_ + 1) & 3;
        const uint8_t* rb = reinterpret_cast<const uint8_t*>(&s_raw[warp][0][0]);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int rr = it < 3 ? it : (lane >> 1), c = it < 3 ? lane : FL_TW + (lane & 1);
            if (it == 3 && lane >= 6) break;
            const int y = oy_ - 1 + rr, x = ox0_ - 1 + c;
            float uv = 0.0f;
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                const uint8_t* s3 = rb + rr * 112 + d + 3 * c;
                const float g = fmaf(0.114f, (float)s3[2], fmaf(0.587f, (float)s3[1], 0.299f * (float)s3[0])) * (1.0f / 255.0f);
                uv = (g - GRAY_GM) * (1.0f / GRAY_GS);
            }
            s_h[warp][buf][rr][c] = uv;
        }
    };
    int n, oy, ox0;
    coords(u, n, oy, ox0);
    fetch_raw(n, oy, ox0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    convert_raw(oy, ox0, 0);
    __syncwarp();

    for (int buf = 0;; buf ^= 1) {
        const int u2 = u + nw;
        const bool more = u2 < units;
        int n2 = 0, oy2 = 0, ox2 = 0;
        if (more) {                      // the next unit's pixels go in flight before this unit's arithmetic
            coords(u2, n2, oy2, ox2);
            fetch_raw(n2, oy2, ox2);
        }
        float acc[8][8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[j][k] = bq[k];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float* hr = &s_h[warp][buf][dy][c0];
            const float4 a0 = *reinterpret_cast<const float4*>(hr);
            const float4 a1 = *reinterpret_cast<const float4*>(hr + 4);
            const float2 a2 = *reinterpret_cast<const float2*>(hr + 8);
            const float a[10] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y};
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4* wp = reinterpret_cast<const float4*>(&s_w1[dy * 3 + dx][q * 8]);
                const float4 wa = wp[0], wb = wp[1];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[j][0] = fmaf(a[j + dx], wa.x, acc[j][0]);
                    acc[j][1] = fmaf(a[j + dx], wa.y, acc[j][1]);
                    acc[j][2] = fmaf(a[j + dx], wa.z, acc[j][2]);
                    acc[j][3] = fmaf(a[j + dx], wa.w, acc[j][3]);
                    acc[j][4] = fmaf(a[j + dx], wb.x, acc[j][4]);
                    acc[j][5] = fmaf(a[j + dx], wb.y, acc[j][5]);
                    acc[j][6] = fmaf(a[j + dx], wb.z, acc[j][6]);
                    acc[j][7] = fmaf(a[j + dx], wb.w, acc[j][7]);
                }
            }
        }
        const int x0 = ox0 + c0;
        const bool edge_row = oy == 0 || oy == p.H - 1;
        if (edge_row || x0 == 0 || x0 + 8 >= p.W) {       // taps that fall outside the image carry no W0 term (border pixels only)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ox = x0 + j;
                if (ox < p.W && (edge_row || ox == 0 || ox == p.W - 1)) {
#pragma unroll 1
                    for (int tp = 0; tp < 9; ++tp) {
                        const int y = oy + tp / 3 - 1, x = ox + tp % 3 - 1;
                        if (y >= 0 && y < p.H && x >= 0 && x < p.W) continue;
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[j][k] -= s_w0[tp][q * 8 + k];
                    }
                }
            }
        }
        const long long o = (((long long)n * p.H + oy) * p.W + x0) * 64 + q * 8;
        uint16_t* ph = p.out_hi + o;
        uint16_t* pl = LO ? p.out_lo + o : nullptr;
        const int npx = min(8, p.W - x0);          // < 8 (or <= 0) only in the last, partial column tile
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= npx) break;
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v0 = fmaxf(acc[j][2 * i], 0.0f), v1 = fmaxf(acc[j][2 * i + 1], 0.0f);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
                if (LO) {
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hw[i] << 16), v1 - __uint_as_float(hw[i] & 0xffff0000u));
                    lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
                }
            }
            *reinterpret_cast<uint4*>(ph + j * 64) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            if (LO) *reinterpret_cast<uint4*>(pl + j * 64) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        if (!more) break;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        convert_raw(oy2, ox2, buf ^ 1);  // s_h[buf ^ 1] was last read two units ago; a __syncwarp lies in between
        __syncwarp();
        u = u2; n = n2; oy = oy2; ox0 = ox2;
    }
}

static int fl_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int first_layer(const void* src, int src_kind, int gray, int N, int H, int W, const float* w, const float* bias,
                void* out_hi, void* out_lo, float* out_f32, cudaStream_t st) {
    RRV_REQUIRE(src && w && bias, "rrv_first_layer: NULL input");
    RRV_REQUIRE(src_kind == 0 || src_kind == 1, "rrv_first_layer: src_kind must be 0 (fp32 NCHW) or 1 (u8 HWC BGR)");
    RRV_REQUIRE(N > 0 && H > 0 && W > 0, "rrv_first_layer: empty input");
    RRV_REQUIRE(out_hi || out_f32, "rrv_first_layer: no output requested");
    FirstDev d{src, w, bias, (uint16_t*)out_hi, (uint16_t*)out_lo, out_f32, N, H, W, src_kind, gray, g_lo_fp16};
    dim3 grid(ceil_div(W, FL_TW) * ceil_div(H, FL_TH), N);
    if (gray && src_kind == 1 && out_hi && !out_f32 && !g_lo_fp16 && W % 4 == 0 && ((uintptr_t)src & 3) == 0 &&
        (long long)N * H * ceil_div(W, FL_TW) < (1ll << 30)) {
        const int tiles_x = ceil_div(W, FL_TW), units = N * H * tiles_x;
        const int blocks = std::min(ceil_div(units, 8), 2 * fl_num_sms());
        if (out_lo) first_layer_gray_rows_kernel<true><<<blocks, 256, 0, st>>>(d, units, tiles_x);
        else first_layer_gray_rows_kernel<false><<<blocks, 256, 0, st>>>(d, units, tiles_x);
        return check_launch("first_layer_gray_rows_kernel");
    }
    if (gray) {
        dim3 ggrid(ceil_div(ceil_div(W, FL_TW), FL_TPB) * ceil_div(H, FL_TH), N);
        first_layer_gray_kernel<<<ggrid, 256, 0, st>>>(d);
        return check_launch("first_layer_gray_kernel");
    }
    first_layer_kernel<<<grid, 256, 0, st>>>(d);
    return check_launch("first_layer_kernel");
}

}  // namespace rrv
