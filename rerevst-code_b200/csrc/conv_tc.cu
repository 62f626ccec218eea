// Tensor-core convolution: im2col-free implicit GEMM on tcgen05 (sm_100a).
//
// Replaces every nn.Conv2d / F.conv2d of the per-frame path except conv1_1
// (test/style_network_global.py:103-105, 181-187, 205, 275-281, 341) together with the nearest x2
// upsample of ResidualBlock.forward (:113) and the pointwise chains that follow each convolution
// (:52-55, :116-122, :364), which run in the epilogue (rrv_common.cuh: apply_epilogue).
//
// GEMM view: D[M = 128 output pixels][N = Cout tile] += A[M][K] * B[N][K]^T with K = taps x Cin.
//   A  TMA boxes of the NHWC input itself (64-channel chunks); rows / columns outside the image are
//      zero-filled by TMA, which IS the conv's zero padding.  The box lands in shared memory as rows
//      of 128 bytes with the 128-byte swizzle, i.e. the canonical K-major operand layout of
//      tcgen05.mma -- no im2col buffer anywhere.  One box serves several taps: the tap only moves
//      the operand descriptor's start address by whole swizzle atoms (see the three main-loop
//      shapes at conv_tc2_kernel / conv2d_tc2).
//   B  weights repacked once at load time to [tap][Cout][Cin] bf16, TMA boxes per k-chunk.
//   D  fp32 accumulators in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//      tile i+1.  Persistent CTAs (one per SM, or CTA pairs with cta_group::2) walk the tile list.
// fp32 accuracy ("x3"): every fp32 operand v is carried as hi = bf16(v), lo = bf16(v - hi) and each
// k-slice issues Ahi*Bhi + Ahi*Blo + Alo*Bhi into the same accumulator (the dropped lo*lo term is
// 2^-18 relative).  With lo == NULL a single bf16 MMA is issued (BASELINE config 3).
// Nearest x2 upsample: the 3x3 conv over the upsampled image collapses, per output parity (py,px),
// to a 2x2 conv over the low-resolution input with summed weights (9 -> 4 taps); the four phases
// are four GEMMs over the same low-res tile that scatter to interleaved output pixels.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) and TMEM owner,
// warps 2..9 = epilogue (tcgen05.ld -> fused pointwise chain -> stores; the merged-tap layers stage
// their rows in shared memory and leave through TMA stores; row-reuse layers with a same-size residual
// bring it in by cp.async one chunk ahead and run the chain in place: epilogue_chunk_rr); the
// per-channel constants of the chain are gathered once per CTA into a shared-memory table.
// Operand rows are 128 bytes (64 channels, SWIZZLE_128B) or, for a 32-channel input in the row-reuse
// loop, 64 bytes (SWIZZLE_64B boxes and descriptors).
#include <cuda.h>
#include <math_constants.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rrv_common.cuh"
#include "tc_ptx.cuh"

#ifndef RRV_EPI_PREFETCH
#define RRV_EPI_PREFETCH 2      // before waiting for the accumulators: 1 = load the first chunk's residual into registers
#endif                          // (spills at 168 registers), 2 = prefetch its cache lines into L1 (no registers)
#ifndef RRV_EPI_PREFETCH_RR
#define RRV_EPI_PREFETCH_RR RRV_EPI_PREFETCH    // the same choice for the row-reuse (not merged-tap) instantiations
#endif
#ifndef RRV_EPI_HALF16
#define RRV_EPI_HALF16 1        // merged-tap epilogue: combine the three taps 16 columns at a time (register pressure)
#endif
#ifndef RRV_EPI_RESVOL
#define RRV_EPI_RESVOL 1        // residual loads as volatile asm at the top of the chunk
#endif
#ifndef RRV_EPI_L2PF
#define RRV_EPI_L2PF 1          // epilogue warps prefetch the NEXT tile's residual lines into L2 (bit 0: row-reuse layers, bit 1: merged-tap layers)
#endif
#ifndef RRV_EPI_FOLD
#define RRV_EPI_FOLD 1          // norm stages as one FMA + clamps with pre-multiplied constants
#endif
// Measurement-only switches (tools/build_variants.sh; results are WRONG with any of them set): what would the epilogue of the
// merged-tap layers cost without its global stores / without the shared-memory constant loads / without the tap-combine shuffles?
#ifndef RRV_EXP_NOSTORE
#define RRV_EXP_NOSTORE 0
#endif
#ifndef RRV_EXP_NOTAB
#define RRV_EXP_NOTAB 0
#endif
#ifndef RRV_EXP_NOSHFL
#define RRV_EXP_NOSHFL 0
#endif
#ifndef RRV_EXP_NORES
#define RRV_EXP_NORES 0
#endif

namespace rrv {

namespace {

constexpr int BM = 128;          // output pixels per tile (= TMEM lanes)
constexpr int BK = 64;           // channels per k-step (128-byte rows)
constexpr int MAX_STAGES = 8;       // A-operand ring slots (barrier arrays)
constexpr int EPI_WARPS = 8;      // two per TMEM lane quadrant, alternating CW-column chunks
constexpr int CW = 32;            // epilogue chunk width (accumulator columns per tcgen05.ld)
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TAB_BYTES = 48;     // per-channel epilogue constants: 11 floats (+1 pad)
constexpr int SMEM_LIMIT = 227 * 1024 - 1024;   // dynamic part: the opt-in maximum minus the static barriers

struct TcTune {
    int max_bn = 256;
    int mt = 2;             // M tiles (128 pixels each) per weight tile
    int dxm = 1;            // merge the three dx taps along N when 3 Cout_pad <= 256 (the 64-channel layers, the RGB head)
    int pair = 1;           // v2: CTA pairs (cta_group::2) for Cout tiles >= pair_min_bn
    int pair_min_bn = 64;   // (<= 64 also overrides resident weights: measured faster on the 64 -> 64 layers)
    int pdl = 1;            // programmatic dependent launch: the next layer's CTAs start (and wait) while this layer's last round runs
};
TcTune g_tune;                  // written by the rrv_tc_tune* setters, read as ONE snapshot per convolution call (both under g_tune_mu):
std::mutex g_tune_mu;           // a call never sees half of a setter's update, whatever thread flips the knobs
TcTune tune_snapshot() {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    return g_tune;
}
// rrv_tc_timeline (measurement only): every tensor-core convolution launched while a buffer is set takes the next 4-word slot
unsigned long long* g_tl_base = nullptr;
int g_tl_slots = 0, g_tl_next = 0;

struct OutDesc {
    int H, W, Cout, out_mode, out_C;     // H, W: the OUTPUT tensor (half the convolution's size when pool is set)
    int pool;                            // 2x2/2 max-pool of the epilogue result (vgg19.features[4|9|18]) before the store
    uint16_t* out_hi;
    uint16_t* out_lo;
    float* out_f32;
    void* out_img;                       // RRV_OUT_BGR_*: the post-processed, cropped HWC BGR frame
    int crop_y0, crop_x0, crop_h, crop_w;
};

// Shared-memory table of the per-channel constants of the fused pointwise chain, one array of
// Cout_pad floats per constant (absent stages get their identity values).
enum { T_BIAS = 0, T_M1, T_R1, T_LO1, T_HI1, T_M2, T_R2, T_LO2, T_HI2, T_SCALE, T_SHIFT, T_COUNT };
enum { EPI_N1 = 1, EPI_RES = 2, EPI_N2 = 4, EPI_AFF = 8, EPI_STATS = 16, EPI_HEAD = 32 };

__device__ __forceinline__ void fill_epilogue_table(float* s_tab, const EpiDev& e, int Cout, int Cout_pad, int nthreads) {
    for (int ch = threadIdx.x; ch < Cout_pad; ch += nthreads) {
        const bool in = ch < Cout;
        const int C = Cout;
        const float m1 = (in && e.norm1) ? e.norm1[ch] : 0.0f, r1 = (in && e.norm1) ? e.norm1[C + ch] : 1.0f;
        const float lo1 = (in && e.norm1) ? e.norm1[2 * C + ch] : -CUDART_INF_F, hi1 = (in && e.norm1) ? e.norm1[3 * C + ch] : CUDART_INF_F;
        const float m2 = (in && e.norm2) ? e.norm2[ch] : 0.0f, r2 = (in && e.norm2) ? e.norm2[C + ch] : 1.0f;
        const float lo2 = (in && e.norm2) ? e.norm2[2 * C + ch] : -CUDART_INF_F, hi2 = (in && e.norm2) ? e.norm2[3 * C + ch] : CUDART_INF_F;
        const float sc = (in && e.affine) ? e.affine[ch] : 1.0f, sh = (in && e.affine) ? e.affine[C + ch] : 0.0f;
        s_tab[T_BIAS * Cout_pad + ch] = (in && e.bias) ? e.bias[ch] : 0.0f;
#if RRV_EPI_FOLD
        // (x - m) r            ->  fma(x, r, -m r)
        // ((x - m) r) s + t    ->  fma(x, r s, t - m r s), the clamp bounds mapped through the same affine (s = a std > 0)
        s_tab[T_M1 * Cout_pad + ch] = -m1 * r1;
        s_tab[T_R1 * Cout_pad + ch] = r1;
        s_tab[T_LO1 * Cout_pad + ch] = lo1;
        s_tab[T_HI1 * Cout_pad + ch] = hi1;
        // a negative scale (never produced by cal_mean_std, std > 0) swaps the bounds; a zero scale makes the output the shift
        s_tab[T_M2 * Cout_pad + ch] = fmaf(-m2 * r2, sc, sh);
        s_tab[T_R2 * Cout_pad + ch] = r2 * sc;
        s_tab[T_LO2 * Cout_pad + ch] = sc == 0.0f ? -CUDART_INF_F : fmaf(sc > 0.0f ? lo2 : hi2, sc, sh);
        s_tab[T_HI2 * Cout_pad + ch] = sc == 0.0f ? CUDART_INF_F : fmaf(sc > 0.0f ? hi2 : lo2, sc, sh);
        s_tab[T_SCALE * Cout_pad + ch] = 1.0f;
        s_tab[T_SHIFT * Cout_pad + ch] = 0.0f;
#else
        s_tab[T_M1 * Cout_pad + ch] = m1;
        s_tab[T_R1 * Cout_pad + ch] = r1;
        s_tab[T_LO1 * Cout_pad + ch] = lo1;
        s_tab[T_HI1 * Cout_pad + ch] = hi1;
        s_tab[T_M2 * Cout_pad + ch] = m2;
        s_tab[T_R2 * Cout_pad + ch] = r2;
        s_tab[T_LO2 * Cout_pad + ch] = lo2;
        s_tab[T_HI2 * Cout_pad + ch] = hi2;
        s_tab[T_SCALE * Cout_pad + ch] = sc;
        s_tab[T_SHIFT * Cout_pad + ch] = sh;
#endif
    }
}

__device__ __forceinline__ void lds8(const float* p, float* v) {
#if RRV_EXP_NOTAB
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.75f;
    return;
#endif
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 8 fp32 values -> 8 bf16 hi + 8 bf16 lo (= bf16(v - hi)), two values per conversion instruction.
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
        const float r0 = v[2 * i] - __uint_as_float(hw[i] << 16);
        const float r1 = v[2 * i + 1] - __uint_as_float(hw[i] & 0xffff0000u);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// Where one accumulator row (= one output pixel) goes, computed once per (tile, M tile) by each epilogue thread.
struct PixCtx {
    long long out_off;      // element offset of the pixel's channel 0 in an NHWC output: ((n H + oy) W + ox) Cout
    long long res_off;      // same for the residual tensor (half resolution when res_shift = 1)
    int n, oy, ox;
    bool valid;
};

__device__ __forceinline__ PixCtx make_pix(const OutDesc& o, const EpiDev& e, int n, int oy, int ox, bool valid) {
    PixCtx px;
    px.n = n; px.oy = oy; px.ox = ox; px.valid = valid;
    px.out_off = (((long long)n * o.H + oy) * o.W + ox) * o.Cout;
    px.res_off = e.res_hi ? (long long)n * e.res_batch_stride + ((long long)(oy >> e.res_shift) * e.res_W + (ox >> e.res_shift)) * e.C : 0;
    return px;
}

// ALL: the NCHW and finished-frame outputs too (the FLAGS < 0 instantiations; the specialised ones only write planes / NHWC, and
// keeping the frame code out of them keeps the full-chain kernel below the register ceiling).
template <bool ALL>
__device__ __forceinline__ void store_group(const OutDesc& p, const float* v, const PixCtx& px, int c0, int nvalid) {
    if (p.out_mode == RRV_OUT_PLANES) {
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(p.out_hi + px.out_off + c0) = hi;
        if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + px.out_off + c0) = lo;
    } else if (p.out_mode == RRV_OUT_F32_NHWC) {
        float* o = p.out_f32 + px.out_off + c0;
        if (nvalid == 8) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nvalid) o[k] = v[k];
        }
    } else if (!ALL) {
        return;
    } else if (p.out_mode == RRV_OUT_F32_NCHW) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nvalid && c0 + k < p.out_C)
                p.out_f32[(((long long)px.n * p.out_C + c0 + k) * p.H + px.oy) * p.W + px.ox] = v[k];
    } else {
        // the RGB head with transform_back_image + tensor2numpy (test/framework.py:39-49) and the crop of
        // generate_real_video.py:167 fused in: img * std + mean (two roundings like the reference's tensor ops), clamp(0, 1),
        // * 255, RGB -> BGR, HWC
        if (c0 != 0) return;
        const int yy = px.oy - p.crop_y0, xx = px.ox - p.crop_x0;
        if (yy < 0 || yy >= p.crop_h || xx < 0 || xx >= p.crop_w) return;
        const float r = __fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(v[0], 0.229f), 0.485f), 0.0f), 1.0f), 255.0f);
        const float g = __fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(v[1], 0.224f), 0.456f), 0.0f), 1.0f), 255.0f);
        const float b = __fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(v[2], 0.225f), 0.406f), 0.0f), 1.0f), 255.0f);
        const long long off = (((long long)px.n * p.crop_h + yy) * p.crop_w + xx) * 3;
        if (p.out_mode == RRV_OUT_BGR_F32) {
            float* o = reinterpret_cast<float*>(p.out_img) + off;
            o[0] = b; o[1] = g; o[2] = r;
        } else {
            uint8_t* o = reinterpret_cast<uint8_t*>(p.out_img) + off;       // cv2.imwrite: saturate_cast<uchar>(cvRound(v))
            o[0] = (uint8_t)__float2int_rn(b); o[1] = (uint8_t)__float2int_rn(g); o[2] = (uint8_t)__float2int_rn(r);
        }
    }
}

// The fused pointwise chain on 8 consecutive channels starting at c0: bias -> act -> saved-stat norm ->
// + residual (rh/rl = the 8 hi / lo residual halves) -> saved-stat norm -> AdaIN.
template <int FLAGS>
__device__ __forceinline__ void chain8(const EpiDev& e, const float* s_tab, int ts, int c0, const uint32_t* acc, float* x,
                                       const uint4& rh, const uint4& rl, bool res_packed, const PixCtx& px, int Cout) {
    const bool has_n1 = FLAGS >= 0 ? (FLAGS & EPI_N1) != 0 : e.norm1 != nullptr;
    const bool has_res = FLAGS >= 0 ? (FLAGS & EPI_RES) != 0 : e.res_hi != nullptr;
    const bool has_n2 = FLAGS >= 0 ? (FLAGS & EPI_N2) != 0 : e.norm2 != nullptr;
    const bool has_aff = FLAGS >= 0 ? (FLAGS & EPI_AFF) != 0 : e.affine != nullptr;
    const bool res_x3 = e.res_lo != nullptr;
    float k0[8], k1[8];
    lds8(s_tab + T_BIAS * ts + c0, k0);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(acc[i]) + k0[i];
    if (e.act == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.0f);
    } else if (e.act == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.2f * x[i]);          // LeakyReLU(0.2): the slope is < 1
    }
    if (has_n1) {
        lds8(s_tab + T_M1 * ts + c0, k0);
        lds8(s_tab + T_R1 * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = RRV_EPI_FOLD ? fmaf(x[i], k1[i], k0[i]) : (x[i] - k0[i]) * k1[i];
        lds8(s_tab + T_LO1 * ts + c0, k0);
        lds8(s_tab + T_HI1 * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fminf(k1[i], fmaxf(k0[i], x[i]));
    }
    if (has_res) {
        if (res_packed && e.res_f32) {        // fp32 residual: rh = channels 0..3, rl = channels 4..7
            x[0] += __uint_as_float(rh.x); x[1] += __uint_as_float(rh.y); x[2] += __uint_as_float(rh.z); x[3] += __uint_as_float(rh.w);
            x[4] += __uint_as_float(rl.x); x[5] += __uint_as_float(rl.y); x[6] += __uint_as_float(rl.z); x[7] += __uint_as_float(rl.w);
        } else if (res_packed) {
            const uint32_t hw[4] = {rh.x, rh.y, rh.z, rh.w};
            const uint32_t lw[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = __uint_as_float(hw[i] << 16), b = __uint_as_float(hw[i] & 0xffff0000u);
                if (res_x3) {
                    a += __uint_as_float(lw[i] << 16);
                    b += __uint_as_float(lw[i] & 0xffff0000u);
                }
                x[2 * i] += a;
                x[2 * i + 1] += b;
            }
        } else {                              // ragged channel tail (never on the per-frame path)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cc = c0 + i;
                float rr = 0.0f;
                if (cc < Cout) rr = e.res_f32 ? reinterpret_cast<const float*>(e.res_hi)[px.res_off + cc] : bf16_to_f32(e.res_hi[px.res_off + cc]);
                if (res_x3 && cc < Cout) rr += bf16_to_f32(e.res_lo[px.res_off + cc]);
                x[i] += rr;
            }
        }
    }
#if RRV_EPI_FOLD
    if (has_n2 || has_aff) {              // Decoder.norm[i] and AdaIN as one FMA; the clamp bounds went through the same affine
        lds8(s_tab + T_M2 * ts + c0, k0);
        lds8(s_tab + T_R2 * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], k1[i], k0[i]);
        if (has_n2) {
            lds8(s_tab + T_LO2 * ts + c0, k0);
            lds8(s_tab + T_HI2 * ts + c0, k1);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fminf(k1[i], fmaxf(k0[i], x[i]));
        }
    }
#else
    if (has_n2) {
        lds8(s_tab + T_M2 * ts + c0, k0);
        lds8(s_tab + T_R2 * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (x[i] - k0[i]) * k1[i];
        lds8(s_tab + T_LO2 * ts + c0, k0);
        lds8(s_tab + T_HI2 * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fminf(k1[i], fmaxf(k0[i], x[i]));
    }
    if (has_aff) {
        lds8(s_tab + T_SCALE * ts + c0, k0);
        lds8(s_tab + T_SHIFT * ts + c0, k1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = x[i] * k0[i] + k1[i];
    }
#endif
}

// ---- statistics fused into the epilogue (rrv_conv.stats) ------------------------------------------------------------------
// A warp holds 32 pixels (lanes) x 32 channels (registers).  The butterfly below leaves lane L with the reduction over the 32
// pixels of channel L in 31 shuffles: at distance d a lane keeps the half of its values its side of the exchange owns and
// receives the partner's copy of that half.
template <class Op>
__device__ __forceinline__ float transpose_reduce32(const float* v, int lane, Op op) {
    float a16[16], a8[8], a4[4], a2[2];
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) a16[i] = op(up ? v[i + 16] : v[i], __shfl_xor_sync(0xffffffffu, up ? v[i] : v[i + 16], 16));
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] = op(up ? a16[i + 8] : a16[i], __shfl_xor_sync(0xffffffffu, up ? a16[i] : a16[i + 8], 8));
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) a4[i] = op(up ? a8[i + 4] : a8[i], __shfl_xor_sync(0xffffffffu, up ? a8[i] : a8[i + 4], 4));
    }
    {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) a2[i] = op(up ? a4[i + 2] : a4[i], __shfl_xor_sync(0xffffffffu, up ? a4[i] : a4[i + 2], 2));
    }
    const bool up = (lane & 1) != 0;
    return op(up ? a2[1] : a2[0], __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 1));
}

constexpr int STAT_SLOTS = 4;          // 32-channel chunks one epilogue warp can own: BN <= 256 over two warps per quadrant
struct StatAcc {                       // per lane: channel (chunk base + lane) of each owned chunk
    double s[STAT_SLOTS], q[STAT_SLOTS];
    float mn[STAT_SLOTS], mx[STAT_SLOTS];
};

template <int SLOT>
__device__ __forceinline__ void stat_add(StatAcc& a, const float* xs, bool valid, int lane, bool minmax) {
    float t[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) t[i] = valid ? xs[i] : 0.0f;
    a.s[SLOT] += (double)transpose_reduce32(t, lane, [](float x, float y) { return x + y; });
    if (minmax) {
#pragma unroll
        for (int i = 0; i < CW; ++i) t[i] = valid ? xs[i] : CUDART_INF_F;
        a.mn[SLOT] = fminf(a.mn[SLOT], transpose_reduce32(t, lane, [](float x, float y) { return fminf(x, y); }));
#pragma unroll
        for (int i = 0; i < CW; ++i) t[i] = valid ? xs[i] : -CUDART_INF_F;
        a.mx[SLOT] = fmaxf(a.mx[SLOT], transpose_reduce32(t, lane, [](float x, float y) { return fmaxf(x, y); }));
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) t[i] = valid ? xs[i] * xs[i] : 0.0f;
    a.q[SLOT] += (double)transpose_reduce32(t, lane, [](float x, float y) { return x + y; });
}

template <int SLOT>
__device__ __forceinline__ void stat_flush(StatAcc& a, double* stats, int C, int ch, bool minmax) {
    if (ch < C) {
        atomicAdd(stats + C + ch, a.s[SLOT]);
        atomicAdd(stats + 2 * C + ch, a.q[SLOT]);
        if (minmax) {
            atomic_min_double(stats + 3 * C + ch, (double)a.mn[SLOT]);
            atomic_max_double(stats + 4 * C + ch, (double)a.mx[SLOT]);
        }
    }
    a.s[SLOT] = 0.0; a.q[SLOT] = 0.0; a.mn[SLOT] = CUDART_INF_F; a.mx[SLOT] = -CUDART_INF_F;
}

// One CW-channel chunk of one accumulator row (= one output pixel): tcgen05.ld, the fused chain, the stores.
// cb = first global output channel of the chunk, col0 = its first column inside the Cout tile.
// FLAGS >= 0 fixes the set of stages at compile time (EPI_* bits); FLAGS < 0 reads it from `e`.
// DXM: the accumulator holds the three dx taps side by side (columns [0,Cp) [Cp,2Cp) [2Cp,3Cp), Cp = ts) for INPUT
// column = lane; output column `lane` = tap0[lane] + tap1[lane+1] + tap2[lane+2] (the lanes of one quadrant are
// 32 consecutive columns of one image row, the first of them one left of the tile).
// MODE 2 (nearest-x2 convolution, column phases merged): the chunk's two column taps b = 0, 1 sit ts columns apart; lane l is
// input column x0 - 1 + l and the output column 2 (x0 - 1 + l) + upx is tap0[l - 1] + tap1[l] (upx = 0) or tap0[l] + tap1[l + 1].
template <int FLAGS, int MODE, bool ST = false>
__device__ __forceinline__ void epilogue_chunk(const OutDesc& o, const EpiDev& e, const float* s_tab, int ts, uint32_t taddr,
                                               const PixCtx& px, int cb, int col0, int BN, bool pre = false,
                                               const uint4* pre_rh = nullptr, const uint4* pre_rl = nullptr, int upx = 0,
                                               float* xs = nullptr, uint32_t stage = 0u, uint32_t stage_lo = 0u) {
    // (staged stores: the caller has NOT yet waited for this warp's previous TMA store; that wait sits right before the first
    //  st.shared below, behind the TMEM loads and the tap combine, so the store's shared-memory read overlaps them)
    // stage != 0 (merged-tap layers, planes output): the finished 16-byte groups go to this warp's shared-memory staging rows
    // instead of global memory; the warp's elected lane then writes the whole row with one TMA store per plane.  A per-lane
    // 16-byte global store at a 128-byte pixel stride costs the LSU one tag per lane (32 per instruction): measured ~0.1 ms of a
    // 0.57 ms full-resolution 64-channel layer (tools/build_variants.sh nostore).
    constexpr bool DXM = MODE == 1;
    constexpr bool STATS = FLAGS >= 0 && (FLAGS & EPI_STATS) != 0;       // xs receives the CW finished values (0 beyond Cout)
    const bool has_res = FLAGS >= 0 ? (FLAGS & EPI_RES) != 0 : e.res_hi != nullptr;
    // the residual of the whole chunk goes in flight first, ahead of the TMEM loads and the tap combine (volatile loads: the
    // compiler would otherwise sink them to their first use to save registers, and the chain then waits for L2)
    uint4 rh[CW / 8], rl[CW / 8];
    const bool full = ST || (cb + CW <= o.Cout && col0 + CW <= BN);        // warp-uniform (staged instantiations: guaranteed by the host)
    if (pre) {                                                     // loaded by the caller while the MMAs were still running
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
            rh[g] = pre_rh[g];
            rl[g] = pre_rl[g];
        }
    } else if (has_res && px.valid && full) {
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
#if RRV_EXP_NORES
            rh[g] = make_uint4(0, 0, 0, 0);
            rl[g] = make_uint4(0, 0, 0, 0);
            continue;
#endif
#if RRV_EPI_RESVOL
            if (e.res_f32) {
                const float* rf = reinterpret_cast<const float*>(e.res_hi) + px.res_off + cb + g * 8;
                rh[g] = ptx::ldg_nc_v4(rf);
                rl[g] = ptx::ldg_nc_v4(rf + 4);
            } else {
                rh[g] = ptx::ldg_nc_v4(e.res_hi + px.res_off + cb + g * 8);
                if (e.res_lo) rl[g] = ptx::ldg_nc_v4(e.res_lo + px.res_off + cb + g * 8);
            }
#else
            rh[g] = *reinterpret_cast<const uint4*>(e.res_hi + px.res_off + cb + g * 8);
            if (e.res_lo) rl[g] = *reinterpret_cast<const uint4*>(e.res_lo + px.res_off + cb + g * 8);
#endif
        }
    }
    uint32_t r[CW];
    if (MODE == 2) {
#pragma unroll
        for (int h = 0; h < CW / 16; ++h) {
            uint32_t a0[16], a1[16];
            ptx::tmem_ld16_issue(taddr + (uint32_t)(16 * h), a0);
            ptx::tmem_ld16_issue(taddr + (uint32_t)(ts + 16 * h), a1);
            ptx::tmem_ld16_wait(a0);
            ptx::tmem_ld16_wait(a1);
            if (upx == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    r[16 * h + i] = __float_as_uint(__shfl_up_sync(0xffffffffu, __uint_as_float(a0[i]), 1) + __uint_as_float(a1[i]));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    r[16 * h + i] = __float_as_uint(__uint_as_float(a0[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(a1[i]), 1));
            }
        }
    } else if (DXM && RRV_EPI_HALF16) {
#pragma unroll
        for (int h = 0; h < CW / 16; ++h) {
            uint32_t a0[16], a1[16], a2[16];
            ptx::tmem_ld16_issue(taddr + (uint32_t)(16 * h), a0);
            ptx::tmem_ld16_issue(taddr + (uint32_t)(ts + 16 * h), a1);
            ptx::tmem_ld16_issue(taddr + (uint32_t)(2 * ts + 16 * h), a2);
            ptx::tmem_ld16_wait(a0);
            ptx::tmem_ld16_wait(a1);
            ptx::tmem_ld16_wait(a2);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float a = __uint_as_float(a0[i]);
#if RRV_EXP_NOSHFL
                const float b = __uint_as_float(a1[i]), c = __uint_as_float(a2[i]);
#else
                const float b = __shfl_down_sync(0xffffffffu, __uint_as_float(a1[i]), 1);
                const float c = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[i]), 2);
#endif
                r[16 * h + i] = __float_as_uint((a + b) + c);
            }
        }
    } else if (MODE != 2) {
        ptx::tmem_ld32_issue(taddr, r);
    }
    if (DXM && !RRV_EPI_HALF16) {
        uint32_t r1[CW], r2[CW];
        ptx::tmem_ld32_issue(taddr + (uint32_t)ts, r1);
        ptx::tmem_ld32_issue(taddr + (uint32_t)(2 * ts), r2);
        ptx::tmem_ld32_wait(r);
        ptx::tmem_ld32_wait(r1);
        ptx::tmem_ld32_wait(r2);
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            const float a = __uint_as_float(r[i]);
            const float b = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[i]), 1);
            const float c = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[i]), 2);
            r[i] = __float_as_uint((a + b) + c);
        }
    }
    if (MODE == 0) ptx::tmem_ld32_wait(r);
    if (ST || ((FLAGS == 0 || STATS) && stage != 0u)) {       // warp-uniform: before any lane leaves
        if ((threadIdx.x & 31) == 0) ptx::bulk_wait_read0();
        __syncwarp();
    }
    if (!STATS && !px.valid) return;       // (with statistics every lane stays for the warp-wide reduction that follows)
    if (ST || (full && o.out_mode == RRV_OUT_PLANES)) {
        // fast path (every per-frame layer but the RGB head): no per-group range checks, planes output
        uint16_t* oh = o.out_hi + px.out_off + cb;
        uint16_t* ol = o.out_lo ? o.out_lo + px.out_off + cb : nullptr;
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
            float x[8];
            chain8<FLAGS>(e, s_tab, ts, cb + g * 8, r + g * 8, x, rh[g], rl[g], true, px, o.Cout);
            if (STATS) {
#pragma unroll
                for (int k = 0; k < 8; ++k) xs[g * 8 + k] = x[k];
                if (!px.valid) continue;
            }
            uint4 hi, lo;
            split8(x, hi, lo);
#if RRV_EXP_NOSTORE
            if (hi.x != 0x7fc07fc0u || lo.y != 0x7fc17fc1u) continue;       // (never true in practice: the values stay live, nothing is stored)
#endif
            if (ST || (FLAGS == 0 && stage != 0u)) {      // ST: only ever staged; FLAGS 0: decided per launch; generic (FLAGS < 0): never
                // rows of 64 bytes (the chunk's 32 channels of one pixel), SWIZZLE_64B: 16-byte piece ^= address bits 7..8;
                // row = lane (MODE 1) or lane - 1 (MODE 2: lanes 1..30 are the tile's columns)
                const int row = (int)(threadIdx.x & 31) - (MODE == 2 ? 1 : 0);
                const uint32_t off = (uint32_t)(row * 64 + ((g ^ ((row >> 1) & 3)) << 4));
                ptx::sts_v4(stage + off, hi);
                if (ol) ptx::sts_v4(stage_lo + off, lo);
                continue;
            }
            if (ST) continue;
            *reinterpret_cast<uint4*>(oh + g * 8) = hi;
            if (ol) *reinterpret_cast<uint4*>(ol + g * 8) = lo;
        }
        return;
    }
#pragma unroll
    for (int g = 0; g < CW / 8; ++g) {
        const int c0 = cb + g * 8;
        if (c0 >= o.Cout || col0 + g * 8 >= BN) {
            if (STATS) {
#pragma unroll
                for (int k = 0; k < 8; ++k) xs[g * 8 + k] = 0.0f;
            }
            continue;
        }
        float x[8];
        chain8<FLAGS>(e, s_tab, ts, c0, r + g * 8, x, rh[g], rl[g], full, px, o.Cout);
        if (STATS) {
#pragma unroll
            for (int k = 0; k < 8; ++k) xs[g * 8 + k] = x[k];
            if (!px.valid) continue;
            if (stage != 0u) {
                // fp32 NHWC through the staging rows (frame mode / pre-pass): rows of 128 bytes = the chunk's 32 fp32 channels of one
                // pixel, SWIZZLE_128B (16-byte piece ^= address bits 7..9); row = lane (MODE 1) or lane - 1 (MODE 2)
                const int row = (int)(threadIdx.x & 31) - (MODE == 2 ? 1 : 0);
                const uint32_t base = stage + (uint32_t)(row * 128), key = (uint32_t)(row & 7);
                ptx::sts_v4(base + ((((uint32_t)(2 * g)) ^ key) << 4),
                            make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3])));
                ptx::sts_v4(base + ((((uint32_t)(2 * g + 1)) ^ key) << 4),
                            make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7])));
                continue;
            }
        }
        store_group<(FLAGS < 0)>(o, x, px, c0, c0 + 8 <= o.Cout ? 8 : o.Cout - c0);
    }
}

// Row-reuse layers whose epilogue is exposed (few MMAs per output value: the KernelFilter up-convolutions, conv2_1): a lane is a
// pixel, so its 16-byte loads / stores of a 512-channel tensor land in 32 different lines per instruction and the LSU, not the
// arithmetic, sets the pace (ncu: the epilogue warps wait on residual loads for 57 % of their samples, then on the load / store
// queue).  Here a warp's chunk -- 4 rows x 8 columns x 32 channels, 64 bytes per pixel and plane -- lives in one of two staging
// buffers (rows of 64 bytes, SWIZZLE_64B):
//   * the residual planes of the NEXT chunk are copied in by cp.async with a transposed lane map (four lanes cover one pixel's 64
//     bytes: 8 lines per instruction instead of 32) while this chunk is processed;
//   * the chain runs in place (each lane reads and rewrites its own row);
//   * the elected lane writes the finished rows with one TMA store per plane (box 32 channels x 8 x 4 pixels).
struct RrNext {                 // where the next chunk's residual comes from (per lane: column lane >> 2, 16-byte piece lane & 3)
    const uint16_t* hi;
    const uint16_t* lo;
    long long row_stride;       // elements between image rows
    int rows;                   // rows of the 4 that lie inside the image; 0: nothing to fetch
};

__device__ __forceinline__ uint32_t rr_piece_off(int row, int piece) { return (uint32_t)(row * 64 + ((piece ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ void rr_fetch(const RrNext& nx, uint32_t buf_hi, uint32_t buf_lo, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i < nx.rows) {
            const uint32_t off = rr_piece_off(8 * i + (lane >> 2), lane & 3);
            ptx::cp_async16(buf_hi + off, nx.hi + i * nx.row_stride);
            if (nx.lo) ptx::cp_async16(buf_lo + off, nx.lo + i * nx.row_stride);
        }
    }
}

template <int FLAGS>
__device__ __forceinline__ void epilogue_chunk_rr(const OutDesc& o, const EpiDev& e, const float* s_tab, int ts, uint32_t taddr,
                                                  const PixCtx& px, int cb, uint32_t buf_hi, uint32_t buf_lo, bool rs, const RrNext& nx,
                                                  uint32_t nbuf_hi, uint32_t nbuf_lo, int lane) {
    uint32_t r[CW];
    ptx::tmem_ld32_issue(taddr, r);
    ptx::tmem_ld32_wait(r);
    // the other buffer was the source of the TMA store issued one chunk ago: it has been read by now
    if (lane == 0) ptx::bulk_wait_read0();
    __syncwarp();
    if (rs) {
        rr_fetch(nx, nbuf_hi, nbuf_lo, lane);
        ptx::cp_async_commit();
        ptx::cp_async_wait1();             // this chunk's residual (committed one chunk ago) has landed
        __syncwarp();
    }
#pragma unroll
    for (int g = 0; g < CW / 8; ++g) {
        const uint32_t off = rr_piece_off(lane, g);
        uint4 rh = make_uint4(0, 0, 0, 0), rl = make_uint4(0, 0, 0, 0);
        if (rs) {
            rh = ptx::lds_v4(buf_hi + off);
            if (e.res_lo) rl = ptx::lds_v4(buf_lo + off);
        }
        float x[8];
        chain8<FLAGS>(e, s_tab, ts, cb + g * 8, r + g * 8, x, rh, rl, true, px, o.Cout);
        uint4 hi, lo;
        split8(x, hi, lo);
        if (px.valid) {
            ptx::sts_v4(buf_hi + off, hi);
            if (o.out_lo) ptx::sts_v4(buf_lo + off, lo);
        }
    }
}

// Pooled variant (Encoder conv1_2 / conv2_2 / conv3_4, whose only consumer is the 2x2 max-pool): the pool runs on
// the raw accumulators -- bias + ReLU are monotonic, so they commute with the max -- and only the pooled pixel goes
// through the chain and to memory (a quarter of the stores; the full-resolution tensor never exists).
//   !DXM: a warp's 32 lanes are 4 rows x 8 columns of the tile: both partners are lanes (xor 8, xor 1).
//    DXM: lanes are 32 consecutive columns of ONE row (quadrant = row): the column partner is lane ^ 1, the row
//         partner lives in the neighbouring warp (warp ^ 1) and is exchanged through `xbuf` (512 B per warp).
// `px` is the POOLED pixel (same for the 4 lanes of a 2x2 group, each of which stores one 8-channel group: `sub`).
template <bool DXM>
__device__ __forceinline__ void epilogue_chunk_pool(const OutDesc& o, const EpiDev& e, const float* s_tab, int ts, uint32_t taddr,
                                                    const PixCtx& px, int cb, int sub, float* xbuf, int warp, int lane) {
    uint32_t r[CW];
    ptx::tmem_ld32_issue(taddr, r);
    if (DXM) {
        uint32_t r1[CW], r2[CW];
        ptx::tmem_ld32_issue(taddr + (uint32_t)ts, r1);
        ptx::tmem_ld32_issue(taddr + (uint32_t)(2 * ts), r2);
        ptx::tmem_ld32_wait(r);
        ptx::tmem_ld32_wait(r1);
        ptx::tmem_ld32_wait(r2);
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            const float a = __uint_as_float(r[i]);
            const float b = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[i]), 1);
            const float c = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[i]), 2);
            const float v = (a + b) + c;
            r[i] = __float_as_uint(fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)));       // column partner
        }
    } else {
        ptx::tmem_ld32_wait(r);
#pragma unroll
        for (int i = 0; i < CW; ++i) {
            float v = __uint_as_float(r[i]);
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            r[i] = __float_as_uint(v);
        }
    }
    // this lane's 8-channel group
    uint32_t sel[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t lo2 = (sub & 1) ? r[8 + k] : r[k];
        const uint32_t hi2 = (sub & 1) ? r[24 + k] : r[16 + k];
        sel[k] = (sub & 2) ? hi2 : lo2;
    }
    if (DXM) {
        // The row partner (warp ^ 1) stores the groups with the other quadrant parity: publish that group of this lane
        // pair (`sub ^ 2`) and take the partner's copy of this lane's own group, 4 channels per round (512 B per warp).
        uint32_t pub[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t lo2 = (sub & 1) ? r[8 + k] : r[k];
            const uint32_t hi2 = (sub & 1) ? r[24 + k] : r[16 + k];
            pub[k] = (sub & 2) ? lo2 : hi2;
        }
        float4* mine = reinterpret_cast<float4*>(xbuf) + (warp - 2) * 32 + lane;
        const float4* theirs = reinterpret_cast<const float4*>(xbuf) + ((warp - 2) ^ 1) * 32 + lane;
        const int bar_id = 1 + ((warp - 2) >> 1);
#pragma unroll
        for (int rd = 0; rd < 2; ++rd) {
            *mine = make_float4(__uint_as_float(pub[4 * rd]), __uint_as_float(pub[4 * rd + 1]), __uint_as_float(pub[4 * rd + 2]),
                                __uint_as_float(pub[4 * rd + 3]));
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float4 q = *theirs;
            sel[4 * rd + 0] = __float_as_uint(fmaxf(__uint_as_float(sel[4 * rd + 0]), q.x));
            sel[4 * rd + 1] = __float_as_uint(fmaxf(__uint_as_float(sel[4 * rd + 1]), q.y));
            sel[4 * rd + 2] = __float_as_uint(fmaxf(__uint_as_float(sel[4 * rd + 2]), q.z));
            sel[4 * rd + 3] = __float_as_uint(fmaxf(__uint_as_float(sel[4 * rd + 3]), q.w));
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");       // the slot may be rewritten
        }
    }
    if (!px.valid) return;
    float x[8];
    const uint4 none = make_uint4(0, 0, 0, 0);
    chain8<0>(e, s_tab, ts, cb + sub * 8, sel, x, none, none, true, px, o.Cout);
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(o.out_hi + px.out_off + cb + sub * 8) = hi;
    if (o.out_lo) *reinterpret_cast<uint4*>(o.out_lo + px.out_off + cb + sub * 8) = lo;
}

// =====================================================================================================
// v2 main loop: fewer bytes through the L2 -> SM crossbar per MMA cycle (v1 is bound by it: ncu shows
// ~10.9 TB/s of xbar2l1tex traffic and the tensor pipe at 50% on the 256-channel layers).
//   * spatial tile = 16 rows x 8 columns per 128-row M tile, MT M tiles stacked vertically;
//   * ONE A box per (64-channel chunk, dx): (16 MT + 2) rows x 8 columns.  The three dy taps read
//     it at row offsets dy * 8 rows = dy * 1024 bytes, i.e. whole 128-byte-swizzle atoms, so the
//     operand descriptor only moves its start address: 3 loads instead of 9 per chunk;
//   * every weight tile (tap x chunk) is used by all MT M tiles before its slot is released
//     (B bytes per MMA halve with MT = 2), or stays resident for the whole kernel when all taps
//     fit (the 64 -> 64 full-resolution layers);
//   * the four phases of a nearest-x2 convolution share the three A boxes (16 -> 3 loads per chunk)
//     and accumulate into 4 MT separate TMEM accumulators.
// A and B travel through separate rings (A: a_stages, B: b_slots) filled by one producer thread in
// consumption order.
constexpr int MAX_B_SLOTS = 16;

struct Grp {                // one weight tile consumed against the current A box
    int btile;              // index into the [tap][Cout_pad][Cin] blob
    int arow;               // row offset (in 8-pixel rows) of the A operand inside the box
    int first;              // first contribution to the accumulators at chunk 0: overwrite
};

struct Tc2Params {
    OutDesc o;
    int N, in_H, in_W;
    int Cout, Cout_pad, kchunks;
    int nk_last;            // k-slices (of 16 channels) of the last chunk that meet non-zero weights (rrv_conv.Cin_used), 1..4
    int nph;                // 1, or 4 output phases of a nearest-x2 convolution (each phase is its own tile)
    int MT;
    int nA;                 // A boxes per chunk: 3 (3x3), 2 (one ups phase), 1 (1x1)
    int a_dx[4][3];         // [phase][box]: column offset of the box
    int a_y0[4];            // [phase]: row offset of the box origin (-1 for 3x3, py - 1 for ups, 0 for 1x1)
    int ngrp[4][3];
    Grp grp[4][3][3];
    int tiles_x, tiles_y, n_ntiles, total_tiles;
    int BN, x3;             // BN = N of the MMA (3 Cout_pad when the dy taps are merged); x3: lo planes present
    int terms;              // MMAs per k-slice beyond hi*Whi: bit 0 = hi*Wlo, bit 1 = lo*Whi (3 = the full fp32-accurate split)
    int BNe;                // output channels per tile (= BN, or Cout_pad when merged)
    int b_tile_rows;        // blob rows per weight tile index (Cout_pad, or 3 Cout_pad when merged)
    int dxm;                // 1: dx taps merged along N: M tile = 4 rows x 32 columns (30 outputs), one TMEM lane quadrant per row
                            // 2: nearest-x2 convolution with Cout <= 64: the two column phases and their two column taps merged
                            //    along N (N = 4 Cout_pad = [px0 b0 | px0 b1 | px1 b0 | px1 b1]); nph = 2 row phases are the tiles
    int dxm_groups;         // weight tiles per A box: 3 (tap rows dy) or 2 (tap rows a of one row phase)
    int b_parts;            // dxm 2: TMA loads per weight slot and CTA (blocks of Cout_pad blob rows)
    int a_stages, b_slots, b_resident, pair;
    int a_plane_bytes;      // (16 MT + 2) * 8 * row_bytes (16 MT for 1x1)
    int row_bytes;          // bytes per operand row: 128 (64-channel chunks, SWIZZLE_128B), or 64 for a 32-channel input (SWIZZLE_64B;
                            // row-reuse loop only: the KernelFilter up-convolutions read their 32 channels without zero padding)
    int acc_stride, set_stride, bufs, tmem_cols;
    double* stats;          // rrv_conv.stats: double[5][Cout] accumulated by the epilogue (EPI_STATS instantiations), or NULL
    int stats_minmax;
    int merge_wlo;          // the RGB head: hi * Whi and hi * Wlo as ONE MMA of N = 2 BN over the adjacent hi | lo weight planes (its
                            // MMAs cost their A fetch whatever N is); lo * Whi goes to a third column block; the epilogue adds the three
    int pdl_attr;           // launch with the programmatic-stream-serialization attribute (the tuning snapshot of this call)
    unsigned long long* tl; // rrv_tc_timeline: {first CTA start, first CTA past griddepcontrol.wait, first CTA end, last CTA end} in ns, or NULL
    int ostage;             // planes output through per-warp staging rows + TMA stores: bytes per plane and buffer (2048), 0 = direct stores
    int ostage_off;         // offset of the staging area (EPI_WARPS x ostage_bufs x 2 planes x ostage bytes) in dynamic shared memory
    int ostage_bufs;        // 1 (merged-tap layers), 2 (row-reuse layers: epilogue_chunk_rr alternates them)
    int rstage;             // row-reuse layers: the residual planes arrive in the staging rows too (cp.async, one chunk ahead)
    EpiDev ep;
};

// Tile index -> coordinates, identically in the producer, the MMA issuer and the epilogue.  mtc = M tiles this tile carries: an
// MT = 2 tile whose second M tile lies entirely below the image issues, loads and stores only the first one (304 rows = 9.5
// tiles of 32: the last tile row of the 256-channel layers at 1080p).
__device__ __forceinline__ void tile_coords(const Tc2Params& p, int tile, int xmul, int xoff, int cols, int rows, int& ph, int& n0,
                                            int& x0, int& y0, int& n, int& mtc) {
    int t = tile;
    ph = t % p.nph; t /= p.nph;
    n0 = (t % p.n_ntiles) * p.BNe; t /= p.n_ntiles;
    x0 = ((t % p.tiles_x) * xmul + xoff) * cols; t /= p.tiles_x;
    y0 = (t % p.tiles_y) * rows;
    n = t / p.tiles_y;
    mtc = (p.MT == 2 && y0 + 16 >= p.in_H) ? 1 : p.MT;
}

// PAIR: two CTAs of a cluster issue one cta_group::2 MMA per k-slice (256 pixels x BN): each CTA loads its own A box
// and HALF of the weight rows, so the per-MMA operand fetch drops from 64 + BN/2 to 64 + BN/4 cycles.
// DXM (dx taps merged along N, the 64-channel layers and the RGB head): everything about the tile is fixed -- one A box
// per chunk, three weight tiles (dy), one M tile, one Cout tile -- so the producer and the MMA issuer run straight-line
// loops in ONE thread each.  These layers have the smallest MMAs (N <= 192), and the generic loops (per-group elect /
// reconverge, parameter tables in constant memory, a barrier wait per resident weight tile) cost more issue cycles than the
// MMAs they launch take to execute (ncu: tensor pipe 43% active, issuer warp never waiting).
// (168 registers is the ceiling for this block: the register file is allocated per 4 warps, so 10 warps cost as much as 12.)
template <int FLAGS, bool PAIR, bool DXM>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                const Tc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_afull[MAX_STAGES], s_aempty[MAX_STAGES], s_bfull[MAX_B_SLOTS], s_bempty[MAX_B_SLOTS],
        s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;

    // Programmatic dependent launch: the next convolution of the frame may be scheduled as soon as SMs free up, so that its
    // prologue (barrier init, TMEM allocation, constants table, descriptor prefetch) overlaps this kernel's tail; it reads
    // and writes activations only after its own griddepcontrol.wait below, i.e. after this whole grid has completed.
    ptx::pdl_launch_dependents();
    if (p.tl != nullptr && threadIdx.x == 0) atomicMin(p.tl + 0, ptx::globaltimer());
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t planes = p.x3 ? 2u : 1u;
    const uint32_t a_stage_bytes = planes * (uint32_t)p.a_plane_bytes;
    const uint32_t rowb = DXM ? 128u : (uint32_t)p.row_bytes;
    const uint32_t desc_hi = rowb == 64u ? ptx::DESC_HI_SW64 : ptx::DESC_HI_SW128;
    const uint32_t b_plane_bytes = (uint32_t)(PAIR ? p.BN / 2 : p.BN) * rowb;    // a pair splits the weight rows
    const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;            // persistent worker index
    const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint32_t tx_mult = PAIR ? 2u : 1u;                                     // the leader's barriers count both CTAs' bytes
    const uint32_t b_slot_bytes = planes * b_plane_bytes;
    const uint32_t b_base = smem_base + (uint32_t)p.a_stages * a_stage_bytes;
    const uint32_t tab_off = (uint32_t)p.a_stages * a_stage_bytes + (uint32_t)p.b_slots * b_slot_bytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.a_stages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_afull[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_aempty[s]), 1);
        }
        for (int s = 0; s < p.b_slots; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_bfull[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_bempty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(ptx::smem_u32(&s_tfull[a]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_tempty[a]), EPI_WARPS * (PAIR ? 2 : 1));
        }
        ptx::fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a_hi);
        ptx::prefetch_tmap(&map_b_hi);
        if (p.x3) {
            ptx::prefetch_tmap(&map_a_lo);
            ptx::prefetch_tmap(&map_b_lo);
        }
    }
    if (p.ostage && warp == 2 && lane == 0) {
        ptx::prefetch_tmap(&map_o_hi);
        if (p.x3) ptx::prefetch_tmap(&map_o_lo);
    }
    float* s_tab = reinterpret_cast<float*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)) + tab_off);
    fill_epilogue_table(s_tab, p.ep, p.Cout, p.Cout_pad, TC_THREADS);
    if (warp == 1) {
        if (PAIR) {
            ptx::tmem_alloc_pair(ptx::smem_u32(&s_tmem_base), (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(ptx::smem_u32(&s_tmem_base), (uint32_t)p.tmem_cols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync();          // the peer's barriers are initialised before anyone signals them
    else __syncthreads();
    ptx::tc_fence_after();
    ptx::pdl_wait();                        // everything above ran while the previous kernel was still finishing
    if (p.tl != nullptr && threadIdx.x == 0) atomicMin(p.tl + 1, ptx::globaltimer());
    const uint32_t tmem_base = s_tmem_base;
    const int rows_per_set = DXM ? 4 : 16 * p.MT;
    const int cols_per_tile = DXM ? 30 : 8;                  // merged taps: 32 input columns give 30 output columns

    if (DXM && warp == 0) {
        // ================= TMA producer, merged-tap layers =================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            bool first_set = true;
            const int kch = p.kchunks;
            const uint32_t a_plane = (uint32_t)p.a_plane_bytes;
            const bool x3 = p.x3 != 0, resident = p.b_resident != 0, lead = !PAIR || cta_rank == 0;
            const int brow0 = PAIR ? (int)cta_rank * (p.BN / 2) : 0;
            for (int tile = cta_id; tile < p.total_tiles; tile += n_workers) {
                int t = tile;
                const int ph = t % p.nph; t /= p.nph;                        // dxm 2: the row phase py (fastest: the two phases of a
                                                                             // spatial tile follow each other, the second A read hits L2)
                const int bx = ((t % p.tiles_x) * (PAIR ? 2 : 1) + (int)cta_rank) * 30 - 1; t /= p.tiles_x;
                const int by = (t % p.tiles_y) * 4 + (p.dxm == 2 ? ph - 1 : -1);
                const int n = t / p.tiles_y;
                for (int kc = 0; kc < kch; ++kc) {
                    ptx::mbar_wait(ptx::smem_u32(&s_aempty[sa]), pa ^ 1u);
                    const uint32_t full = ptx::smem_u32(&s_afull[sa]);
                    const uint32_t dst = smem_base + (uint32_t)sa * a_stage_bytes;
                    if (lead) ptx::mbar_expect_tx(full, tx_mult * a_stage_bytes);
                    if (PAIR) {
                        ptx::tma_load_4d_pair(dst, &map_a_hi, full, kc * BK, bx, by, n);
                        if (x3) ptx::tma_load_4d_pair(dst + a_plane, &map_a_lo, full, kc * BK, bx, by, n);
                    } else {
                        ptx::tma_load_4d(dst, &map_a_hi, full, kc * BK, bx, by, n);
                        if (x3) ptx::tma_load_4d(dst + a_plane, &map_a_lo, full, kc * BK, bx, by, n);
                    }
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
                    if (resident && !first_set) continue;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        if (dy >= p.dxm_groups) break;
                        int slot;
                        if (resident) {
                            slot = dy * kch + kc;
                        } else {
                            slot = sb;
                            ptx::mbar_wait(ptx::smem_u32(&s_bempty[sb]), pb ^ 1u);
                            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
                        }
                        const uint32_t bfull = ptx::smem_u32(&s_bfull[slot]);
                        const uint32_t bdst = b_base + (uint32_t)slot * b_slot_bytes;
                        if (lead) ptx::mbar_expect_tx(bfull, tx_mult * b_slot_bytes);
                        if (p.dxm == 2) {
                            // weight tile of (row phase ph, tap row a = dy): blocks (px, b) of Cout_pad rows from the 16-matrix blob
                            // [(py 2 + px) 4 + a 2 + b]; a pair splits them by column phase (CTA r holds px = r)
                            for (int q = 0; q < p.b_parts; ++q) {
                                const int px = PAIR ? (int)cta_rank : (q >> 1), b = q & 1;
                                const int brow = (((ph * 2 + px) * 4) + dy * 2 + b) * p.Cout_pad;
                                const uint32_t dq = bdst + (uint32_t)(q * p.Cout_pad) * 128u;
                                if (PAIR) {
                                    ptx::tma_load_2d_pair(dq, &map_b_hi, bfull, kc * BK, brow);
                                    if (x3) ptx::tma_load_2d_pair(dq + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                                } else {
                                    ptx::tma_load_2d(dq, &map_b_hi, bfull, kc * BK, brow);
                                    if (x3) ptx::tma_load_2d(dq + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                                }
                            }
                            continue;
                        }
                        const int brow = dy * p.b_tile_rows + brow0;
                        if (PAIR) {
                            ptx::tma_load_2d_pair(bdst, &map_b_hi, bfull, kc * BK, brow);
                            if (x3) ptx::tma_load_2d_pair(bdst + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                        } else {
                            ptx::tma_load_2d(bdst, &map_b_hi, bfull, kc * BK, brow);
                            if (x3) ptx::tma_load_2d(bdst + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                        }
                    }
                }
                first_set = false;
            }
        }
    } else if (DXM && warp == 1 && cta_rank == 0) {
        // ================= MMA issuer, merged-tap layers: one thread, 36 (x3) MMAs per chunk back to back =================
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, p.BN);
            int sa = 0, sb = 0, as = 0;
            uint32_t pa = 0, pb = 0, aphase = 0;
            bool first_set = true;
            const int kch = p.kchunks;
            const uint32_t a_plane = (uint32_t)p.a_plane_bytes;
            const uint32_t x3 = (uint32_t)p.terms;
            const bool resident = p.b_resident != 0;
            for (int tile = cta_id; tile < p.total_tiles; tile += n_workers) {
                ptx::mbar_wait(ptx::smem_u32(&s_tempty[as]), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(as * p.set_stride);
                for (int kc = 0; kc < kch; ++kc) {
                    ptx::mbar_wait(ptx::smem_u32(&s_afull[sa]), pa);
                    ptx::tc_fence_after();
                    const uint32_t a_base = smem_base + (uint32_t)sa * a_stage_bytes;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        if (dy >= p.dxm_groups) break;
                        int slot;
                        if (resident) {
                            slot = dy * kch + kc;
                            if (first_set) {
                                ptx::mbar_wait(ptx::smem_u32(&s_bfull[slot]), 0u);
                                ptx::tc_fence_after();
                            }
                        } else {
                            slot = sb;
                            ptx::mbar_wait(ptx::smem_u32(&s_bfull[sb]), pb);
                            ptx::tc_fence_after();
                        }
                        const uint32_t bs = b_base + (uint32_t)slot * b_slot_bytes;
                        const uint32_t a0 = a_base + (uint32_t)dy * 4096u;          // tap row dy: 32 pixels x 128 bytes further down the box
                        const bool overwrite = kc == 0 && dy == 0;
                        const uint32_t nk = kc == kch - 1 ? (uint32_t)p.nk_last : 4u;
                        if (FLAGS == EPI_HEAD && p.merge_wlo) {
                            // A_hi x [Whi | Wlo] (the two weight planes of the slot are adjacent: one operand of 2 BN rows), then A_lo x Whi
                            const uint32_t idesc2 = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, 2 * p.BN);
                            const uint32_t ah = ptx::desc_lo_sw128(a0), al = ptx::desc_lo_sw128(a0 + a_plane), bh = ptx::desc_lo_sw128(bs);
#pragma unroll
                            for (uint32_t k = 0; k < 4; ++k) {
                                if (k >= nk) break;
                                const uint32_t acc = (overwrite && k == 0) ? 0u : 1u;
                                if (PAIR) {
                                    ptx::mma_bf16_lo_pair(d0, ah + 2 * k, bh + 2 * k, idesc2, acc);
                                    ptx::mma_bf16_lo_pair(d0 + 2u * (uint32_t)p.BN, al + 2 * k, bh + 2 * k, idesc, acc);
                                } else {
                                    ptx::mma_bf16_lo(d0, ah + 2 * k, bh + 2 * k, idesc2, acc);
                                    ptx::mma_bf16_lo(d0 + 2u * (uint32_t)p.BN, al + 2 * k, bh + 2 * k, idesc, acc);
                                }
                            }
                        } else if (PAIR) ptx::mma_kblock_pair(d0, a0, a0 + a_plane, bs, bs + b_plane_bytes, idesc, x3, overwrite, nk);
                        else ptx::mma_kblock(d0, a0, a0 + a_plane, bs, bs + b_plane_bytes, idesc, x3, overwrite, nk);
                        if (!resident) {
                            if (PAIR) ptx::mma_commit_pair(ptx::smem_u32(&s_bempty[sb]));
                            else ptx::mma_commit(ptx::smem_u32(&s_bempty[sb]));
                            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
                        }
                    }
                    if (PAIR) {
                        ptx::mma_commit_pair(ptx::smem_u32(&s_aempty[sa]));
                        if (kc == kch - 1) ptx::mma_commit_pair(ptx::smem_u32(&s_tfull[as]));
                    } else {
                        ptx::mma_commit(ptx::smem_u32(&s_aempty[sa]));
                        if (kc == kch - 1) ptx::mma_commit(ptx::smem_u32(&s_tfull[as]));
                    }
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
                }
                first_set = false;
                if (++as == p.bufs) { as = 0; aphase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            bool first_set = true;
            for (int tile = cta_id; tile < p.total_tiles; tile += n_workers) {
                int ph, n0, x0, y0, n, mtc;
                tile_coords(p, tile, PAIR ? 2 : 1, (int)cta_rank, cols_per_tile, rows_per_set, ph, n0, x0, y0, n, mtc);
                const int by = y0 + p.a_y0[ph];
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    for (int j = 0; j < p.nA; ++j) {
                        const int bx = x0 + p.a_dx[ph][j];
                        ptx::mbar_wait(ptx::smem_u32(&s_aempty[sa]), pa ^ 1u);
                        const uint32_t full = ptx::smem_u32(&s_afull[sa]);
                        const uint32_t dst = smem_base + (uint32_t)sa * a_stage_bytes;
                        if (!PAIR || cta_rank == 0) ptx::mbar_expect_tx(full, tx_mult * a_stage_bytes);
                        if (PAIR) {
                            ptx::tma_load_4d_pair(dst, &map_a_hi, full, kc * BK, bx, by, n);
                            if (p.x3) ptx::tma_load_4d_pair(dst + (uint32_t)p.a_plane_bytes, &map_a_lo, full, kc * BK, bx, by, n);
                        } else {
                            ptx::tma_load_4d(dst, &map_a_hi, full, kc * BK, bx, by, n);
                            if (p.x3) ptx::tma_load_4d(dst + (uint32_t)p.a_plane_bytes, &map_a_lo, full, kc * BK, bx, by, n);
                        }
                        if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
                        for (int g = 0; g < p.ngrp[ph][j]; ++g) {
                            const int btile = p.grp[ph][j][g].btile;
                            int slot;
                            if (p.b_resident) {
                                if (!first_set) continue;
                                slot = btile * p.kchunks + kc;
                            } else {
                                slot = sb;
                                ptx::mbar_wait(ptx::smem_u32(&s_bempty[sb]), pb ^ 1u);
                                if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
                            }
                            const uint32_t bfull = ptx::smem_u32(&s_bfull[slot]);
                            const uint32_t bdst = b_base + (uint32_t)slot * b_slot_bytes;
                            const int brow = btile * p.b_tile_rows + n0 + (PAIR ? (int)cta_rank * (p.BN / 2) : 0);
                            if (!PAIR || cta_rank == 0) ptx::mbar_expect_tx(bfull, tx_mult * b_slot_bytes);
                            if (PAIR) {
                                ptx::tma_load_2d_pair(bdst, &map_b_hi, bfull, kc * BK, brow);
                                if (p.x3) ptx::tma_load_2d_pair(bdst + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                            } else {
                                ptx::tma_load_2d(bdst, &map_b_hi, bfull, kc * BK, brow);
                                if (p.x3) ptx::tma_load_2d(bdst + b_plane_bytes, &map_b_lo, bfull, kc * BK, brow);
                            }
                        }
                    }
                }
                first_set = false;
            }
        }
    } else if (warp == 1 && cta_rank == 0) {
        // ================= MMA issuer (the leader CTA of a pair) =================
        const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, p.BN);
        int sa = 0, sb = 0, as = 0;
        uint32_t pa = 0, pb = 0, aphase = 0;
        for (int tile = cta_id; tile < p.total_tiles; tile += n_workers) {
            int ph, n0_, x0_, y0_, n_, mtc;
            tile_coords(p, tile, PAIR ? 2 : 1, 0, cols_per_tile, rows_per_set, ph, n0_, x0_, y0_, n_, mtc);
            ptx::mbar_wait(ptx::smem_u32(&s_tempty[as]), aphase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_set = tmem_base + (uint32_t)(as * p.set_stride);
            for (int kc = 0; kc < p.kchunks; ++kc) {
                for (int j = 0; j < p.nA; ++j) {
                    ptx::mbar_wait(ptx::smem_u32(&s_afull[sa]), pa);
                    const uint32_t a_base = smem_base + (uint32_t)sa * a_stage_bytes;
                    for (int g = 0; g < p.ngrp[ph][j]; ++g) {
                        const Grp gr = p.grp[ph][j][g];
                        int slot;
                        if (p.b_resident) {
                            slot = gr.btile * p.kchunks + kc;
                            ptx::mbar_wait(ptx::smem_u32(&s_bfull[slot]), 0u);
                        } else {
                            slot = sb;
                            ptx::mbar_wait(ptx::smem_u32(&s_bfull[sb]), pb);
                        }
                        ptx::tc_fence_after();
                        const uint32_t bs = b_base + (uint32_t)slot * b_slot_bytes;
                        const bool overwrite = kc == 0 && gr.first != 0;
                        const uint32_t a0 = a_base + (uint32_t)gr.arow * 8u * rowb;          // 8 pixels per tile row
                        const uint32_t a1 = a0 + 128u * rowb;                                // second M tile: 16 rows further down
                        const uint32_t d0 = d_set;
                        const uint32_t bempty_bar = ptx::smem_u32(&s_bempty[sb]);
                        const uint32_t nk = kc == p.kchunks - 1 ? (uint32_t)p.nk_last : 4u;
                        if (ptx::elect_one()) {
                            if (PAIR) {
                                ptx::mma_kblock_pair(d0, a0, a0 + (uint32_t)p.a_plane_bytes, bs, bs + b_plane_bytes, idesc, (uint32_t)p.terms, overwrite, nk,
                                                     desc_hi);
                                if (mtc == 2)
                                    ptx::mma_kblock_pair(d0 + (uint32_t)p.acc_stride, a1, a1 + (uint32_t)p.a_plane_bytes, bs,
                                                         bs + b_plane_bytes, idesc, (uint32_t)p.terms, overwrite, nk, desc_hi);
                                if (!p.b_resident) ptx::mma_commit_pair(bempty_bar);
                            } else {
                                ptx::mma_kblock(d0, a0, a0 + (uint32_t)p.a_plane_bytes, bs, bs + b_plane_bytes, idesc, (uint32_t)p.terms, overwrite, nk, desc_hi);
                                if (mtc == 2)
                                    ptx::mma_kblock(d0 + (uint32_t)p.acc_stride, a1, a1 + (uint32_t)p.a_plane_bytes, bs,
                                                    bs + b_plane_bytes, idesc, (uint32_t)p.terms, overwrite, nk, desc_hi);
                                if (!p.b_resident) ptx::mma_commit(bempty_bar);
                            }
                        }
                        __syncwarp();
                        if (!p.b_resident) {
                            if (++sb == p.b_slots) { sb = 0; pb ^= 1u; }
                        }
                    }
                    const uint32_t aempty_bar = ptx::smem_u32(&s_aempty[sa]), tfull_bar = ptx::smem_u32(&s_tfull[as]);
                    if (ptx::elect_one()) {
                        if (PAIR) {
                            ptx::mma_commit_pair(aempty_bar);
                            if (kc == p.kchunks - 1 && j == p.nA - 1) ptx::mma_commit_pair(tfull_bar);
                        } else {
                            ptx::mma_commit(aempty_bar);
                            if (kc == p.kchunks - 1 && j == p.nA - 1) ptx::mma_commit(tfull_bar);
                        }
                    }
                    __syncwarp();
                    if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
                }
            }
            if (++as == p.bufs) { as = 0; aphase ^= 1u; }
        }
    } else if (warp >= 2) {
        // ================= epilogue (warps 2..9; TMEM lane quadrant = warp % 4) =================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        const int ty = DXM ? quad : (m >> 3), tx = DXM ? lane : (m & 7);
        const int nchunks = (p.BNe + CW - 1) / CW;
        const EpiDev& e = p.ep;
        // the full-chain and norm1-only merged-tap instantiations (slice2.conv2, slice2.conv1) always store through the staging rows
        constexpr bool kStaged = DXM && (FLAGS == (EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF) || FLAGS == EPI_N1);
        // this warp's output staging rows (hi plane, then lo plane), 1024-byte aligned
        const uint32_t o_stage = p.ostage ? smem_base + (uint32_t)p.ostage_off + (uint32_t)((warp - 2) * 2 * p.ostage_bufs * p.ostage) : 0u;
        constexpr bool STATS = FLAGS >= 0 && (FLAGS & EPI_STATS) != 0;
        StatAcc sacc;
        if (STATS) {
#pragma unroll
            for (int j = 0; j < STAT_SLOTS; ++j) { sacc.s[j] = 0.0; sacc.q[j] = 0.0; sacc.mn[j] = CUDART_INF_F; sacc.mx[j] = -CUDART_INF_F; }
        }
        const bool st_mm = STATS && p.stats_minmax != 0;
        int st_n0 = 0;
        int as = 0;
        uint32_t aphase = 0;
        // staged row-reuse epilogue (epilogue_chunk_rr): where a chunk's residual comes from, per lane
        constexpr bool kRR = !DXM && (FLAGS == EPI_RES || FLAGS == (EPI_RES | EPI_N2 | EPI_AFF));
        const bool rr = kRR && p.ostage != 0;
        const bool rr_rs = rr && p.rstage != 0;
        int rr_par = 0;
        auto rr_next_of = [&](int tile_, int mt_, int ch_) {
            RrNext nx;
            nx.hi = nullptr; nx.lo = nullptr; nx.row_stride = 0; nx.rows = 0;
            if (!rr_rs || tile_ >= p.total_tiles) return nx;
            int ph_, n0_, x0_, y0_, n_, mtc_;
            tile_coords(p, tile_, PAIR ? 2 : 1, (int)cta_rank, cols_per_tile, rows_per_set, ph_, n0_, x0_, y0_, n_, mtc_);
            const int yq = y0_ + 16 * mt_ + quad * 4, xc = x0_ + (lane >> 2);
            if (xc >= p.in_W || yq >= p.in_H) return nx;
            nx.rows = p.in_H - yq < 4 ? p.in_H - yq : 4;
            nx.row_stride = (long long)e.res_W * e.C;
            const long long off = (long long)n_ * e.res_batch_stride + ((long long)yq * e.res_W + xc) * e.C + n0_ + ch_ * CW + (lane & 3) * 8;
            nx.hi = e.res_hi + off;
            nx.lo = e.res_lo ? e.res_lo + off : nullptr;
            return nx;
        };
        if (rr_rs) {                                  // the first chunk's residual
            if (half < nchunks) rr_fetch(rr_next_of(cta_id, 0, half), o_stage, o_stage + (uint32_t)p.ostage, lane);
            ptx::cp_async_commit();
        }
        for (int tile = cta_id; tile < p.total_tiles; tile += n_workers) {
            int ph, n0, x0, y0, n, mtc;
            tile_coords(p, tile, PAIR ? 2 : 1, (int)cta_rank, cols_per_tile, rows_per_set, ph, n0, x0, y0, n, mtc);
#if RRV_EPI_L2PF
            // The residual of a tile is first touched by this kernel (it comes from DRAM): ask L2 for the NEXT tile's lines now, a
            // whole tile time before the loads that need them (ncu: 57 % of the epilogue warps' stall samples of the KernelFilter
            // up-convolutions were long-scoreboard waits on exactly those loads).  One 128-byte line per plane and lane; the two warps
            // of a TMEM quadrant alternate lines.
            {
                constexpr bool kResS = FLAGS >= 0 && (FLAGS & EPI_RES) != 0;
                const bool res_any = FLAGS >= 0 ? kResS : e.res_hi != nullptr;
                const bool want = DXM ? (RRV_EPI_L2PF & 2) != 0 : (RRV_EPI_L2PF & 1) != 0;
                if (want && res_any && (!rr_rs || (RRV_EPI_L2PF & 4) != 0) && tile + n_workers < p.total_tiles) {
                    int ph2, n02, x02, y02, n2, mtc2;
                    tile_coords(p, tile + n_workers, PAIR ? 2 : 1, (int)cta_rank, cols_per_tile, rows_per_set, ph2, n02, x02, y02, n2, mtc2);
                    const int esz = e.res_f32 ? 4 : 2;
                    const int lines = (p.BNe * esz) >> 7;             // whole lines inside the pixel's channel range only
                    for (int mt = 0; mt < mtc2; ++mt) {
                        const int iy = y02 + (DXM ? 0 : 16 * mt) + ty, ix = x02 + tx;
                        bool v2 = iy < p.in_H && ix < p.in_W && (!DXM || tx < 30);
                        const int oy = p.nph == 4 ? 2 * iy + (ph2 >> 1) : iy;
                        const int ox = p.nph == 4 ? 2 * ix + (ph2 & 1) : ix;
                        if (e.res_shift) v2 = v2 && ((oy | ox) & 1) == 0;      // one lane per shared low-resolution pixel
                        if (!v2) continue;
                        const long long off = ((long long)n2 * e.res_batch_stride +
                                               ((long long)(oy >> e.res_shift) * e.res_W + (ox >> e.res_shift)) * e.C + n02) * esz;
                        const char* bh = reinterpret_cast<const char*>(e.res_hi) + off;
                        const char* bl = e.res_lo ? reinterpret_cast<const char*>(e.res_lo) + off : nullptr;
                        for (int l = half; l < lines; l += 2) {
                            ptx::prefetch_l2(bh + l * 128);
                            if (bl) ptx::prefetch_l2(bl + l * 128);
                        }
                    }
                }
            }
#endif
            // (the full-chain and statistics instantiations never run the merged-phase layout: the host sends those to the generic one)
            constexpr bool kMode2 = DXM && FLAGS != (EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF) && FLAGS != EPI_HEAD;
            if (kMode2 && p.dxm == 2) {
                // nearest-x2 convolution, column phases merged: warp (quad, half) = (tile row, column phase px); lanes 1..30 are the
                // tile's 30 low-resolution columns; output pixel (2 iy + py, 2 ix + px)
                const int iy = y0 + quad, ix = x0 + lane - 1;
                const bool valid = lane >= 1 && lane <= 30 && iy < p.in_H && ix < p.in_W;
                const PixCtx px = make_pix(p.o, e, n, 2 * iy + ph, 2 * ix + half, valid);
                ptx::mbar_wait(ptx::smem_u32(&s_tfull[as]), aphase);
                ptx::tc_fence_after();
                const uint32_t ta = tmem_base + (uint32_t)(as * p.set_stride) + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 2 * p.Cout_pad);
                if (STATS) {                         // (frame mode / pre-pass: fp32 NHWC output + the statistics of what is written)
#define RRV_STAT_CHUNK2(J)                                                                                                             \
    if ((J) < p.Cout_pad / CW) {                                                                                                       \
        float xs[CW];                                                                                                                  \
        epilogue_chunk<FLAGS, 2>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)((J) * CW), px, (J) * CW, (J) * CW, p.Cout_pad, false, nullptr, \
                                 nullptr, half, xs, o_stage);                                                                          \
        if (p.ostage) {                                                                                                                \
            ptx::fence_proxy_async();                                                                                                  \
            __syncwarp();                                                                                                              \
            if (lane == 0 && iy < p.in_H) {                                                                                            \
                ptx::tma_store_5d(&map_o_hi, o_stage, (J) * CW, half, x0, 2 * iy + ph, n);                                             \
                ptx::bulk_commit();                                                                                                    \
            }                                                                                                                          \
        }                                                                                                                              \
        stat_add<(J)>(sacc, xs, valid, lane, st_mm);                                                                                   \
    }
                    RRV_STAT_CHUNK2(0) RRV_STAT_CHUNK2(1)
#undef RRV_STAT_CHUNK2
                } else
                for (int c = 0; c < p.Cout_pad / CW; ++c) {
                    epilogue_chunk<FLAGS, 2, kStaged>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(c * CW), px, c * CW, c * CW, p.Cout_pad, false, nullptr,
                                                      nullptr, half, nullptr, o_stage, o_stage + (uint32_t)p.ostage);
                    if (kStaged || p.ostage) {
                        // 30 output pixels of one column phase of one output row x 32 channels per plane: tensor view
                        // (C, column parity, W / 2, H, N)
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && iy < p.in_H) {
                            ptx::tma_store_5d(&map_o_hi, o_stage, c * CW, half, x0, 2 * iy + ph, n);
                            if (p.o.out_lo) ptx::tma_store_5d(&map_o_lo, o_stage + (uint32_t)p.ostage, c * CW, half, x0, 2 * iy + ph, n);
                            ptx::bulk_commit();
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) ptx::mbar_arrive_cluster(ptx::smem_u32(&s_tempty[as]), 0u);
                    else ptx::mbar_arrive(ptx::smem_u32(&s_tempty[as]));
                }
                if (++as == p.bufs) { as = 0; aphase ^= 1u; }
                continue;
            }
            if constexpr (DXM && FLAGS == EPI_HEAD) {
                // the RGB head (Decoder.slice1, 64 -> 3, bias only): three useful accumulator columns per tap.  One warp per tile row
                // loads 4 columns of each tap, combines the taps of its 3 channels and writes the pixel (NCHW or the finished BGR
                // frame); nothing else of the generic chunk machinery runs.
                const int iy = y0 + quad, ix = x0 + lane;
                const bool valid = iy < p.in_H && ix < p.in_W && lane < 30;
                const PixCtx px = make_pix(p.o, e, n, iy, ix, valid);
                ptx::mbar_wait(ptx::smem_u32(&s_tfull[as]), aphase);
                ptx::tc_fence_after();
                if (half == 0) {
                    const uint32_t ta = tmem_base + (uint32_t)(as * p.set_stride) + ((uint32_t)(quad * 32) << 16);
                    uint32_t a0[4], a1[4], a2[4];
                    if (p.merge_wlo) {
                        // column blocks [hi * Whi | hi * Wlo] (a pair interleaves them per CTA: each holds half of the rows of both
                        // planes) and lo * Whi from column 2 BN on; row r of the weight tile = tap (r / Cout_pad), channel (r % Cout_pad)
                        const int hb = PAIR ? p.BN / 2 : p.BN;
                        uint32_t* acc3[3] = {a0, a1, a2};
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const int r0 = t * p.Cout_pad;
                            const int chi = PAIR ? (r0 < hb ? r0 : 2 * hb + (r0 - hb)) : r0;
                            uint32_t u[4], v[4];
                            ptx::tmem_ld4_issue(ta + (uint32_t)chi, acc3[t]);
                            ptx::tmem_ld4_issue(ta + (uint32_t)(chi + hb), u);
                            ptx::tmem_ld4_issue(ta + (uint32_t)(2 * p.BN + r0), v);
                            ptx::tmem_ld4_wait(acc3[t]);
                            ptx::tmem_ld4_wait(u);
                            ptx::tmem_ld4_wait(v);
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                acc3[t][c] = __float_as_uint((__uint_as_float(acc3[t][c]) + __uint_as_float(u[c])) + __uint_as_float(v[c]));
                        }
                    } else {
                    ptx::tmem_ld4_issue(ta, a0);
                    ptx::tmem_ld4_issue(ta + (uint32_t)p.Cout_pad, a1);
                    ptx::tmem_ld4_issue(ta + (uint32_t)(2 * p.Cout_pad), a2);
                    ptx::tmem_ld4_wait(a0);
                    ptx::tmem_ld4_wait(a1);
                    ptx::tmem_ld4_wait(a2);
                    }
                    float x[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[c] = 0.0f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float b1 = __shfl_down_sync(0xffffffffu, __uint_as_float(a1[c]), 1);
                        const float b2 = __shfl_down_sync(0xffffffffu, __uint_as_float(a2[c]), 2);
                        x[c] = ((__uint_as_float(a0[c]) + b1) + b2) + s_tab[T_BIAS * p.Cout_pad + c];
                    }
                    if (valid) store_group<true>(p.o, x, px, 0, p.o.Cout < 4 ? p.o.Cout : 4);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) ptx::mbar_arrive_cluster(ptx::smem_u32(&s_tempty[as]), 0u);
                    else ptx::mbar_arrive(ptx::smem_u32(&s_tempty[as]));
                }
                if (++as == p.bufs) { as = 0; aphase ^= 1u; }
                continue;
            }
            // the first chunk's output pixel and residual do not depend on the accumulators: fetch them while the MMAs run
            constexpr bool kHasResStatic = FLAGS >= 0 && (FLAGS & EPI_RES) != 0;
            const bool has_res = FLAGS >= 0 ? kHasResStatic : e.res_hi != nullptr;
            uint4 pre_rh[CW / 8], pre_rl[CW / 8];
            bool pre = false;
            constexpr int kPrefetch = DXM ? RRV_EPI_PREFETCH : RRV_EPI_PREFETCH_RR;
            if (kPrefetch && has_res && !RRV_EXP_NORES && !rr_rs) {
                const int iy = y0 + ty, ix = x0 + tx;
                const bool valid = iy < p.in_H && ix < p.in_W && (!DXM || tx < 30);
                const int oy = p.nph == 4 ? 2 * iy + (ph >> 1) : iy;
                const int ox = p.nph == 4 ? 2 * ix + (ph & 1) : ix;
                const PixCtx px = make_pix(p.o, e, n, oy, ox, valid);
                const int cb = n0 + half * CW;
                const bool can = half < nchunks && cb + CW <= p.o.Cout && half * CW + CW <= p.BNe;      // warp-uniform
                if (kPrefetch == 1 && !e.res_f32) {
                    pre = can;
                    if (pre && valid) {
#pragma unroll
                        for (int g = 0; g < CW / 8; ++g) {
                            pre_rh[g] = *reinterpret_cast<const uint4*>(e.res_hi + px.res_off + cb + g * 8);
                            if (e.res_lo) pre_rl[g] = *reinterpret_cast<const uint4*>(e.res_lo + px.res_off + cb + g * 8);
                        }
                    }
                } else if (can && valid) {                 // CW channels = 64 bytes = half a line per plane
                    if (e.res_f32) {
                        ptx::prefetch_l1(reinterpret_cast<const float*>(e.res_hi) + px.res_off + cb);
                    } else {
                        ptx::prefetch_l1(e.res_hi + px.res_off + cb);
                        if (e.res_lo) ptx::prefetch_l1(e.res_lo + px.res_off + cb);
                    }
                }
            }
            ptx::mbar_wait(ptx::smem_u32(&s_tfull[as]), aphase);
            ptx::tc_fence_after();
            const uint32_t t_set = tmem_base + (uint32_t)(as * p.set_stride) + ((uint32_t)(quad * 32) << 16);
            for (int mt = 0; mt < mtc; ++mt) {
                const int iy = y0 + (DXM ? 0 : 16 * mt) + ty, ix = x0 + tx;
                const uint32_t ta = t_set + (uint32_t)(mt * p.acc_stride);
                if (FLAGS == 0 && p.o.pool) {
                    // fused 2x2 max-pool: the pooled pixel (iy >> 1, ix >> 1); an odd last row / column is dropped (floor)
                    const bool pvalid = (iy >> 1) < p.o.H && (ix >> 1) < p.o.W && (!DXM || tx < 30);
                    const PixCtx px = make_pix(p.o, e, n, iy >> 1, ix >> 1, pvalid);
                    float* xbuf = s_tab + (TAB_BYTES / 4) * p.Cout_pad;      // after the constants table (merged-tap layers only)
                    for (int ch = half; ch < nchunks; ch += EPI_WARPS / 4) {
                        if (DXM) epilogue_chunk_pool<true>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(ch * CW), px, n0 + ch * CW,
                                                           (lane & 1) | ((quad & 1) << 1), xbuf, warp, lane);
                        else epilogue_chunk_pool<false>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(ch * CW), px, n0 + ch * CW,
                                                        (lane & 1) | (((lane >> 3) & 1) << 1), xbuf, warp, lane);
                    }
                    continue;
                }
                const bool valid = iy < p.in_H && ix < p.in_W && (!DXM || tx < 30);
                const int oy = p.nph == 4 ? 2 * iy + (ph >> 1) : iy;
                const int ox = p.nph == 4 ? 2 * ix + (ph & 1) : ix;
                const PixCtx px = make_pix(p.o, e, n, oy, ox, valid);
                if (STATS) {
                    // the finished values of each chunk also go into this warp's per-channel accumulators (slot j = its j-th chunk)
#define RRV_STAT_CHUNK(J)                                                                                                              \
    if (half + 2 * (J) < nchunks) {                                                                                                    \
        const int ch = half + 2 * (J);                                                                                                 \
        float xs[CW];                                                                                                                  \
        epilogue_chunk<FLAGS, DXM ? 1 : 0>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(ch * CW), px, n0 + ch * CW, ch * CW, p.BNe, false, \
                                           nullptr, nullptr, 0, xs, DXM ? o_stage : 0u);                                               \
        if (DXM && p.ostage) {     /* merged taps, Cout_pad <= 64: this warp's only chunk; its row of 30 pixels x 32 fp32 channels */ \
            ptx::fence_proxy_async();                                                                                                  \
            __syncwarp();                                                                                                              \
            if (lane == 0 && iy < p.in_H) {                                                                                            \
                ptx::tma_store_4d(&map_o_hi, o_stage, ch * CW, x0, iy, n);                                                             \
                ptx::bulk_commit();                                                                                                    \
            }                                                                                                                          \
        }                                                                                                                              \
        stat_add<(J)>(sacc, xs, valid, lane, st_mm);                                                                                   \
    }
                    RRV_STAT_CHUNK(0) RRV_STAT_CHUNK(1) RRV_STAT_CHUNK(2) RRV_STAT_CHUNK(3)
#undef RRV_STAT_CHUNK
                    st_n0 = n0;
                    continue;
                }
                if (DXM && (kStaged || (FLAGS == 0 && p.ostage))) {  // (one chunk per warp: Cout_pad <= 64)
                    if (half < nchunks)
                        epilogue_chunk<FLAGS, 1, kStaged>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(half * CW), px, n0 + half * CW, half * CW, p.BNe,
                                                          pre && mt == 0, pre_rh, pre_rl, 0, nullptr, o_stage, o_stage + (uint32_t)p.ostage);
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && half < nchunks && iy < p.in_H) {      // this warp's row: 30 pixels x its 32 channels per plane
                        ptx::tma_store_4d(&map_o_hi, o_stage, half * CW, x0, iy, n);
                        if (p.o.out_lo) ptx::tma_store_4d(&map_o_lo, o_stage + (uint32_t)p.ostage, half * CW, x0, iy, n);
                        ptx::bulk_commit();
                    }
                    continue;
                }
                if (kRR && rr) {
                    for (int ch = half; ch < nchunks; ch += EPI_WARPS / 4) {
                        // the chunk after this one: next Cout chunk, next M tile, or the first chunk of this worker's next tile
                        const RrNext nx = ch + EPI_WARPS / 4 < nchunks ? rr_next_of(tile, mt, ch + EPI_WARPS / 4)
                                          : (mt + 1 < mtc ? rr_next_of(tile, mt + 1, half) : rr_next_of(tile + n_workers, 0, half));
                        const uint32_t bh = o_stage + (uint32_t)(rr_par * 2 * p.ostage), nbh = o_stage + (uint32_t)((rr_par ^ 1) * 2 * p.ostage);
                        epilogue_chunk_rr<FLAGS>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(ch * CW), px, n0 + ch * CW, bh, bh + (uint32_t)p.ostage,
                                                 rr_rs, nx, nbh, nbh + (uint32_t)p.ostage, lane);
                        ptx::fence_proxy_async();
                        __syncwarp();
                        const int yq = y0 + 16 * mt + quad * 4;
                        if (lane == 0 && yq < p.in_H) {        // this warp's 4 rows x 8 columns x 32 channels per plane
                            ptx::tma_store_4d(&map_o_hi, bh, n0 + ch * CW, x0, yq, n);
                            if (p.o.out_lo) ptx::tma_store_4d(&map_o_lo, bh + (uint32_t)p.ostage, n0 + ch * CW, x0, yq, n);
                            ptx::bulk_commit();
                        }
                        rr_par ^= 1;
                    }
                    continue;
                }
                for (int ch = half; ch < nchunks; ch += EPI_WARPS / 4)
                    epilogue_chunk<FLAGS, DXM ? 1 : 0>(p.o, e, s_tab, p.Cout_pad, ta + (uint32_t)(ch * CW), px, n0 + ch * CW, ch * CW, p.BNe,
                                               pre && mt == 0 && ch == half, pre_rh, pre_rl);
            }
            if (STATS && p.n_ntiles > 1) {           // the next tile may cover other channels: flush (one atomic per channel and quantity)
                stat_flush<0>(sacc, p.stats, p.Cout, st_n0 + (half + 0) * CW + lane, st_mm);
                stat_flush<1>(sacc, p.stats, p.Cout, st_n0 + (half + 2) * CW + lane, st_mm);
                stat_flush<2>(sacc, p.stats, p.Cout, st_n0 + (half + 4) * CW + lane, st_mm);
                stat_flush<3>(sacc, p.stats, p.Cout, st_n0 + (half + 6) * CW + lane, st_mm);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) ptx::mbar_arrive_cluster(ptx::smem_u32(&s_tempty[as]), 0u);     // the leader's MMA warp waits for both CTAs
                else ptx::mbar_arrive(ptx::smem_u32(&s_tempty[as]));
            }
            if (++as == p.bufs) { as = 0; aphase ^= 1u; }
        }
        if (p.ostage && lane == 0) ptx::bulk_wait0();    // this lane's TMA stores have completed before the CTA's shared memory goes away
        if (STATS && p.n_ntiles == 1) {              // one Cout tile: the lane-to-channel map never changed, one flush per kernel
            const bool m2 = DXM && p.dxm == 2;       // merged phases: a warp owns chunks 0, 1 (slot = chunk), else chunks half, half + 2, ...
            stat_flush<0>(sacc, p.stats, p.Cout, (m2 ? 0 : half + 0) * CW + lane, st_mm);
            stat_flush<1>(sacc, p.stats, p.Cout, (m2 ? 1 : half + 2) * CW + lane, st_mm);
            if (!m2) {
                stat_flush<2>(sacc, p.stats, p.Cout, (half + 4) * CW + lane, st_mm);
                stat_flush<3>(sacc, p.stats, p.Cout, (half + 6) * CW + lane, st_mm);
            }
        }
    }

    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync();
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
        else ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
    if (p.tl != nullptr && threadIdx.x == 0) {
        const unsigned long long t = ptx::globaltimer();
        atomicMin(p.tl + 2, t);
        atomicMax(p.tl + 3, t);
    }
}

// ---- weight repack: OIHW fp32 -> [tap][Cout_pad][Cin] bf16 hi / lo (3x3: tap = dy*3 + dx) -------------
__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ w, int Cin, int Cout, int Cout_pad, int ksize,
                                                      int ups, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int ntaps = ups ? 16 : ksize * ksize;
    const long long total = (long long)ntaps * Cout_pad * Cin;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout_pad);
        const int t = (int)(i / ((long long)Cin * Cout_pad));
        float v = 0.0f;
        if (co < Cout) {
            const float* wk = w + ((long long)co * Cin + ci) * ksize * ksize;
            if (!ups) {
                // 3x3: blob tap t = dy * 3 + dx (PyTorch order: the three dx taps of one dy are adjacent row blocks)
                v = wk[t];
            } else {
                // phase (py,px), tap (a,b): sum of the 3x3 weights whose upsampled sample falls on
                // low-res offset (py-1+a, px-1+b):  floor((py + dy - 1) / 2) == py - 1 + a
                const int ph = t >> 2, a = (t >> 1) & 1, b = t & 1;
                const int py = ph >> 1, px = ph & 1;
                for (int dy = 0; dy < 3; ++dy) {
                    if (((py + dy + 1) >> 1) - 1 != py - 1 + a) continue;
                    for (int dx = 0; dx < 3; ++dx) {
                        if (((px + dx + 1) >> 1) - 1 != px - 1 + b) continue;
                        v += wk[dy * 3 + dx];
                    }
                }
            }
        }
        uint16_t h, l;
        split_hi_lo(v, 0, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    });
    return fn;
}

// NHWC bf16 activation [N][H][W][C]: box = 64 channels x TW x TH x 1, 128-byte swizzle.
int encode_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int TW, int TH, int row_bytes = 128) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)TW, (cuuint32_t)TH, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
        return 1;
    }
    return 0;
}

// Weights [rows = taps * Cout_pad][Cin] bf16: box = 64 channels x BN rows.
int encode_w_map(CUtensorMap* m, const void* base, int rows, int Cin, int BN, int row_bytes = 128) {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)BN};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights %dx%d) failed: %d", rows, Cin, (int)r);
        return 1;
    }
    return 0;
}

// Output planes [N][H][W][C] for the staged TMA stores of the merged-tap layers.
//   mode 1: box = 32 channels x 30 pixels of one row, rows of 64 bytes, SWIZZLE_64B (one epilogue warp's share of a tile row);
//   mode 2: the tensor seen as (C, column parity, W / 2, H, N): box = 32 channels x 1 parity x 30 low-res columns, SWIZZLE_64B
//           (one 32-channel chunk of one column phase of one output row of the nearest-x2 convolution).
// f32: the same boxes over an fp32 NHWC tensor (rows of 128 bytes, SWIZZLE_128B): the statistics-collecting instantiations.
int encode_out_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int mode, bool f32 = false) {
    CUresult r;
    const cuuint64_t eb = f32 ? 4 : 2;
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle sw = f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (mode == 1 || mode == 3) {       // mode 3 (row-reuse layers): box = 32 channels x 8 columns x 4 rows, one epilogue warp's chunk
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
        const cuuint32_t box[4] = {32, (cuuint32_t)(mode == 3 ? 8 : 30), (cuuint32_t)(mode == 3 ? 4 : 1), 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        r = encode_fn()(m, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)(W / 2), (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[4] = {(cuuint64_t)C * eb, (cuuint64_t)C * 2 * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
        const cuuint32_t box[5] = {32, 1, 30, 1, 1};
        const cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = encode_fn()(m, dt, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(output %dx%dx%dx%d, mode %d) failed: %d", N, H, W, C, mode, (int)r);
        return 1;
    }
    return 0;
}

int cout_pad_of(int Cout) { return (Cout + 15) / 16 * 16; }

OutDesc make_out(const rrv_conv* p) {
    OutDesc o;
    o.pool = p->pool ? 1 : 0;
    o.H = p->H >> o.pool; o.W = p->W >> o.pool; o.Cout = p->Cout; o.out_mode = p->out_mode; o.out_C = p->out_C;
    o.out_hi = (uint16_t*)p->out_hi; o.out_lo = (uint16_t*)p->out_lo; o.out_f32 = p->out_f32;
    o.out_img = p->out_img;
    o.crop_y0 = p->crop_y0; o.crop_x0 = p->crop_x0; o.crop_h = p->crop_h; o.crop_w = p->crop_w;
    return o;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}


int epi_flags(const rrv_epilogue& e) {
    return (e.norm1 ? EPI_N1 : 0) | (e.res_hi ? EPI_RES : 0) | (e.norm2 ? EPI_N2 : 0) | (e.affine ? EPI_AFF : 0);
}

// (the two output maps of the staged TMA stores travel in file-scope slots set by conv2d_tc2 right before the launch: every
//  launch helper below forwards them without growing its signature)
static thread_local CUtensorMap t_mo_hi, t_mo_lo;

template <int FLAGS, bool PAIR, bool DXM>
int launch_tc2p(int grid, int smem, cudaStream_t st, const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mb_hi,
                const CUtensorMap& mb_lo, const Tc2Params& d) {
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<FLAGS, PAIR, DXM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        RRV_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(conv_tc2_kernel): %s", cudaGetErrorString(e));
        attr_set = true;
    }
    static const bool no_pdl = getenv("RRV_NO_PDL") != nullptr;          // A/B switch for measurements
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (PAIR) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (!no_pdl && d.pdl_attr) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<FLAGS, PAIR, DXM>, ma_hi, ma_lo, mb_hi, mb_lo, t_mo_hi, t_mo_lo, d);
    RRV_REQUIRE(e == cudaSuccess, "cudaLaunchKernelEx(conv_tc2_kernel%s): %s", PAIR ? ", cluster 2" : "", cudaGetErrorString(e));
    return check_launch("conv_tc2_kernel");
}

template <int FLAGS>
int launch_tc2(int grid, int smem, cudaStream_t st, const CUtensorMap& ma_hi, const CUtensorMap& ma_lo, const CUtensorMap& mb_hi,
               const CUtensorMap& mb_lo, const Tc2Params& d) {
    return d.pair ? launch_tc2p<FLAGS, true, false>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d)
                  : launch_tc2p<FLAGS, false, false>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
}

int conv2d_tc2(const rrv_conv* p, cudaStream_t st) {
    const TcTune tune = tune_snapshot();
    const int ups = p->ups ? 1 : 0;
    Tc2Params d;
    memset(&d, 0, sizeof(d));
    d.o = make_out(p);
    d.N = p->N;
    d.in_H = p->H >> ups; d.in_W = p->W >> ups;
    d.Cout = p->Cout;
    d.Cout_pad = cout_pad_of(p->Cout);
    const int rowb = p->Cin == 32 ? 64 : 128;        // a 32-channel input: half-width operand rows (row-reuse loop only)
    d.row_bytes = rowb;
    d.kchunks = (p->Cin + BK - 1) / BK;
    d.nk_last = p->Cin == 32 ? 2 : 4;
    if (p->Cin_used > 0 && p->Cin_used < p->Cin) {      // trailing input channels that only meet zero weights: whole chunks, then k-slices
        d.kchunks = (p->Cin_used + BK - 1) / BK;
        d.nk_last = (p->Cin_used - (d.kchunks - 1) * BK + 15) / 16;
    }
    d.x3 = p->in_lo != nullptr;
    d.terms = !d.x3 ? 0 : (p->terms == RRV_TERMS_NO_WLO ? 2 : (p->terms == RRV_TERMS_NO_ALO ? 1 : 3));
    d.nph = ups ? 4 : 1;
    const int halo = (p->ksize == 3) ? 1 : 0;
    const int planes = d.x3 ? 2 : 1;
    const int btiles = ups ? 16 : p->ksize * p->ksize;     // weight tiles per chunk in the blob
    const int btiles_tile = ups ? 4 : btiles;               // ... of which one tile (= one phase) uses this many

    int a_stage = 0, b_slot = 0, box_w = 8, box_rows = 0;
    d.dxm = (tune.dxm && p->ksize == 3 && !ups && 3 * d.Cout_pad <= 256 && d.in_W >= 16 && rowb == 128) ? 1 : 0;
    static const bool no_ups_merge = getenv("RRV_NO_UPS_MERGE") != nullptr;      // A/B switch for measurements
    if (!no_ups_merge && tune.dxm && p->ksize == 3 && ups && 4 * d.Cout_pad <= 256 && d.Cout_pad % CW == 0 && d.in_W >= 16 &&
        (p->out_mode == RRV_OUT_PLANES || (p->out_mode == RRV_OUT_F32_NHWC && p->stats != nullptr && p->Cout == d.Cout_pad)))
        d.dxm = 2;
    const int xchg_bytes = (p->pool && d.dxm) ? EPI_WARPS * 512 : 0;       // row-partner exchange of the fused max-pool
    // merged-tap layers writing planes: per-warp output staging + TMA stores (see epilogue_chunk); RRV_NO_OSTAGE=1 is the A/B switch
    static const bool no_ostage = getenv("RRV_NO_OSTAGE") != nullptr;
    d.ostage = 0;
    if (!no_ostage && d.dxm && p->out_mode == RRV_OUT_PLANES && !p->pool && p->stats == nullptr && p->Cout == d.Cout_pad) {
        const int fl = epi_flags(p->ep);             // the instantiations that carry the staging code
        if (d.dxm == 1 && p->Cout % CW == 0 && (fl == 0 || fl == (EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF))) d.ostage = 2048;
        if (d.dxm == 2 && p->Cout == 64 && fl == EPI_N1) d.ostage = 2048;
    }
    // the statistics-collecting instantiations write fp32 NHWC: the same staging, one 4 KB buffer of 128-byte rows per warp
    const bool ostage_f32 = !no_ostage && d.dxm && p->out_mode == RRV_OUT_F32_NHWC && p->stats != nullptr && !p->pool && p->Cout == d.Cout_pad &&
                            ((d.dxm == 1 && p->Cout % CW == 0) || (d.dxm == 2 && p->Cout == 64));
    if (ostage_f32) d.ostage = 2048;
    if (d.ostage) {      // the staging rows must leave room for two A stages and two weight slots (CTA pairs halve the weight slots)
        const bool pair_ok = tune.pair && num_sms() % 2 == 0;
        const int bn = (d.dxm == 2 ? 4 : 3) * d.Cout_pad;
        const bool pr = d.dxm == 2 ? pair_ok : (pair_ok && bn >= std::min(tune.pair_min_bn, 48) && (bn / 2) % 8 == 0);
        const int need = 2 * planes * (d.dxm == 2 ? 5 : 6) * 4096 + 2 * planes * bn * 128 / (pr ? 2 : 1);
        if (need > SMEM_LIMIT - 1024 - d.Cout_pad * TAB_BYTES - xchg_bytes - (EPI_WARPS * 2 * d.ostage + 1024)) d.ostage = 0;
    }
    d.ostage_bufs = 1;
    // row-reuse layers with few MMAs per output value (the KernelFilter up-convolutions, conv2_1): staged in-place epilogue
    // (epilogue_chunk_rr); RRV_NO_RSTAGE=1 is the A/B switch
    static const bool no_rstage = getenv("RRV_NO_RSTAGE") != nullptr;
    bool rr_stage = false;
    if (!no_rstage && !no_ostage && !d.dxm && p->out_mode == RRV_OUT_PLANES && !p->pool && !ups && p->stats == nullptr &&
        p->Cout % 64 == 0 && (p->ksize * p->ksize) * (p->Cin_used > 0 ? p->Cin_used : p->Cin) <= 576) {
        const int fl = epi_flags(p->ep);
        const bool res_ok = !p->ep.res_f32 && p->ep.res_shift == 0 && p->ep.res_H == p->H && p->ep.res_W == p->W &&
                            (p->ep.res_lo != nullptr) == (p->out_lo != nullptr);
        // (staging the stores alone -- layers without a residual, e.g. conv2_1 -- measured slower than the direct stores: the
        //  staging rows cost ring slots and the stores were not what those epilogues wait for)
        if ((fl == EPI_RES || fl == (EPI_RES | EPI_N2 | EPI_AFF)) && res_ok && p->ep.res_hi != nullptr) {
            rr_stage = true;
            d.ostage = 2048;
            d.rstage = 1;
            d.ostage_bufs = 2;
        }
    }
    const int ostage_bytes = d.ostage ? EPI_WARPS * 2 * d.ostage_bufs * d.ostage + 1024 : 0;
    const int budget = SMEM_LIMIT - 1024 - d.Cout_pad * TAB_BYTES - xchg_bytes - ostage_bytes;
    if (d.dxm == 2) {
        // ---- nearest-x2 convolution with Cout <= 64 (ResidualBlock slice2.conv1): per output parity the 3x3 convolution is a 2x2
        //      convolution over the low-resolution input.  With N = Cout = 64 an MMA costs the A fetch (64 cycles) however small N is,
        //      so the two column phases px and their two column taps b go side by side along N (N = 4 Cout_pad = 256, one A read for
        //      four weight blocks) and meet in the epilogue through lane shifts, as in the merged 3x3 layers; the two row phases py
        //      are separate tiles, each with two tap rows a (A box of 5 rows x 32 columns, row a at a 4096-byte offset). ----
        d.MT = 1;
        d.nph = 2;
        d.BN = 4 * d.Cout_pad; d.BNe = d.Cout_pad; d.b_tile_rows = d.Cout_pad;
        d.n_ntiles = 1;
        d.pair = (tune.pair && num_sms() % 2 == 0) ? 1 : 0;
        d.a_plane_bytes = 5 * 4096;
        a_stage = planes * d.a_plane_bytes;
        b_slot = planes * d.BN * 128 / (d.pair ? 2 : 1);
        d.acc_stride = d.BN;
        d.set_stride = d.acc_stride;
        d.bufs = 2 * d.set_stride <= 512 ? 2 : 1;
        d.tmem_cols = 32;
        while (d.tmem_cols < d.bufs * d.set_stride) d.tmem_cols *= 2;
        // weight tiles through the ring (reloaded per tile from L2: keeping one row phase's tiles resident and walking the image
        // once per phase measured the same and doubles the DRAM reads of the input)
        d.b_resident = 0; d.a_stages = 2; d.b_slots = 2;
        int rem = budget - 2 * a_stage - 2 * b_slot;
        RRV_REQUIRE(rem >= 0, "rrv_conv2d(tcgen05 v2): merged-phase tile does not fit shared memory");
        for (;;) {
            if (d.b_slots < 4 && rem >= b_slot) { ++d.b_slots; rem -= b_slot; continue; }
            if (d.a_stages < 3 && rem >= a_stage) { ++d.a_stages; rem -= a_stage; continue; }
            break;
        }
        d.dxm_groups = 2;
        d.b_parts = d.pair ? 2 : 4;
        d.nA = 1;
        box_w = 32; box_rows = 5;
        d.tiles_x = ceil_div(d.in_W, 30 * (d.pair ? 2 : 1));
        d.tiles_y = ceil_div(d.in_H, 4);
    } else if (d.dxm) {
        d.dxm_groups = 3;
        // ---- merged dx taps: N = 3 Cout_pad, M tile = 4 rows x 32 input columns (one TMEM lane quadrant per row).  ONE A box
        //      of 6 rows x 32 columns per chunk serves all nine taps: tap row dy reads it from row dy (a 4096-byte offset, whole
        //      swizzle atoms), the three dx taps are the three column blocks of the weight tile and meet in the epilogue. ----
        d.MT = 1;
        d.BN = 3 * d.Cout_pad; d.BNe = d.Cout_pad; d.b_tile_rows = 3 * d.Cout_pad;
        d.n_ntiles = 1;
        // (each CTA of a pair holds BN / 2 weight rows: whole 8-row swizzle atoms.  The RGB head, N = 48, is a pair too -- its MMAs
        //  cost their A fetch, 64 cycles, whatever N is, and a pair covers twice the pixels per MMA; RRV_HEAD_PAIR=0 is the A/B switch)
        static const bool head_pair = !(getenv("RRV_HEAD_PAIR") && atoi(getenv("RRV_HEAD_PAIR")) == 0);
        const int min_bn = head_pair ? std::min(tune.pair_min_bn, 48) : tune.pair_min_bn;
        d.pair = (tune.pair && num_sms() % 2 == 0 && d.BN >= min_bn && (d.BN / 2) % 8 == 0) ? 1 : 0;
        d.a_plane_bytes = 6 * 4096;
        a_stage = planes * d.a_plane_bytes;
        b_slot = planes * d.BN * 128 / (d.pair ? 2 : 1);
        d.acc_stride = (d.BN + 31) / 32 * 32;
        d.set_stride = d.acc_stride;
        // the RGB head (same conditions as the EPI_HEAD dispatch below): three column blocks per accumulator set
        static const bool no_merge_wlo = getenv("RRV_NO_MERGE_WLO") != nullptr;
        const bool head_like = p->out_mode != RRV_OUT_PLANES && p->out_mode != RRV_OUT_F32_NHWC && p->Cout <= 4 && epi_flags(p->ep) == 0 &&
                               p->ep.act == 0 && !p->pool && p->stats == nullptr;
        if (!no_merge_wlo && head_like && d.x3 && d.terms == 3 && d.BN % 16 == 0 && 2 * d.BN <= 256 && (!d.pair || (d.BN / 2) % 8 == 0)) {
            d.merge_wlo = 1;
            d.set_stride = (3 * d.BN + 31) / 32 * 32;
        }
        d.bufs = 2 * d.set_stride <= 512 ? 2 : 1;
        d.tmem_cols = 32;
        while (d.tmem_cols < d.bufs * d.set_stride) d.tmem_cols *= 2;
        const int b_all = 3 * d.kchunks;
        if (b_all <= MAX_B_SLOTS && b_all * b_slot + 2 * a_stage <= budget) {
            d.b_resident = 1; d.b_slots = b_all;
            d.a_stages = std::min(MAX_STAGES, (budget - b_all * b_slot) / a_stage);
        } else {
            d.b_resident = 0; d.a_stages = 2; d.b_slots = 2;
            int rem = budget - 2 * a_stage - 2 * b_slot;
            RRV_REQUIRE(rem >= 0, "rrv_conv2d(tcgen05 v2): merged-tap tile does not fit shared memory");
            for (;;) {
                if (d.b_slots < 6 && rem >= b_slot) { ++d.b_slots; rem -= b_slot; continue; }
                if (d.a_stages < 4 && rem >= a_stage) { ++d.a_stages; rem -= a_stage; continue; }
                break;
            }
        }
        d.nA = 1; d.a_y0[0] = -1; d.a_dx[0][0] = -1; d.ngrp[0][0] = 3;
        for (int dy = 0; dy < 3; ++dy) d.grp[0][0][dy] = Grp{dy, 4 * dy, dy == 0 ? 1 : 0};   // weight tile dy = the three dx taps of row dy
        box_w = 32; box_rows = 6;
        d.tiles_x = ceil_div(d.in_W, 30 * (d.pair ? 2 : 1));
        d.tiles_y = ceil_div(d.in_H, 4);
    } else {
        // ---- tile shape: Cout tile BN, M tiles per weight tile MT ----
        int BN = std::min(d.Cout_pad, tune.max_bn);
        int MT = std::max(1, std::min(tune.mt, 2));
        // few MMAs per output value (the 1x1 shortcuts, the KernelFilter up-convolutions): the epilogue is the kernel's time and
        // nothing is gained from sharing a weight tile between two M tiles; smaller work items balance the 148 SMs better
        const int k_real = (ups ? 4 : p->ksize * p->ksize) * (p->Cin_used > 0 ? p->Cin_used : p->Cin);
        if (k_real < 576 && tune.mt == 2 && tune.max_bn == 256) {
            MT = 1;
            // (a 32-channel input issues 2 k-slices per tap: the MMAs, 64 + N / 4 cycles each whatever K is, set the pace, and wide
            //  ones cost less per output: the KernelFilter up-convolution 0.060 -> 0.048 ms at N = 256)
            if (rowb == 128) BN = std::min(BN, 128);
        }
        while (BN > 16 && d.Cout_pad % BN != 0) BN -= 16;
        if (d.in_H <= 16) MT = 1;
        // all weight tiles resident beats sharing them between two M tiles: prefer MT = 1 if that is what fits
        const bool resident_shape = !ups && BN == d.Cout_pad && btiles * d.kchunks <= MAX_B_SLOTS;
        if (MT == 2 && resident_shape) {
            const int b_all = btiles * d.kchunks * planes * BN * rowb;
            const int a2 = planes * (32 + 2 * halo) * 8 * rowb, a1 = planes * (16 + 2 * halo) * 8 * rowb;
            if (b_all + 2 * a2 > budget && b_all + 2 * a1 <= budget) MT = 1;
        }
        for (;;) {
            const int acc_stride = (BN + 31) / 32 * 32;
            a_stage = planes * (16 * MT + 2 * halo) * 8 * rowb;
            b_slot = planes * BN * rowb;
            const bool tmem_ok = MT * acc_stride <= 512;
            const bool smem_ok = 2 * a_stage + 2 * b_slot <= budget;
            if (tmem_ok && smem_ok) break;
            if (MT > 1) { MT = 1; continue; }
            RRV_REQUIRE(BN > 16, "rrv_conv2d(tcgen05 v2): no tile shape fits (Cout=%d)", p->Cout);
            BN -= 16;
            while (BN > 16 && d.Cout_pad % BN != 0) BN -= 16;
        }
        if (rr_stage && BN % 64 != 0) {      // (a tuned-down Cout tile: each of a quadrant's two warps needs whole 32-channel chunks)
            rr_stage = false;
            d.ostage = 0; d.ostage_bufs = 1; d.rstage = 0;
        }
        d.BN = BN; d.MT = MT; d.BNe = BN; d.b_tile_rows = d.Cout_pad;
        // CTA pairs everywhere except where all weight tiles can stay resident (the 64 -> 64 layers)
        const bool resident_fits = resident_shape && btiles * d.kchunks * b_slot + 2 * a_stage <= budget;
        d.pair = (tune.pair && (!resident_fits || tune.pair_min_bn <= 64) && BN >= tune.pair_min_bn && BN % 32 == 0 && num_sms() % 2 == 0) ? 1 : 0;
        if (d.pair) b_slot /= 2;
        d.n_ntiles = d.Cout_pad / BN;
        d.acc_stride = (BN + 31) / 32 * 32;
        d.set_stride = MT * d.acc_stride;
        d.bufs = 2 * d.set_stride <= 512 ? 2 : 1;
        d.tmem_cols = 32;
        while (d.tmem_cols < d.bufs * d.set_stride) d.tmem_cols *= 2;
        d.a_plane_bytes = (16 * MT + 2 * halo) * 8 * rowb;

        // ---- rings ----
        const int b_all = btiles * d.kchunks;
        if (!d.pair && resident_fits && d.n_ntiles == 1) {
            d.b_resident = 1;
            d.b_slots = b_all;
            d.a_stages = std::min(4, (budget - b_all * b_slot) / a_stage);
        } else {
            d.b_resident = 0;
            d.a_stages = 2; d.b_slots = 2;
            int rem = budget - 2 * a_stage - 2 * b_slot;
            for (;;) {
                if (d.b_slots < 4 && rem >= b_slot) { ++d.b_slots; rem -= b_slot; continue; }
                if (d.a_stages < 3 && rem >= a_stage) { ++d.a_stages; rem -= a_stage; continue; }
                if (d.b_slots < 8 && rem >= b_slot) { ++d.b_slots; rem -= b_slot; continue; }
                if (rowb == 64 && d.a_stages < 6 && rem >= a_stage) { ++d.a_stages; rem -= a_stage; continue; }      // (three boxes per tile)
                break;
            }
        }
        (void)btiles_tile;

        // ---- which weight tiles meet which A box ----
        if (p->ksize == 1) {
            d.nA = 1; d.a_dx[0][0] = 0; d.a_y0[0] = 0; d.ngrp[0][0] = 1;
            d.grp[0][0][0] = Grp{0, 0, 1};
        } else if (!ups) {
            // box = rows y0-1 .. y0+16MT, columns x0+dx ..; tap (dy, dx) reads it from row dy
            d.nA = 3; d.a_y0[0] = -1;
            for (int j = 0; j < 3; ++j) {
                d.a_dx[0][j] = j - 1; d.ngrp[0][j] = 3;
                for (int dy = 0; dy < 3; ++dy) d.grp[0][j][dy] = Grp{dy * 3 + j, dy, (j == 0 && dy == 0) ? 1 : 0};
            }
        } else {
            // phase (py, px) of the nearest-x2 convolution = a 2x2 convolution over the low-res input with taps at rows
            // py-1+a and columns px-1+b: box origin row y0+py-1, boxes at columns x0+px-1 and x0+px, tap a reads from row a
            d.nA = 2;
            for (int ph = 0; ph < 4; ++ph) {
                const int py = ph >> 1, px = ph & 1;
                d.a_y0[ph] = py - 1;
                for (int b = 0; b < 2; ++b) {
                    d.a_dx[ph][b] = px - 1 + b; d.ngrp[ph][b] = 2;
                    for (int a = 0; a < 2; ++a) d.grp[ph][b][a] = Grp{ph * 4 + a * 2 + b, a, (a == 0 && b == 0) ? 1 : 0};
                }
            }
        }

        box_w = 8; box_rows = 16 * MT + 2 * halo;
        d.tiles_x = ceil_div(d.in_W, d.pair ? 16 : 8);          // a pair takes two horizontally adjacent tiles
        d.tiles_y = ceil_div(d.in_H, 16 * MT);
    }
    const long long total = (long long)d.N * d.tiles_y * d.tiles_x * d.n_ntiles * d.nph;
    RRV_REQUIRE(total < (1LL << 31), "rrv_conv2d: too many tiles");
    d.total_tiles = (int)total;
    d.ep = make_epi(p->ep, p->Cout);
    d.ep.lo_fp16 = 0;
    d.stats = p->stats;
    d.stats_minmax = p->stats_minmax;
    d.pdl_attr = tune.pdl;
    d.tl = nullptr;
    if (g_tl_base != nullptr && g_tl_next < g_tl_slots) d.tl = g_tl_base + 4 * (g_tl_next++);

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    const int rows = btiles * d.Cout_pad;
    const uint16_t* w_hi = (const uint16_t*)p->w_tc;
    const uint16_t* w_lo = w_hi + (long long)rows * p->Cin;
    if (encode_act_map(&ma_hi, p->in_hi, d.N, d.in_H, d.in_W, p->Cin, box_w, box_rows, rowb)) return 1;
    const int b_box = d.dxm == 2 ? d.Cout_pad : (d.pair ? d.BN / 2 : d.BN);
    if (encode_w_map(&mb_hi, w_hi, rows, p->Cin, b_box, rowb)) return 1;
    if (d.x3) {
        if (encode_act_map(&ma_lo, p->in_lo, d.N, d.in_H, d.in_W, p->Cin, box_w, box_rows, rowb)) return 1;
        if (encode_w_map(&mb_lo, w_lo, rows, p->Cin, b_box, rowb)) return 1;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    d.ostage_off = (d.a_stages * a_stage + d.b_slots * b_slot + d.Cout_pad * TAB_BYTES + xchg_bytes + 1023) / 1024 * 1024;
    memset(&t_mo_hi, 0, sizeof(t_mo_hi));
    memset(&t_mo_lo, 0, sizeof(t_mo_lo));
    if (d.ostage && ostage_f32) {
        if (encode_out_map(&t_mo_hi, p->out_f32, d.N, d.o.H, d.o.W, p->Cout, d.dxm, true)) return 1;
        t_mo_lo = t_mo_hi;
    } else if (d.ostage) {
        const int omode = d.dxm ? d.dxm : 3;
        if (encode_out_map(&t_mo_hi, p->out_hi, d.N, d.o.H, d.o.W, p->Cout, omode)) return 1;
        if (p->out_lo) {
            if (encode_out_map(&t_mo_lo, p->out_lo, d.N, d.o.H, d.o.W, p->Cout, omode)) return 1;
        } else {
            t_mo_lo = t_mo_hi;
        }
    }
    const int smem = d.a_stages * a_stage + d.b_slots * b_slot + d.Cout_pad * TAB_BYTES + xchg_bytes + (d.ostage ? ostage_bytes : 0) + 1024;
    // the specialised instantiations write planes / NHWC only; NCHW and the finished BGR frame (the RGB head) take the generic one
    const bool plain_out = p->out_mode == RRV_OUT_PLANES || p->out_mode == RRV_OUT_F32_NHWC;
    const int flags = plain_out ? epi_flags(p->ep) : -1;
    const int grid = d.pair ? 2 * std::min(d.total_tiles, num_sms() / 2) : std::min(d.total_tiles, num_sms());
    if (p->stats != nullptr) {          // statistics of the written values from the epilogue (bias + activation only)
        if (d.dxm) {
            if (d.pair) return launch_tc2p<EPI_STATS, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
            return launch_tc2p<EPI_STATS, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        }
        return launch_tc2<EPI_STATS>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
    }
    if (d.dxm == 1 && !plain_out && p->Cout <= 4 && epi_flags(p->ep) == 0 && p->ep.act == 0 && !p->pool) {     // the RGB head
        if (d.pair) return launch_tc2p<EPI_HEAD, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        return launch_tc2p<EPI_HEAD, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
    }
    if (d.dxm) {
        // merged-tap layers: conv1_2 (bias + ReLU [+ pool]), slice2.conv2 (the full chain), the RGB head / anything else (generic).
        // The full-chain and norm1-only instantiations store through the staging rows only: without staging, the generic one.
        int f = (flags == 0 || flags == (EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF)) ? flags : -1;      // (flags < 0 stays generic)
        if (f > 0 && (!d.ostage || d.dxm == 2)) f = -1;
        if (flags == EPI_N1 && d.ostage) {          // slice2.conv1 (merged column phases)
            if (d.pair) return launch_tc2p<EPI_N1, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
            return launch_tc2p<EPI_N1, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        }
        if (d.pair) {
            if (f == 0) return launch_tc2p<0, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
            if (f > 0) return launch_tc2p<EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
            return launch_tc2p<-1, true, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        }
        if (f == 0) return launch_tc2p<0, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        if (f > 0) return launch_tc2p<EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        return launch_tc2p<-1, false, true>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
    }
    switch (flags) {
        case 0: return launch_tc2<0>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        case EPI_N1: return launch_tc2<EPI_N1>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        case EPI_RES: return launch_tc2<EPI_RES>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        case EPI_RES | EPI_N2 | EPI_AFF: return launch_tc2<EPI_RES | EPI_N2 | EPI_AFF>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        case EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF:
            return launch_tc2<EPI_N1 | EPI_RES | EPI_N2 | EPI_AFF>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
        default: return launch_tc2<-1>(grid, smem, st, ma_hi, ma_lo, mb_hi, mb_lo, d);
    }
}

}  // namespace

int tc_tune(int max_bn, int mt) {
    RRV_REQUIRE(max_bn >= 16 && max_bn <= 256 && max_bn % 16 == 0, "rrv_tc_tune: max_bn must be a multiple of 16 in [16, 256]");
    RRV_REQUIRE(mt == 1 || mt == 2, "rrv_tc_tune: mt must be 1 or 2");
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tune.max_bn = max_bn;
    g_tune.mt = mt;
    return 0;
}

int tc_tune_pair(int enable, int min_bn) {
    RRV_REQUIRE(min_bn >= 32 && min_bn <= 256 && min_bn % 32 == 0, "rrv_tc_tune_pair: min_bn must be a multiple of 32 in [32, 256]");
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tune.pair = enable ? 1 : 0;
    g_tune.pair_min_bn = min_bn;
    return 0;
}

int tc_tune_merge(int enable) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tune.dxm = enable ? 1 : 0;
    return 0;
}

int tc_timeline(unsigned long long* buf, int nslots) {
    g_tl_base = nslots > 0 ? buf : nullptr;
    g_tl_slots = nslots > 0 ? nslots : 0;
    g_tl_next = 0;
    return 0;
}

int tc_tune_pdl(int enable) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    g_tune.pdl = enable ? 1 : 0;
    return 0;
}

long long tc_weight_bytes(int Cin, int Cout, int ksize, int ups) {
    if (Cin <= 0 || Cout <= 0 || (Cin % BK != 0 && !(Cin == 32 && !ups))) return 0;          // the FFMA kernel takes the other shapes
    if (!(ksize == 3 || (ksize == 1 && !ups))) return 0;
    const long long ntaps = ups ? 16 : ksize * ksize;
    return 2LL * ntaps * cout_pad_of(Cout) * Cin * 2;
}

int pack_weights_tc(const float* w, int Cin, int Cout, int ksize, int ups, void* blob, cudaStream_t st) {
    RRV_REQUIRE(w && blob, "rrv_pack_weights_tc: NULL tensor");
    const long long bytes = tc_weight_bytes(Cin, Cout, ksize, ups);
    RRV_REQUIRE(bytes > 0, "rrv_pack_weights_tc: unsupported shape Cin=%d Cout=%d k=%d ups=%d", Cin, Cout, ksize, ups);
    const int cp = cout_pad_of(Cout);
    const long long total = bytes / 4;
    uint16_t* hi = (uint16_t*)blob;
    uint16_t* lo = hi + total;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
    pack_tc_kernel<<<grid, 256, 0, st>>>(w, Cin, Cout, cp, ksize, ups ? 1 : 0, hi, lo);
    return check_launch("pack_tc_kernel");
}

int conv2d_tc(const rrv_conv* p, cudaStream_t st) {
    RRV_REQUIRE(encode_fn() != nullptr, "rrv_conv2d(tcgen05): cuTensorMapEncodeTiled is not available from the driver");
    RRV_REQUIRE(g_lo_fp16 == 0, "rrv_conv2d(tcgen05): the tensor-core path needs bf16 lo planes (rrv_set_lo_format(0))");
    const int ups = p->ups ? 1 : 0;
    RRV_REQUIRE(tc_weight_bytes(p->Cin, p->Cout, p->ksize, ups) > 0,
                "rrv_conv2d(tcgen05): unsupported shape Cin=%d Cout=%d k=%d ups=%d (Cin must be 32 or a multiple of 64)", p->Cin,
                p->Cout, p->ksize, ups);
    RRV_REQUIRE(p->w_tc != nullptr, "rrv_conv2d(tcgen05): w_tc is NULL");
    RRV_REQUIRE(p->in_hi != nullptr, "rrv_conv2d: in_hi is NULL");
    RRV_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "rrv_conv2d: empty output %dx%dx%d", p->N, p->H, p->W);
    RRV_REQUIRE(!ups || (p->H % 2 == 0 && p->W % 2 == 0), "rrv_conv2d: ups needs even output size");
    if (p->out_mode == RRV_OUT_PLANES) {
        RRV_REQUIRE(p->out_hi != nullptr, "rrv_conv2d: out_hi is NULL");
        RRV_REQUIRE(p->Cout % 8 == 0, "rrv_conv2d: planes output needs Cout %% 8 == 0");
        RRV_REQUIRE((p->in_lo == nullptr) == (p->out_lo == nullptr), "rrv_conv2d: planes in/out must both be x3 or both bf16");
    } else if (p->out_mode == RRV_OUT_BGR_F32 || p->out_mode == RRV_OUT_BGR_U8) {
        RRV_REQUIRE(p->out_img != nullptr, "rrv_conv2d: out_img is NULL");
        RRV_REQUIRE(p->Cout == 3 && !p->pool, "rrv_conv2d: the BGR frame output belongs to the 3-channel RGB head");
        RRV_REQUIRE(p->crop_h > 0 && p->crop_w > 0 && p->crop_y0 >= 0 && p->crop_x0 >= 0 && p->crop_y0 + p->crop_h <= p->H &&
                        p->crop_x0 + p->crop_w <= p->W,
                    "rrv_conv2d: crop window (%d,%d,%d,%d) outside the %dx%d result", p->crop_y0, p->crop_x0, p->crop_h, p->crop_w, p->H, p->W);
    } else {
        RRV_REQUIRE(p->out_mode == RRV_OUT_F32_NHWC || p->out_mode == RRV_OUT_F32_NCHW, "rrv_conv2d: unknown out_mode %d", p->out_mode);
        RRV_REQUIRE(p->out_f32 != nullptr, "rrv_conv2d: out_f32 is NULL");
        RRV_REQUIRE(p->out_mode != RRV_OUT_F32_NCHW || p->out_C > 0, "rrv_conv2d: out_C must be set for NCHW output");
    }
    RRV_REQUIRE(p->terms >= RRV_TERMS_FULL && p->terms <= RRV_TERMS_NO_ALO, "rrv_conv2d: bad terms %d", p->terms);
    if (p->stats != nullptr) {
        RRV_REQUIRE(!p->pool && epi_flags(p->ep) == 0, "rrv_conv2d(stats): only bias + activation may precede the fused statistics");
        RRV_REQUIRE(p->out_mode == RRV_OUT_PLANES || p->out_mode == RRV_OUT_F32_NHWC, "rrv_conv2d(stats): planes or NHWC output");
    }

    if (p->pool) {
        RRV_REQUIRE(!ups && p->out_mode == RRV_OUT_PLANES && p->Cout % 32 == 0,
                    "rrv_conv2d(pool): needs a plain (not upsampling) convolution, planes output and Cout %% 32 == 0");
        RRV_REQUIRE(epi_flags(p->ep) == 0, "rrv_conv2d(pool): only bias + activation may precede the fused max-pool");
        RRV_REQUIRE(p->H >= 2 && p->W >= 2, "rrv_conv2d(pool): empty pooled output");
    }
    return conv2d_tc2(p, st);
}

}  // namespace rrv
