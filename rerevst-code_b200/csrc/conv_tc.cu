// Tensor-core convolution: im2col-free implicit GEMM on tcgen05 (sm_100a).
//
// Replaces every nn.Conv2d / F.conv2d of the per-frame path except conv1_1
// (test/style_network_global.py:103-105, 181-187, 205, 275-281, 341) together with the nearest x2
// upsample of ResidualBlock.forward (:113) and the pointwise chains that follow each convolution
// (:52-55, :116-122, :364), which run in the epilogue (rrv_common.cuh: apply_epilogue).
//
// GEMM view: D[M = 128 output pixels][N = Cout tile] += A[M][K] * B[N][K]^T with K = taps x Cin.
//   A  one TMA box per (tap, 64-channel chunk): a TH x TW window of the NHWC input shifted by the
//      tap offset; rows outside the image are zero-filled by TMA, which IS the conv's zero padding.
//      The box lands in shared memory as 128 rows of 128 bytes with the 128-byte swizzle, i.e. the
//      canonical K-major operand layout of tcgen05.mma -- no im2col buffer anywhere.
//   B  weights repacked once at load time to [tap][Cout][Cin] bf16, one TMA box per k-step.
//   D  fp32 accumulators in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of
//      tile i+1.  Persistent CTAs (one per SM) walk the tile list.
// fp32 accuracy ("x3"): every fp32 operand v is carried as hi = bf16(v), lo = bf16(v - hi) and each
// k-slice issues Ahi*Bhi + Ahi*Blo + Alo*Bhi into the same accumulator (the dropped lo*lo term is
// 2^-18 relative).  With lo == NULL a single bf16 MMA is issued (BASELINE config 3).
// Nearest x2 upsample: the 3x3 conv over the upsampled image collapses, per output parity (py,px),
// to a 2x2 conv over the low-resolution input with summed weights (9 -> 4 taps); the four phases
// are four GEMMs over the same low-res tile that scatter to interleaved output pixels.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) and TMEM owner,
// warps 2..5 = epilogue (tcgen05.ld -> fused pointwise chain -> 16-byte stores).
#include <cuda.h>

#include <mutex>

#include "rrv_common.cuh"
#include "tc_ptx.cuh"

namespace rrv {

namespace {

constexpr int BM = 128;          // output pixels per tile (= TMEM lanes)
constexpr int BK = 64;           // channels per k-step (128-byte rows)
constexpr int A_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 8;
constexpr int TC_THREADS = 192;
constexpr int SMEM_LIMIT = 227 * 1024 - 1024;   // dynamic part: the opt-in maximum minus the static barriers

struct TcTune {
    int max_bn = 256;
    int tile_w = 16;
    int max_stages = 6;
};
TcTune g_tune;

struct TcParams {
    int N, H, W;            // output
    int in_H, in_W;         // input (H/2, W/2 when ups)
    int Cout, Cout_pad;
    int kchunks;            // Cin / 64
    int ntaps;              // taps per phase: 9, 1, or 4 (ups)
    int nphase;             // 1, or 4 (ups)
    int ksize;
    int tiles_x, tiles_y, n_ntiles, total_tiles;
    int tw_shift;           // TW = 1 << tw_shift, TH = 128 >> tw_shift
    int BN;
    int stages;
    int x3;
    int acc_stride, tmem_cols;
    int out_mode, out_C;
    uint16_t* out_hi;
    uint16_t* out_lo;
    float* out_f32;
    EpiDev ep;
};

struct TileCoord {
    int n, y0, x0, n0, py, px, phase;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
    TileCoord c;
    c.phase = t % p.nphase; t /= p.nphase;
    const int nt = t % p.n_ntiles; t /= p.n_ntiles;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    c.n = t / p.tiles_y;
    c.y0 = ty * (BM >> p.tw_shift);
    c.x0 = tx << p.tw_shift;
    c.n0 = nt * p.BN;
    c.py = c.phase >> 1;
    c.px = c.phase & 1;
    return c;
}

__device__ __forceinline__ void tap_offset(const TcParams& p, const TileCoord& c, int t, int& oy, int& ox) {
    if (p.nphase == 4) {          // 2x2 taps of one upsample phase
        oy = c.py - 1 + (t >> 1);
        ox = c.px - 1 + (t & 1);
    } else if (p.ksize == 3) {
        oy = t / 3 - 1;
        ox = t % 3 - 1;
    } else {
        oy = 0;
        ox = 0;
    }
}

__device__ __forceinline__ void store_group(const TcParams& p, const float* v, int n, int oy, int ox, int c0, int nvalid) {
    const long long pix = ((long long)n * p.H + oy) * p.W + ox;
    if (p.out_mode == RRV_OUT_PLANES) {
        store8(p.out_hi + pix * p.Cout + c0, p.out_lo ? p.out_lo + pix * p.Cout + c0 : nullptr, 0, v);
    } else if (p.out_mode == RRV_OUT_F32_NHWC) {
        float* o = p.out_f32 + pix * p.Cout + c0;
        if (nvalid == 8) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nvalid) o[k] = v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nvalid && c0 + k < p.out_C)
                p.out_f32[(((long long)n * p.out_C + c0 + k) * p.H + oy) * p.W + ox] = v[k];
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[MAX_STAGES], s_empty[MAX_STAGES], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t stage_bytes = (p.x3 ? 2u : 1u) * ((uint32_t)A_BYTES + b_bytes);
    const uint32_t off_a_lo = A_BYTES;
    const uint32_t off_b_hi = p.x3 ? 2u * A_BYTES : (uint32_t)A_BYTES;
    const uint32_t off_b_lo = off_b_hi + b_bytes;
    const int ksteps = p.ntaps * p.kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(ptx::smem_u32(&s_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(ptx::smem_u32(&s_tfull[a]), 1);
            ptx::mbar_init(ptx::smem_u32(&s_tempty[a]), 4);     // one arrival per epilogue warp
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
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&s_tmem_base), (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord c = decode_tile(p, tile);
                for (int t = 0; t < p.ntaps; ++t) {
                    int oy, ox;
                    tap_offset(p, c, t, oy, ox);
                    const int brow = (c.phase * p.ntaps + t) * p.Cout_pad + c.n0;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1u);
                        const uint32_t full = ptx::smem_u32(&s_full[stage]);
                        const uint32_t sb = smem_base + (uint32_t)stage * stage_bytes;
                        ptx::mbar_expect_tx(full, stage_bytes);
                        ptx::tma_load_4d(sb, &map_a_hi, full, kc * BK, c.x0 + ox, c.y0 + oy, c.n);
                        ptx::tma_load_2d(sb + off_b_hi, &map_b_hi, full, kc * BK, brow);
                        if (p.x3) {
                            ptx::tma_load_4d(sb + off_a_lo, &map_a_lo, full, kc * BK, c.x0 + ox, c.y0 + oy, c.n);
                            ptx::tma_load_2d(sb + off_b_lo, &map_b_lo, full, kc * BK, brow);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = ptx::make_idesc_bf16(BM, p.BN);
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            ptx::mbar_wait(ptx::smem_u32(&s_tempty[as]), aphase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride);
            for (int ks = 0; ks < ksteps; ++ks) {
                ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t sb = smem_base + (uint32_t)stage * stage_bytes;
                    const uint64_t a_hi = ptx::make_smem_desc_sw128(sb);
                    const uint64_t b_hi = ptx::make_smem_desc_sw128(sb + off_b_hi);
                    const uint64_t a_lo = ptx::make_smem_desc_sw128(sb + off_a_lo);
                    const uint64_t b_lo = ptx::make_smem_desc_sw128(sb + off_b_lo);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);       // 16 bf16 = 32 bytes = 2 x 16-byte units
                        ptx::mma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, (ks | k) != 0);
                        if (p.x3) {
                            ptx::mma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                            ptx::mma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                    }
                    ptx::mma_commit(ptx::smem_u32(&s_empty[stage]));          // frees the smem slot when the MMAs retire
                    if (ks == ksteps - 1) ptx::mma_commit(ptx::smem_u32(&s_tfull[as]));
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else {
        // ================= epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int ty = m >> p.tw_shift, tx = m & ((1 << p.tw_shift) - 1);
        const int nchunks = (p.BN + 31) / 32;
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const TileCoord c = decode_tile(p, tile);
            const int iy = c.y0 + ty, ix = c.x0 + tx;
            const bool valid = iy < p.in_H && ix < p.in_W;
            const int oy = p.nphase == 4 ? 2 * iy + c.py : iy;
            const int ox = p.nphase == 4 ? 2 * ix + c.px : ix;
            ptx::mbar_wait(ptx::smem_u32(&s_tfull[as]), aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(as * p.acc_stride) + ((uint32_t)(quad * 32) << 16);
            for (int ch = 0; ch < nchunks; ++ch) {
                float v[32];
                ptx::tmem_ld32(taddr + (uint32_t)(ch * 32), v);
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c0 = c.n0 + ch * 32 + g * 8;
                        if (c0 >= p.Cout || ch * 32 + g * 8 >= p.BN) continue;
                        if (c0 + 8 <= p.Cout) {
                            apply_epilogue<8>(p.ep, v + g * 8, c.n, oy, ox, c0);
                            store_group(p, v + g * 8, c.n, oy, ox, c0, 8);
                        } else {
                            const int nv = p.Cout - c0;
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k < nv) apply_epilogue<1>(p.ep, v + g * 8 + k, c.n, oy, ox, c0 + k);
                            store_group(p, v + g * 8, c.n, oy, ox, c0, nv);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&s_tempty[as]));
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---- weight repack: OIHW fp32 -> [tap][Cout_pad][Cin] bf16 hi / lo ---------------------------------
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
int encode_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int TW, int TH) {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
        return 1;
    }
    return 0;
}

// Weights [rows = taps * Cout_pad][Cin] bf16: box = 64 channels x BN rows.
int encode_w_map(CUtensorMap* m, const void* base, int rows, int Cin, int BN) {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights %dx%d) failed: %d", rows, Cin, (int)r);
        return 1;
    }
    return 0;
}

int cout_pad_of(int Cout) { return (Cout + 15) / 16 * 16; }

int pick_bn(int Cout_pad) {
    int bn = std::min(Cout_pad, g_tune.max_bn);
    while (bn > 16 && Cout_pad % bn != 0) bn -= 16;
    return bn;
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

}  // namespace

int tc_tune(int max_bn, int tile_w, int max_stages) {
    RRV_REQUIRE(max_bn >= 16 && max_bn <= 256 && max_bn % 16 == 0, "rrv_tc_tune: max_bn must be a multiple of 16 in [16, 256]");
    RRV_REQUIRE(tile_w == 8 || tile_w == 16 || tile_w == 32 || tile_w == 64 || tile_w == 128, "rrv_tc_tune: tile_w must be 8..128, power of 2");
    RRV_REQUIRE(max_stages >= 2 && max_stages <= MAX_STAGES, "rrv_tc_tune: max_stages must be in [2, %d]", MAX_STAGES);
    g_tune.max_bn = max_bn;
    g_tune.tile_w = tile_w;
    g_tune.max_stages = max_stages;
    return 0;
}

long long tc_weight_bytes(int Cin, int Cout, int ksize, int ups) {
    if (Cin <= 0 || Cout <= 0 || Cin % BK != 0) return 0;          // the FFMA kernel takes the other shapes
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
                "rrv_conv2d(tcgen05): unsupported shape Cin=%d Cout=%d k=%d ups=%d (Cin must be a multiple of 64)", p->Cin,
                p->Cout, p->ksize, ups);
    RRV_REQUIRE(p->w_tc != nullptr, "rrv_conv2d(tcgen05): w_tc is NULL");
    RRV_REQUIRE(p->in_hi != nullptr, "rrv_conv2d: in_hi is NULL");
    RRV_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "rrv_conv2d: empty output %dx%dx%d", p->N, p->H, p->W);
    RRV_REQUIRE(!ups || (p->H % 2 == 0 && p->W % 2 == 0), "rrv_conv2d: ups needs even output size");
    if (p->out_mode == RRV_OUT_PLANES) {
        RRV_REQUIRE(p->out_hi != nullptr, "rrv_conv2d: out_hi is NULL");
        RRV_REQUIRE(p->Cout % 8 == 0, "rrv_conv2d: planes output needs Cout %% 8 == 0");
        RRV_REQUIRE((p->in_lo == nullptr) == (p->out_lo == nullptr), "rrv_conv2d: planes in/out must both be x3 or both bf16");
    } else {
        RRV_REQUIRE(p->out_f32 != nullptr, "rrv_conv2d: out_f32 is NULL");
        RRV_REQUIRE(p->out_mode != RRV_OUT_F32_NCHW || p->out_C > 0, "rrv_conv2d: out_C must be set for NCHW output");
    }

    TcParams d;
    d.N = p->N; d.H = p->H; d.W = p->W;
    d.in_H = p->H >> ups; d.in_W = p->W >> ups;
    d.Cout = p->Cout;
    d.Cout_pad = cout_pad_of(p->Cout);
    d.kchunks = p->Cin / BK;
    d.ksize = p->ksize;
    d.nphase = ups ? 4 : 1;
    d.ntaps = ups ? 4 : p->ksize * p->ksize;
    d.BN = pick_bn(d.Cout_pad);
    d.n_ntiles = d.Cout_pad / d.BN;
    int tw = g_tune.tile_w;
    while (tw > 8 && tw / 2 >= d.in_W) tw /= 2;                  // narrow images: fewer wasted columns
    d.tw_shift = 0;
    while ((1 << d.tw_shift) < tw) ++d.tw_shift;
    const int th = BM / tw;
    d.tiles_x = ceil_div(d.in_W, tw);
    d.tiles_y = ceil_div(d.in_H, th);
    const long long total = (long long)d.N * d.tiles_y * d.tiles_x * d.n_ntiles * d.nphase;
    RRV_REQUIRE(total < (1LL << 31), "rrv_conv2d: too many tiles");
    d.total_tiles = (int)total;
    d.x3 = p->in_lo != nullptr;
    const int stage_bytes = (d.x3 ? 2 : 1) * (A_BYTES + d.BN * 128);
    d.stages = std::min(g_tune.max_stages, (SMEM_LIMIT - 2048) / stage_bytes);
    RRV_REQUIRE(d.stages >= 2, "rrv_conv2d(tcgen05): tile does not fit shared memory (BN=%d)", d.BN);
    d.acc_stride = (d.BN + 31) / 32 * 32;
    d.tmem_cols = 32;
    while (d.tmem_cols < 2 * d.acc_stride) d.tmem_cols *= 2;
    d.out_mode = p->out_mode;
    d.out_C = p->out_C;
    d.out_hi = (uint16_t*)p->out_hi;
    d.out_lo = (uint16_t*)p->out_lo;
    d.out_f32 = p->out_f32;
    d.ep = make_epi(p->ep, p->Cout);
    d.ep.lo_fp16 = 0;

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    const int rows = d.nphase * d.ntaps * d.Cout_pad;
    const uint16_t* w_hi = (const uint16_t*)p->w_tc;
    const uint16_t* w_lo = w_hi + (long long)rows * p->Cin;
    if (encode_act_map(&ma_hi, p->in_hi, d.N, d.in_H, d.in_W, p->Cin, tw, th)) return 1;
    if (encode_w_map(&mb_hi, w_hi, rows, p->Cin, d.BN)) return 1;
    if (d.x3) {
        if (encode_act_map(&ma_lo, p->in_lo, d.N, d.in_H, d.in_W, p->Cin, tw, th)) return 1;
        if (encode_w_map(&mb_lo, w_lo, rows, p->Cin, d.BN)) return 1;
    } else {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }

    const int smem = d.stages * stage_bytes + 1024;
    static int smem_set = 0;
    if (smem > smem_set) {
        const cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        RRV_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(conv_tc_kernel): %s", cudaGetErrorString(e));
        smem_set = SMEM_LIMIT;
    }
    const int grid = std::min(d.total_tiles, num_sms());
    conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, d);
    return check_launch("conv_tc_kernel");
}

}  // namespace rrv
