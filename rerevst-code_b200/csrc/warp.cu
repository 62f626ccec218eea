// Optical-flow warp of the temporal loss: train/loss_networks.py:20-38 (warp) and :106-111
// (TemporalLoss.forward).  HBM-bound gather: 8 B flow + 4C B read + 4C B written per pixel
// (+ 4C B of the second frame for the loss).
//
// The integer source index must equal the reference's bit for bit, so the coordinate math
// repeats the reference's fp32 operations one by one with explicitly rounded intrinsics (no FMA
// contraction): grid - flo; 2*v; /max(size-1,1); -1  (loss_networks.py:30-35), then ATen's
// grid_sampler_unnormalize (align_corners=False): ((g+1)*size - 1)/2, clip_coordinates to
// [0, size-1], nearbyint (ties to even).
#include "rrv_common.cuh"

namespace rrv {

__device__ __forceinline__ int src_index(float pos, float flow, int size) {
    const float v = __fsub_rn(pos, flow);
    const float t = __fmul_rn(2.0f, v);
    const float q = __fdiv_rn(t, (float)(size - 1 > 1 ? size - 1 : 1));
    const float g = __fsub_rn(q, 1.0f);
    const float a = __fadd_rn(g, 1.0f);
    const float m = __fmul_rn(a, (float)size);
    const float s = __fsub_rn(m, 1.0f);
    const float u = __fdiv_rn(s, 2.0f);
    const float c = fminf((float)(size - 1), fmaxf(u, 0.0f));
    return (int)rintf(c);
}

// MODE 0: warp only.  MODE 1: warp + sum |warped - second| (TemporalLoss.forward).
// Grid (ceil(W / (VEC * 128)), H, B): blockIdx.y is the row and blockIdx.z the batch item, so there is no 64-bit div/mod per
// pixel; a thread owns VEC = 4 consecutive pixels of one row: the two flow planes arrive as one 16-byte load each, every channel
// leaves as one 16-byte store (W % 4 == 0; VEC = 1 otherwise), and the four gathers per channel are independent loads in flight.
template <int MODE, int VEC>
__global__ void __launch_bounds__(128) warp_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                                                   const float* __restrict__ second, int C, int H, int W,
                                                   float* __restrict__ out, int32_t* __restrict__ src, double* loss_accum) {
    const int y = blockIdx.y, b = blockIdx.z;
    const int x0 = (blockIdx.x * 128 + threadIdx.x) * VEC;
    const unsigned hw = (unsigned)H * (unsigned)W;                 // one plane: < 2^31 elements (checked by the host)
    float local = 0.0f;
    if (x0 < W) {
        const unsigned p = (unsigned)y * (unsigned)W + (unsigned)x0;
        const float* fu = flo + (size_t)b * 2 * hw + p;
        float u[VEC], v[VEC];
        if constexpr (VEC == 4) {
            const float4 u4 = __ldg(reinterpret_cast<const float4*>(fu)), v4 = __ldg(reinterpret_cast<const float4*>(fu + hw));
            u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
            v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
        } else {
            u[0] = __ldg(fu);
            v[0] = __ldg(fu + hw);
        }
        unsigned sp[VEC];
        const int iy_row = y;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int ix = src_index((float)(x0 + k), u[k], W);
            const int iy = src_index((float)iy_row, v[k], H);
            sp[k] = (unsigned)iy * (unsigned)W + (unsigned)ix;
            if (src != nullptr) {
                int32_t* s2 = src + ((size_t)b * hw + p + k) * 2;
                s2[0] = iy; s2[1] = ix;
            }
        }
        for (int c = 0; c < C; ++c) {
            const size_t plane = ((size_t)b * C + c) * hw;
            float val[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) val[k] = __ldg(x + plane + sp[k]);
            if constexpr (VEC == 4) {
                *reinterpret_cast<float4*>(out + plane + p) = make_float4(val[0], val[1], val[2], val[3]);
            } else {
                out[plane + p] = val[0];
            }
            if (MODE == 1) {
                if constexpr (VEC == 4) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(second + plane + p));
                    local += (fabsf(val[0] - s4.x) + fabsf(val[1] - s4.y)) + (fabsf(val[2] - s4.z) + fabsf(val[3] - s4.w));
                } else {
                    local += fabsf(val[0] - __ldg(second + plane + p));
                }
            }
        }
    }
    if (MODE == 1) {
        // a thread sums at most 4 C values in fp32; everything above that is double: warp shuffles, then one atomic per block
        double acc = (double)local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        __shared__ double s_part[4];
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(loss_accum, (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
    }
}

__global__ void loss_finish_kernel(const double* accum, double count, float* loss) { *loss = (float)(*accum / count); }

static int warp_shape_ok(const char* what, int B, int C, int H, int W) {
    RRV_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, "%s: negative size", what);
    RRV_REQUIRE((long long)H * W < (1LL << 31) && H <= 65535 && B <= 65535, "%s: %dx%d planes / batch %d exceed the kernel's 32-bit plane indexing",
                what, H, W, B);
    return 0;
}

template <int MODE>
static void launch_warp(const float* x, const float* flo, const float* second, int B, int C, int H, int W, float* out, int32_t* src,
                        double* loss_accum, cudaStream_t st) {
    const bool vec = W % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)flo % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                     (second == nullptr || (uintptr_t)second % 16 == 0);
    if (vec) {
        dim3 grid((unsigned)((W / 4 + 127) / 128), (unsigned)H, (unsigned)B);
        warp_kernel<MODE, 4><<<grid, 128, 0, st>>>(x, flo, second, C, H, W, out, src, loss_accum);
    } else {
        dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)B);
        warp_kernel<MODE, 1><<<grid, 128, 0, st>>>(x, flo, second, C, H, W, out, src, loss_accum);
    }
}

int warp_nearest_border(const float* x, const float* flo, int B, int C, int H, int W, float* out, int32_t* src,
                        cudaStream_t st) {
    RRV_REQUIRE(x && flo && out, "rrv_warp_nearest_border: NULL tensor");
    if (warp_shape_ok("rrv_warp_nearest_border", B, C, H, W)) return 1;
    const long long total = (long long)B * H * W;
    if (total == 0 || C == 0) return 0;
    launch_warp<0>(x, flo, nullptr, B, C, H, W, out, src, nullptr, st);
    return check_launch("warp_kernel<0>");
}

int temporal_loss(const float* first, const float* second, const float* flo, int B, int C, int H, int W, float* warped,
                  double* loss_accum, float* loss, cudaStream_t st) {
    RRV_REQUIRE(first && second && flo && warped && loss_accum && loss, "rrv_temporal_loss: NULL tensor");
    if (warp_shape_ok("rrv_temporal_loss", B, C, H, W)) return 1;
    const long long total = (long long)B * H * W;
    RRV_REQUIRE(total > 0 && C > 0, "rrv_temporal_loss: empty input");
    cudaMemsetAsync(loss_accum, 0, sizeof(double), st);
    launch_warp<1>(first, flo, second, B, C, H, W, warped, nullptr, loss_accum, st);
    if (check_launch("warp_kernel<1>")) return 1;
    loss_finish_kernel<<<1, 1, 0, st>>>(loss_accum, (double)total * C, loss);
    return check_launch("loss_finish_kernel");
}

// d(out)/d(x): nearest sampling copies one source pixel, so grad_x[src] += grad_out[dst];
// the gradient w.r.t. the grid is zero.  grad_x must be zero-initialised by the caller.
__global__ void __launch_bounds__(128) warp_backward_kernel(const float* __restrict__ go, const float* __restrict__ flo,
                                                            int C, int H, int W, float* __restrict__ gx) {
    const int y = blockIdx.y, b = blockIdx.z;
    const int xx = blockIdx.x * 128 + threadIdx.x;
    if (xx >= W) return;
    const unsigned hw = (unsigned)H * (unsigned)W;
    const unsigned p = (unsigned)y * (unsigned)W + (unsigned)xx;
    const float* fu = flo + (size_t)b * 2 * hw + p;
    const int ix = src_index((float)xx, __ldg(fu), W);
    const int iy = src_index((float)y, __ldg(fu + hw), H);
    const unsigned sp = (unsigned)iy * (unsigned)W + (unsigned)ix;
    for (int c = 0; c < C; ++c) {
        const size_t plane = ((size_t)b * C + c) * hw;
        atomicAdd(gx + plane + sp, __ldg(go + plane + p));
    }
}

int warp_backward(const float* go, const float* flo, int B, int C, int H, int W, float* gx, cudaStream_t st) {
    RRV_REQUIRE(go && flo && gx, "rrv_warp_backward: NULL tensor");
    if (warp_shape_ok("rrv_warp_backward", B, C, H, W)) return 1;
    const long long total = (long long)B * H * W;
    if (total == 0 || C == 0) return 0;
    dim3 grid((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)B);
    warp_backward_kernel<<<grid, 128, 0, st>>>(go, flo, C, H, W, gx);
    return check_launch("warp_backward_kernel");
}

}  // namespace rrv
