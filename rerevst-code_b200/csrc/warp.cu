// Optical-flow warp of the temporal loss: train/loss_networks.py:20-38 (warp) and :106-111
// (TemporalLoss.forward).  HBM-bound gather: 8 B flow + 4C B read + 4C B written per pixel.
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
template <int MODE>
__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                                                   const float* __restrict__ second, int B, int C, int H, int W,
                                                   float* __restrict__ out, int32_t* __restrict__ src, double* loss_accum) {
    const long long hw = (long long)H * W;
    const long long total = (long long)B * hw;
    double local = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int b = (int)(i / hw);
        const long long p = i - (long long)b * hw;
        const int y = (int)(p / W), xx = (int)(p - (long long)y * W);
        const float u = flo[((long long)b * 2) * hw + p];
        const float v = flo[((long long)b * 2 + 1) * hw + p];
        const int ix = src_index((float)xx, u, W);
        const int iy = src_index((float)y, v, H);
        if (src != nullptr) { src[i * 2] = iy; src[i * 2 + 1] = ix; }
        const long long sp = (long long)iy * W + ix;
        for (int c = 0; c < C; ++c) {
            const long long plane = ((long long)b * C + c) * hw;
            const float val = __ldg(x + plane + sp);
            out[plane + p] = val;
            if (MODE == 1) local += (double)fabsf(val - second[plane + p]);
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        __shared__ double s_part[8];
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += s_part[k];
            atomicAdd(loss_accum, t);
        }
    }
}

__global__ void loss_finish_kernel(const double* accum, double count, float* loss) { *loss = (float)(*accum / count); }

static int warp_grid(long long total) { return (int)std::min<long long>((total + 255) / 256, 148LL * 16); }

int warp_nearest_border(const float* x, const float* flo, int B, int C, int H, int W, float* out, int32_t* src,
                        cudaStream_t st) {
    RRV_REQUIRE(x && flo && out, "rrv_warp_nearest_border: NULL tensor");
    const long long total = (long long)B * H * W;
    if (total == 0 || C == 0) return 0;
    warp_kernel<0><<<warp_grid(total), 256, 0, st>>>(x, flo, nullptr, B, C, H, W, out, src, nullptr);
    return check_launch("warp_kernel<0>");
}

int temporal_loss(const float* first, const float* second, const float* flo, int B, int C, int H, int W, float* warped,
                  double* loss_accum, float* loss, cudaStream_t st) {
    RRV_REQUIRE(first && second && flo && warped && loss_accum && loss, "rrv_temporal_loss: NULL tensor");
    const long long total = (long long)B * H * W;
    RRV_REQUIRE(total > 0 && C > 0, "rrv_temporal_loss: empty input");
    cudaMemsetAsync(loss_accum, 0, sizeof(double), st);
    warp_kernel<1><<<warp_grid(total), 256, 0, st>>>(first, flo, second, B, C, H, W, warped, nullptr, loss_accum);
    if (check_launch("warp_kernel<1>")) return 1;
    loss_finish_kernel<<<1, 1, 0, st>>>(loss_accum, (double)total * C, loss);
    return check_launch("loss_finish_kernel");
}

// d(out)/d(x): nearest sampling copies one source pixel, so grad_x[src] += grad_out[dst];
// the gradient w.r.t. the grid is zero.  grad_x must be zero-initialised by the caller.
__global__ void __launch_bounds__(256) warp_backward_kernel(const float* __restrict__ go, const float* __restrict__ flo,
                                                            int B, int C, int H, int W, float* __restrict__ gx) {
    const long long hw = (long long)H * W;
    const long long total = (long long)B * hw;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int b = (int)(i / hw);
        const long long p = i - (long long)b * hw;
        const int y = (int)(p / W), xx = (int)(p - (long long)y * W);
        const int ix = src_index((float)xx, flo[((long long)b * 2) * hw + p], W);
        const int iy = src_index((float)y, flo[((long long)b * 2 + 1) * hw + p], H);
        for (int c = 0; c < C; ++c) {
            const long long plane = ((long long)b * C + c) * hw;
            atomicAdd(gx + plane + (long long)iy * W + ix, go[plane + p]);
        }
    }
}

int warp_backward(const float* go, const float* flo, int B, int C, int H, int W, float* gx, cudaStream_t st) {
    RRV_REQUIRE(go && flo && gx, "rrv_warp_backward: NULL tensor");
    const long long total = (long long)B * H * W;
    if (total == 0 || C == 0) return 0;
    warp_backward_kernel<<<warp_grid(total), 256, 0, st>>>(go, flo, B, C, H, W, gx);
    return check_launch("warp_backward_kernel");
}

}  // namespace rrv
