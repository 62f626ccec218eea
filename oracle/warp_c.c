/* CPU oracle (plain C) for the temporal-loss warp.  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product path.
 *
 * Restates train/loss_networks.py:20-38 (warp) + ATen grid_sampler_2d nearest/border with
 * align_corners=False (torch/include/ATen/native/GridSampler.h: grid_sampler_unnormalize,
 * clip_coordinates; nearbyint rounding).  Every float op is a separately rounded fp32 op:
 * compile with -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline int src_index(float pos, float flow, int size) {
    volatile float v = pos - flow;                 /* loss_networks.py:30 */
    volatile float t = 2.0f * v;                   /* :33-34 */
    int den = size - 1 > 1 ? size - 1 : 1;
    volatile float q = t / (float)den;
    volatile float g = q - 1.0f;
    volatile float a = g + 1.0f;                   /* grid_sampler_unnormalize */
    volatile float m = a * (float)size;
    volatile float s = m - 1.0f;
    volatile float u = s / 2.0f;
    float c = u > 0.0f ? u : 0.0f;                 /* clip_coordinates */
    float lim = (float)(size - 1);
    c = c < lim ? c : lim;
    return (int)nearbyintf(c);                     /* ties to even */
}

/* flo: [B,2,H,W]; iy, ix: [B,H,W] */
void rrv_oracle_warp_indices(const float* flo, int B, int H, int W, int32_t* iy, int32_t* ix) {
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t p = ((size_t)b * H + y) * W + x;
                float u = flo[((size_t)b * 2 + 0) * H * W + (size_t)y * W + x];
                float v = flo[((size_t)b * 2 + 1) * H * W + (size_t)y * W + x];
                ix[p] = src_index((float)x, u, W);
                iy[p] = src_index((float)y, v, H);
            }
}

/* x, out: [B,C,H,W]; returns mean |warp(x) - second| accumulated in double (loss_networks.py:106-111);
 * second may be NULL (then 0 is returned). */
double rrv_oracle_warp_l1(const float* x, const float* flo, const float* second,
                          int B, int C, int H, int W, float* out) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                float u = flo[((size_t)b * 2 + 0) * H * W + (size_t)y * W + xx];
                float v = flo[((size_t)b * 2 + 1) * H * W + (size_t)y * W + xx];
                int sx = src_index((float)xx, u, W), sy = src_index((float)y, v, H);
                for (int c = 0; c < C; ++c) {
                    size_t o = (((size_t)b * C + c) * H + y) * W + xx;
                    float val = x[(((size_t)b * C + c) * H + sy) * W + sx];
                    out[o] = val;
                    if (second) acc += fabs((double)val - (double)second[o]);
                }
            }
    return second ? acc / ((double)B * C * H * W) : 0.0;
}
