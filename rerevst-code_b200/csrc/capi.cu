// extern "C" entry points of librerevst_b200.so (see include/rerevst_b200.h).
#include <cstdarg>
#include <cstdio>
#include <atomic>

#include "rrv_common.cuh"

namespace rrv {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
int g_lo_fp16 = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

int conv2d_ffma(const rrv_conv* p, cudaStream_t st);
int conv2d_tc(const rrv_conv* p, cudaStream_t st);
long long tc_weight_bytes(int Cin, int Cout, int ksize, int ups);
int pack_weights_tc(const float* w, int Cin, int Cout, int ksize, int ups, void* blob, cudaStream_t st);
int tc_tune(int max_bn, int mt);
int tc_tune_pair(int enable, int min_bn);
int tc_tune_merge(int enable);
int tc_tune_pdl(int enable);
int tc_timeline(unsigned long long* buf, int nslots);
int first_layer(const void* src, int src_kind, int gray, int N, int H, int W, const float* w, const float* bias,
                void* out_hi, void* out_lo, float* out_f32, cudaStream_t st);
int reflect_pad_u8(const void* src, int N, int H, int W, int top, int left, int PH, int PW, void* dst, cudaStream_t st);
int maxpool2x2(const void* in_hi, const void* in_lo, int N, int H, int W, int C, void* out_hi, void* out_lo, cudaStream_t st);
int pointwise(const float* in, long long in_bs, int N, int H, int W, int C, const rrv_epilogue* ep, int out_mode,
              void* out_hi, void* out_lo, float* out_f32, double* stats, int stats_minmax, cudaStream_t st);
int planes_to_nchw(const void* hi, const void* lo, int N, int H, int W, int C, float* out, cudaStream_t st);
int nchw_to_planes(const float* in, int N, int H, int W, int C, void* hi, void* lo, cudaStream_t st);
int postprocess_bgr(const float* in, int N, int H, int W, int y0, int x0, int h, int w, float* out, cudaStream_t st);
int postprocess_bgr_u8(const float* in, int N, int H, int W, int y0, int x0, int h, int w, uint8_t* out, cudaStream_t st);
int pack_weights_f32(const float* w, int Cin, int Cout, int ksize, int Cin_pad, int Cout_pad, float* out, cudaStream_t st);
int relu_backward(const float* g, const float* g2, const float* y, long long n, void* hi, void* lo, cudaStream_t st);
int maxpool2x2_backward(const float* g, const float* y, int N, int H, int W, int C, float* gx, cudaStream_t st);
int fold_filter(const float* wf1, const float* wf2, const float* down_w, const float* down_b, const float* up_w, void* down_blob,
                float* down_bias, void* up_blob, cudaStream_t st);
int channel_stats(const float* x, long long npix, int C, double* part, cudaStream_t st);
int stats_init(double* part, int C, double count, cudaStream_t st);
int conv3x3_output_sum(const void* in_hi, const void* in_lo, int N, int H, int W, int Cin, const float* w_oihw, const float* bias,
                       int Cout, double* scratch, double* part, cudaStream_t st);
int stats_sums_to_m2(double* part, int C, cudaStream_t st);
int stats_merge(const double* parts, int nparts, int C, double* merged, cudaStream_t st);
int stats_finalize(const double* part, int C, int kind, float eps, float* out, cudaStream_t st);
int filter_fc(const float* w, const float* b, const float* c_mean, const float* s_mean, float* out, cudaStream_t st);
int warp_nearest_border(const float* x, const float* flo, int B, int C, int H, int W, float* out, int32_t* src, cudaStream_t st);
int temporal_loss(const float* first, const float* second, const float* flo, int B, int C, int H, int W, float* warped,
                  double* loss_accum, float* loss, cudaStream_t st);
int warp_backward(const float* go, const float* flo, int B, int C, int H, int W, float* gx, cudaStream_t st);

}  // namespace rrv

using namespace rrv;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int rrv_abi_version(void) { return RRV_ABI_VERSION; }
const char* rrv_last_error(void) { return g_err; }
uint64_t rrv_launch_count(void) { return g_launches.load(); }
int rrv_set_lo_format(int fmt) {
    RRV_REQUIRE(fmt == 0 || fmt == 1, "rrv_set_lo_format: fmt must be 0 (bf16) or 1 (fp16)");
    g_lo_fp16 = fmt;
    return 0;
}
int rrv_get_lo_format(void) { return g_lo_fp16; }

int rrv_conv2d(const rrv_conv* p, int impl, void* stream) {
    RRV_REQUIRE(p != nullptr, "rrv_conv2d: NULL descriptor");
    if (impl == RRV_IMPL_FFMA) return conv2d_ffma(p, ST(stream));
    if (impl == RRV_IMPL_TCGEN05) return conv2d_tc(p, ST(stream));
    set_error("rrv_conv2d: unknown impl %d", impl);
    return 1;
}
int64_t rrv_tc_weight_bytes(int Cin, int Cout, int ksize, int ups) { return tc_weight_bytes(Cin, Cout, ksize, ups); }
int rrv_pack_weights_tc(const float* w, int Cin, int Cout, int ksize, int ups, void* blob, void* stream) {
    return pack_weights_tc(w, Cin, Cout, ksize, ups, blob, ST(stream));
}
int rrv_tc_tune(int max_bn, int mt) { return tc_tune(max_bn, mt); }
int rrv_tc_tune_pair(int enable, int min_bn) { return tc_tune_pair(enable, min_bn); }
int rrv_tc_tune_merge(int enable) { return tc_tune_merge(enable); }
int rrv_tc_tune_pdl(int enable) { return tc_tune_pdl(enable); }
int rrv_tc_timeline(void* buf, int nslots) { return tc_timeline((unsigned long long*)buf, nslots); }
int rrv_pack_weights_f32(const float* w, int Cin, int Cout, int ksize, int Cin_pad, int Cout_pad, float* out, void* stream) {
    return pack_weights_f32(w, Cin, Cout, ksize, Cin_pad, Cout_pad, out, ST(stream));
}
int rrv_relu_backward(const float* g, const float* g2, const float* y, int64_t n, void* out_hi, void* out_lo, void* stream) {
    return relu_backward(g, g2, y, n, out_hi, out_lo, ST(stream));
}
int rrv_maxpool2x2_backward(const float* g, const float* y, int N, int H, int W, int C, float* gx, void* stream) {
    return maxpool2x2_backward(g, y, N, H, W, C, gx, ST(stream));
}
int rrv_fold_filter(const float* wf1, const float* wf2, const float* down_w, const float* down_b, const float* up_w,
                    void* down_blob, float* down_bias, void* up_blob, void* stream) {
    return fold_filter(wf1, wf2, down_w, down_b, up_w, down_blob, down_bias, up_blob, ST(stream));
}
int rrv_first_layer(const void* src, int src_kind, int gray, int N, int H, int W, const float* w, const float* bias,
                    void* out_hi, void* out_lo, float* out_f32, void* stream) {
    return first_layer(src, src_kind, gray, N, H, W, w, bias, out_hi, out_lo, out_f32, ST(stream));
}
int rrv_reflect_pad_u8(const void* src, int N, int H, int W, int top, int left, int PH, int PW, void* dst, void* stream) {
    return reflect_pad_u8(src, N, H, W, top, left, PH, PW, dst, ST(stream));
}
int rrv_maxpool2x2(const void* in_hi, const void* in_lo, int N, int H, int W, int C, void* out_hi, void* out_lo, void* stream) {
    return maxpool2x2(in_hi, in_lo, N, H, W, C, out_hi, out_lo, ST(stream));
}
int rrv_pointwise(const float* in, int64_t in_bs, int N, int H, int W, int C, const rrv_epilogue* ep, int out_mode,
                  void* out_hi, void* out_lo, float* out_f32, void* stream) {
    return pointwise(in, in_bs, N, H, W, C, ep, out_mode, out_hi, out_lo, out_f32, nullptr, 0, ST(stream));
}
int rrv_pointwise_stats(const float* in, int64_t in_bs, int N, int H, int W, int C, const rrv_epilogue* ep, int out_mode,
                        void* out_hi, void* out_lo, float* out_f32, double* stats, int stats_minmax, void* stream) {
    RRV_REQUIRE(stats != nullptr, "rrv_pointwise_stats: stats is NULL");
    return pointwise(in, in_bs, N, H, W, C, ep, out_mode, out_hi, out_lo, out_f32, stats, stats_minmax, ST(stream));
}
int rrv_conv3x3_output_sum(const void* in_hi, const void* in_lo, int N, int H, int W, int Cin, const float* w_oihw,
                           const float* bias, int Cout, double* scratch, double* part, void* stream) {
    return conv3x3_output_sum(in_hi, in_lo, N, H, W, Cin, w_oihw, bias, Cout, scratch, part, ST(stream));
}
int rrv_stats_init(double* part, int C, double count, void* stream) { return stats_init(part, C, count, ST(stream)); }
int rrv_stats_sums_to_m2(double* part, int C, void* stream) { return stats_sums_to_m2(part, C, ST(stream)); }
int rrv_planes_to_nchw(const void* hi, const void* lo, int N, int H, int W, int C, float* out, void* stream) {
    return planes_to_nchw(hi, lo, N, H, W, C, out, ST(stream));
}
int rrv_nchw_to_planes(const float* in, int N, int H, int W, int C, void* hi, void* lo, void* stream) {
    return nchw_to_planes(in, N, H, W, C, hi, lo, ST(stream));
}
int rrv_postprocess_bgr(const float* in, int N, int H, int W, int y0, int x0, int h, int w, float* out, void* stream) {
    return postprocess_bgr(in, N, H, W, y0, x0, h, w, out, ST(stream));
}
int rrv_postprocess_bgr_u8(const float* in, int N, int H, int W, int y0, int x0, int h, int w, uint8_t* out, void* stream) {
    return postprocess_bgr_u8(in, N, H, W, y0, x0, h, w, out, ST(stream));
}
int rrv_channel_stats(const float* x, int64_t npix, int C, double* part, void* stream) {
    return channel_stats(x, npix, C, part, ST(stream));
}
int rrv_stats_merge(const double* parts, int nparts, int C, double* merged, void* stream) {
    return stats_merge(parts, nparts, C, merged, ST(stream));
}
int rrv_stats_finalize(const double* part, int C, int kind, float eps, float* out, void* stream) {
    return stats_finalize(part, C, kind, eps, out, ST(stream));
}
int rrv_filter_fc(const float* w, const float* b, const float* c_mean, const float* s_mean, float* out, void* stream) {
    return filter_fc(w, b, c_mean, s_mean, out, ST(stream));
}
int rrv_warp_nearest_border(const float* x, const float* flo, int B, int C, int H, int W, float* out, int32_t* src_index,
                            void* stream) {
    return warp_nearest_border(x, flo, B, C, H, W, out, src_index, ST(stream));
}
int rrv_temporal_loss(const float* first, const float* second, const float* flo, int B, int C, int H, int W, float* warped,
                      double* loss_accum, float* loss, void* stream) {
    return temporal_loss(first, second, flo, B, C, H, W, warped, loss_accum, loss, ST(stream));
}
int rrv_warp_backward(const float* grad_out, const float* flo, int B, int C, int H, int W, float* grad_x, void* stream) {
    return warp_backward(grad_out, flo, B, C, H, W, grad_x, ST(stream));
}

}  // extern "C"
