// Motion-JPEG .avi writer for frames that live on the device (include/rerevst_b200_io.h).
//
// Replaces test/generate_real_video.py:175-186 (frames re-read from disk, cv2.VideoWriter('MJPG') = a CPU JPEG encoder): the
// uint8 BGR frame the RGB head wrote is encoded by nvJPEG on the GPU and only the bitstream is downloaded; the muxer writes the
// container OpenCV writes for fourcc 'MJPG' (RIFF AVI, one video stream, '00dc' chunks, idx1).  Host code only: the CUDA work
// is nvJPEG's.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/rerevst_b200_io.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

struct IndexEntry {
    uint32_t offset, length;       // offset of the chunk header relative to the 'movi' fourcc
};

struct Writer {
    FILE* f = nullptr;
    int width = 0, height = 0, fps = 0, quality = 0;
    long movi_fourcc_pos = 0;      // file offset of the 'movi' fourcc
    std::vector<IndexEntry> index;
    int64_t bytes = 0;
    uint32_t max_chunk = 0;
    // nvJPEG, created on the first device frame (the muxer alone needs no GPU)
    nvjpegHandle_t nv = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    std::vector<nvjpegEncoderState_t> states;
    int n_states = 1;
    std::vector<unsigned char> scratch;
};

void put32(FILE* f, uint32_t v) {
    unsigned char b[4] = {(unsigned char)v, (unsigned char)(v >> 8), (unsigned char)(v >> 16), (unsigned char)(v >> 24)};
    fwrite(b, 1, 4, f);
}
void put16(FILE* f, uint16_t v) {
    unsigned char b[2] = {(unsigned char)v, (unsigned char)(v >> 8)};
    fwrite(b, 1, 2, f);
}
void put4cc(FILE* f, const char* s) { fwrite(s, 1, 4, f); }

// Header with the counts known so far (called once with zeros at open, again at close with the real numbers).
void write_headers(Writer* w, uint32_t riff_size, uint32_t movi_size, uint32_t frames) {
    FILE* f = w->f;
    fseek(f, 0, SEEK_SET);
    put4cc(f, "RIFF"); put32(f, riff_size); put4cc(f, "AVI ");
    put4cc(f, "LIST"); put32(f, 4 + (8 + 56) + (12 + (8 + 56) + (8 + 40))); put4cc(f, "hdrl");
    put4cc(f, "avih"); put32(f, 56);
    put32(f, (uint32_t)(1000000 / (w->fps > 0 ? w->fps : 1)));       // dwMicroSecPerFrame
    put32(f, w->max_chunk * (uint32_t)(w->fps > 0 ? w->fps : 1));    // dwMaxBytesPerSec
    put32(f, 0);                                                     // dwPaddingGranularity
    put32(f, 0x10);                                                  // dwFlags: AVIF_HASINDEX
    put32(f, frames);                                                // dwTotalFrames
    put32(f, 0);                                                     // dwInitialFrames
    put32(f, 1);                                                     // dwStreams
    put32(f, w->max_chunk);                                          // dwSuggestedBufferSize
    put32(f, (uint32_t)w->width); put32(f, (uint32_t)w->height);
    put32(f, 0); put32(f, 0); put32(f, 0); put32(f, 0);              // dwReserved
    put4cc(f, "LIST"); put32(f, 4 + (8 + 56) + (8 + 40)); put4cc(f, "strl");
    put4cc(f, "strh"); put32(f, 56);
    put4cc(f, "vids"); put4cc(f, "MJPG");
    put32(f, 0);                                                     // dwFlags
    put16(f, 0); put16(f, 0);                                        // wPriority, wLanguage
    put32(f, 0);                                                     // dwInitialFrames
    put32(f, 1); put32(f, (uint32_t)w->fps);                         // dwScale, dwRate
    put32(f, 0); put32(f, frames);                                   // dwStart, dwLength
    put32(f, w->max_chunk);                                          // dwSuggestedBufferSize
    put32(f, 0xFFFFFFFFu);                                           // dwQuality (default)
    put32(f, 0);                                                     // dwSampleSize
    put16(f, 0); put16(f, 0); put16(f, (uint16_t)w->width); put16(f, (uint16_t)w->height);      // rcFrame
    put4cc(f, "strf"); put32(f, 40);
    put32(f, 40); put32(f, (uint32_t)w->width); put32(f, (uint32_t)w->height);
    put16(f, 1); put16(f, 24);                                       // biPlanes, biBitCount
    put4cc(f, "MJPG");
    put32(f, (uint32_t)(w->width * w->height * 3));
    put32(f, 0); put32(f, 0); put32(f, 0); put32(f, 0);
    put4cc(f, "LIST"); put32(f, movi_size);
    w->movi_fourcc_pos = ftell(f);
    put4cc(f, "movi");
}

int append_chunk(Writer* w, const void* jpeg, int64_t n) {
    if (n <= 0 || n > 0x7fffffff) return fail("rrv_mjpg: bad JPEG size %lld", (long long)n);
    FILE* f = w->f;
    fseek(f, 0, SEEK_END);
    const long pos = ftell(f);
    if ((unsigned long long)pos + (unsigned long long)n + 16ull * (w->index.size() + 2) > 0xF0000000ull)
        return fail("rrv_mjpg: the file would pass the 4 GB limit of a RIFF AVI");
    put4cc(f, "00dc");
    put32(f, (uint32_t)n);
    if (fwrite(jpeg, 1, (size_t)n, f) != (size_t)n) return fail("rrv_mjpg: write failed");
    if (n & 1) fputc(0, f);
    w->index.push_back(IndexEntry{(uint32_t)(pos - w->movi_fourcc_pos), (uint32_t)n});
    w->bytes += n;
    if ((uint32_t)n > w->max_chunk) w->max_chunk = (uint32_t)n;
    return 0;
}

int ensure_nvjpeg(Writer* w, cudaStream_t st) {
    if (w->nv != nullptr) return 0;
    nvjpegStatus_t s = nvjpegCreateSimple(&w->nv);
    if (s != NVJPEG_STATUS_SUCCESS) { w->nv = nullptr; return fail("nvjpegCreateSimple failed: %d", (int)s); }
    s = nvjpegEncoderParamsCreate(w->nv, &w->params, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderParamsCreate failed: %d", (int)s);
    nvjpegEncoderParamsSetQuality(w->params, w->quality, st);
    nvjpegEncoderParamsSetSamplingFactors(w->params, NVJPEG_CSS_420, st);
    nvjpegEncoderParamsSetOptimizedHuffman(w->params, 0, st);
    w->states.resize((size_t)w->n_states, nullptr);
    for (int i = 0; i < w->n_states; ++i) {
        s = nvjpegEncoderStateCreate(w->nv, &w->states[(size_t)i], st);
        if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncoderStateCreate failed: %d", (int)s);
    }
    return 0;
}

}  // namespace

extern "C" {

int rrv_io_abi_version(void) { return RRV_IO_ABI_VERSION; }
const char* rrv_io_last_error(void) { return g_err; }

void* rrv_mjpg_open(const char* path, int width, int height, int fps, int quality, int n_states) {
    if (path == nullptr || width <= 0 || height <= 0 || width > 65535 || height > 65535 || fps <= 0 || quality < 1 || quality > 100 ||
        n_states < 1 || n_states > 8) {
        fail("rrv_mjpg_open: bad arguments (%dx%d, %d fps, quality %d, %d states)", width, height, fps, quality, n_states);
        return nullptr;
    }
    FILE* f = fopen(path, "wb+");
    if (f == nullptr) {
        fail("rrv_mjpg_open: cannot open %s", path);
        return nullptr;
    }
    Writer* w = new Writer();
    w->f = f; w->width = width; w->height = height; w->fps = fps; w->quality = quality; w->n_states = n_states;
    write_headers(w, 0, 4, 0);
    return w;
}

int rrv_mjpg_encode(void* writer, int state, const void* dev_bgr_u8, void* stream) {
    Writer* w = (Writer*)writer;
    if (w == nullptr || dev_bgr_u8 == nullptr) return fail("rrv_mjpg_encode: NULL argument");
    if (state < 0 || state >= w->n_states) return fail("rrv_mjpg_encode: state %d of %d", state, w->n_states);
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_nvjpeg(w, st)) return 1;
    nvjpegImage_t img;
    memset(&img, 0, sizeof(img));
    img.channel[0] = (unsigned char*)const_cast<void*>(dev_bgr_u8);
    img.pitch[0] = (size_t)w->width * 3;
    const nvjpegStatus_t s = nvjpegEncodeImage(w->nv, w->states[(size_t)state], w->params, &img, NVJPEG_INPUT_BGRI, w->width, w->height, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeImage failed: %d", (int)s);
    return 0;
}

int rrv_mjpg_flush(void* writer, int state, void* stream) {
    Writer* w = (Writer*)writer;
    if (w == nullptr || w->nv == nullptr) return fail("rrv_mjpg_flush: nothing was encoded");
    if (state < 0 || state >= w->n_states) return fail("rrv_mjpg_flush: state %d of %d", state, w->n_states);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaStreamSynchronize(st);          // the encode enqueued by rrv_mjpg_encode has finished
    if (e != cudaSuccess) return fail("rrv_mjpg_flush: %s", cudaGetErrorString(e));
    size_t len = 0;
    nvjpegStatus_t s = nvjpegEncodeRetrieveBitstream(w->nv, w->states[(size_t)state], nullptr, &len, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream(size) failed: %d", (int)s);
    if (w->scratch.size() < len) w->scratch.resize(len);
    s = nvjpegEncodeRetrieveBitstream(w->nv, w->states[(size_t)state], w->scratch.data(), &len, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream failed: %d", (int)s);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("rrv_mjpg_flush: %s", cudaGetErrorString(e));
    return append_chunk(w, w->scratch.data(), (int64_t)len);
}

int rrv_mjpg_retrieve(void* writer, int state, void* stream, void* buf, int64_t capacity, int64_t* nbytes) {
    Writer* w = (Writer*)writer;
    if (w == nullptr || w->nv == nullptr || nbytes == nullptr) return fail("rrv_mjpg_retrieve: nothing was encoded");
    if (state < 0 || state >= w->n_states) return fail("rrv_mjpg_retrieve: state %d of %d", state, w->n_states);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("rrv_mjpg_retrieve: %s", cudaGetErrorString(e));
    size_t len = 0;
    nvjpegStatus_t s = nvjpegEncodeRetrieveBitstream(w->nv, w->states[(size_t)state], nullptr, &len, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream(size) failed: %d", (int)s);
    *nbytes = (int64_t)len;
    if (buf == nullptr) return 0;                  // size query
    if ((int64_t)len > capacity) return fail("rrv_mjpg_retrieve: buffer of %lld bytes, bitstream of %lld", (long long)capacity, (long long)len);
    s = nvjpegEncodeRetrieveBitstream(w->nv, w->states[(size_t)state], (unsigned char*)buf, &len, st);
    if (s != NVJPEG_STATUS_SUCCESS) return fail("nvjpegEncodeRetrieveBitstream failed: %d", (int)s);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail("rrv_mjpg_retrieve: %s", cudaGetErrorString(e));
    *nbytes = (int64_t)len;
    return 0;
}

int rrv_mjpg_write_jpeg(void* writer, const void* jpeg, int64_t nbytes) {
    Writer* w = (Writer*)writer;
    if (w == nullptr || jpeg == nullptr) return fail("rrv_mjpg_write_jpeg: NULL argument");
    return append_chunk(w, jpeg, nbytes);
}

int64_t rrv_mjpg_frames(void* writer) { return writer ? (int64_t)((Writer*)writer)->index.size() : -1; }
int64_t rrv_mjpg_bytes(void* writer) { return writer ? ((Writer*)writer)->bytes : -1; }

int rrv_mjpg_close(void* writer) {
    Writer* w = (Writer*)writer;
    if (w == nullptr) return fail("rrv_mjpg_close: NULL writer");
    FILE* f = w->f;
    fseek(f, 0, SEEK_END);
    const long movi_end = ftell(f);
    put4cc(f, "idx1");
    put32(f, (uint32_t)(16 * w->index.size()));
    for (const IndexEntry& e : w->index) {
        put4cc(f, "00dc");
        put32(f, 0x10);                     // AVIIF_KEYFRAME
        put32(f, e.offset);
        put32(f, e.length);
    }
    const long end = ftell(f);
    write_headers(w, (uint32_t)(end - 8), (uint32_t)(movi_end - w->movi_fourcc_pos), (uint32_t)w->index.size());
    const int rc = fclose(f) == 0 ? 0 : fail("rrv_mjpg_close: close failed");
    for (nvjpegEncoderState_t s : w->states)
        if (s) nvjpegEncoderStateDestroy(s);
    if (w->params) nvjpegEncoderParamsDestroy(w->params);
    if (w->nv) nvjpegDestroy(w->nv);
    delete w;
    return rc;
}

}  // extern "C"
