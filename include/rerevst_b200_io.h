/* rerevst_b200_io.h -- C ABI of the optional video-output side library (librerevst_b200_io.so).
 *
 * Replaces the tail of the reference's inference script, test/generate_real_video.py:175-186: there the stylized frames are
 * written to disk, read back with cv2.imread and pushed through cv2.VideoWriter('MJPG') -- a CPU JPEG encoder -- into an .avi.
 * Here a frame that is still on the device (the uint8 HWC BGR frame the RGB head's epilogue wrote, RRV_OUT_BGR_U8) is JPEG-encoded
 * on the GPU (nvJPEG: colour conversion, DCT, quantisation and entropy coding on the device) and only the compressed bitstream
 * (~10 % of the frame) crosses PCIe; the muxer below writes the same container (RIFF AVI, one 'MJPG' video stream, idx1 index).
 * The B200 has no NVENC; Motion-JPEG is also what the reference writes.
 *
 * A separate shared library on purpose: it links libnvjpeg, the core library (rerevst_b200.h) links nothing but the CUDA runtime,
 * and nothing on the stylization path loads this one.  Plain pointers and sizes, no torch types; device pointers are CUDA device
 * pointers, `stream` is a cudaStream_t.  Functions return 0 on success; rrv_io_last_error() gives the message.
 */
#ifndef RERVST_B200_IO_H
#define RERVST_B200_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRV_IO_ABI_VERSION 1

int         rrv_io_abi_version(void);
const char* rrv_io_last_error(void);

/* cv2.VideoWriter(path, fourcc('MJPG'), fps, (width, height)) (generate_real_video.py:181-183).  quality: 1..100 (OpenCV's
 * VIDEOWRITER_PROP_QUALITY default for MJPG is 75; chroma 4:2:0 like OpenCV's encoder).  n_states encoder states let that many
 * frames be in flight between rrv_mjpg_encode and rrv_mjpg_flush (1..8).  Returns NULL on failure. */
void* rrv_mjpg_open(const char* path, int width, int height, int fps, int quality, int n_states);

/* videoWriter.write(frame) (:185), first half: enqueue the JPEG encode of a device frame [height][width][3] uint8 BGR on `stream`
 * with encoder state `state` (asynchronous; the frame must stay valid until the matching rrv_mjpg_flush returns). */
int rrv_mjpg_encode(void* writer, int state, const void* dev_bgr_u8, void* stream);

/* ... second half: wait for `stream`, fetch state `state`'s bitstream and append it as the next frame of the file.  Frames appear
 * in the file in the order of the flush calls. */
int rrv_mjpg_flush(void* writer, int state, void* stream);

/* Instead of rrv_mjpg_flush: wait for `stream` and copy state `state`'s bitstream to the caller (buf == NULL: only its size in
 * *nbytes) without appending it -- the reference writes the video from the SORTED list of frame files (:176-177) while it processes
 * frames in glob order, so a caller that streams frames through the encoder keeps the bitstreams and appends them in file-name
 * order with rrv_mjpg_write_jpeg. */
int rrv_mjpg_retrieve(void* writer, int state, void* stream, void* buf, int64_t capacity, int64_t* nbytes);

/* Append an already encoded JPEG (host bytes) as the next frame: the muxer alone, no GPU involved. */
int rrv_mjpg_write_jpeg(void* writer, const void* jpeg, int64_t nbytes);

/* Frames appended so far / compressed bytes appended so far. */
int64_t rrv_mjpg_frames(void* writer);
int64_t rrv_mjpg_bytes(void* writer);

/* videoWriter.release() (:186): writes the index, patches the headers, closes the file and frees the encoder. */
int rrv_mjpg_close(void* writer);

#ifdef __cplusplus
}
#endif
#endif
