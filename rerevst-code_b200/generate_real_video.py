"""Drop-in entry point for the reference's inference script (``test/generate_real_video.py``).

The reference is a module-level script; here the same steps are a function with the script's
constants (:21-43) as defaults, so ``python -m rerevst_code_b200.generate_real_video`` reproduces
``python generate_real_video.py`` and a caller can also pass its own paths:

  read style → Stylization → prepare_style                               (:91-99)
  global pre-pass: clean, add every 8th frame + the last one, compute     (:129-148; frames in
                   glob order and UNPADDED, quirk Q3 -- kept)
  per frame: read → reflect-pad by 64 to a multiple of 64 → transfer → crop → imwrite   (:152-171)
  optional MJPG .avi of the written frames                                (:175-186; ``video_writer="nvjpeg"``: the frames
                   are JPEG-encoded on the GPU while they are still there -- video_io.MjpgWriter -- instead of being read
                   back from disk and encoded by cv2 on the CPU)

Host image I/O (cv2) stays on the host like in the reference; the frame loop uses
``Stylization.transfer_stream`` so the upload of frame i+1 and the download of frame i-1 overlap the
kernels of frame i.
"""
from __future__ import annotations

import glob
import os
import sys


def padded_size(h, w):
    """ReshapeTool (:61-78): +128, rounded up to a multiple of 64."""
    nh, nw = h + 128, w + 128
    nh += (64 - nh % 64) % 64
    nw += (64 - nw % 64) % 64
    return nh, nw


class ReshapeTool:
    """Same contract as the reference class (:61-83): the padded size is fixed by the first frame."""

    def __init__(self):
        self.record_H = 0
        self.record_W = 0

    def process(self, img):
        import cv2
        H, W, _ = img.shape
        if self.record_H == 0 and self.record_W == 0:
            self.record_H, self.record_W = padded_size(H, W)
        return cv2.copyMakeBorder(img, 64, self.record_H - 64 - H, 64, self.record_W - 64 - W, cv2.BORDER_REFLECT)


def main(style_img="./inputs/plum_flower.jpg", content_video="./inputs/ambush_4/*.png",
         checkpoint_path="./Model/style_net-TIP-final.pth", cuda=True, use_Global=True, save_video=True, fps=24,
         result_frames_path="./result_frames/", result_videos_path="./result_videos/", precision="x3", verbose=True,
         video_writer="cv2", video_quality=75):
    import cv2
    from .framework import Stylization

    say = print if verbose else (lambda *a, **k: None)
    for d in (result_frames_path, result_videos_path):
        os.makedirs(d, exist_ok=True)
    if not os.path.exists(style_img):
        sys.exit("Style image %s not exists" % style_img)
    style = cv2.imread(style_img)

    framework = Stylization(checkpoint_path, cuda, use_Global, precision=precision)
    framework.prepare_style(style)

    frame_list = glob.glob(content_video)                      # unsorted, like the reference (:102)
    style_name = os.path.basename(style_img).split(".")[0]
    video_name = content_video.split("/")[-2]
    name = "ReReVST-" + style_name + "-" + video_name + ("" if use_Global else "-no-global")
    out_dir = os.path.join(result_frames_path, name)
    os.makedirs(out_dir, exist_ok=True)
    frame_num = len(frame_list)

    if use_Global:
        say("Preparations for Sequence-Level Global Feature Sharing")
        framework.clean()
        interval = 8
        sample_sum = (frame_num - 1) // interval
        for s in range(sample_sum):
            say("Add frame %d , %d frames in total" % (s, sample_sum))
            framework.add(cv2.imread(frame_list[s * interval]))
        framework.add(cv2.imread(frame_list[-1]))
        say("Computing global features")
        framework.compute()
        say("Preparations finish!")

    # ReshapeTool.process (:66-83) runs on the device: the raw frame is uploaded and reflect-padded there (rrv_reflect_pad_u8)
    first = cv2.imread(frame_list[0]) if frame_num else None
    pad_to = padded_size(first.shape[0], first.shape[1]) if first is not None else None
    raw_frames = (cv2.imread(frame_list[i]) for i in range(frame_num))
    # out_dtype="u8": the frame arrives as the uint8 image cv2.imwrite would make of the reference's float32 result (:170,
    # saturate_cast<uchar>(cvRound(v))) -- the same file, a quarter of the download
    if video_writer not in ("cv2", "nvjpeg"):
        raise ValueError("video_writer must be 'cv2' (the reference's CPU encoder) or 'nvjpeg'")
    gpu_video = save_video and frame_num and video_writer == "nvjpeg"
    sink, jpegs = None, {}
    if gpu_video:
        # Motion-JPEG on the GPU: every finished frame is encoded while it is still on the device (three encoder states rotate
        # with the frames in flight); the bitstreams are kept and muxed in FILE-NAME order below, like the reference's sorted
        # list of written frames (:176-177) -- frames are processed in glob order
        from .video_io import MjpgWriter
        vw = MjpgWriter(os.path.join(result_videos_path, name + ".avi"), fps, (first.shape[1], first.shape[0]), quality=video_quality, states=4)

        def sink(i, dev_frame, stream):
            k = i % 4
            vw.encode(dev_frame, k, stream)

            def fin():
                jpegs[os.path.basename(frame_list[i])] = vw.retrieve(k)
            return fin

    for i, styled in enumerate(framework.transfer_stream(raw_frames, pad_to=pad_to, copy=False, out_dtype="u8", depth=4,
                                                         device_sink=sink)):     # written out before the next one
        say("Stylizing frame %d" % i)
        cv2.imwrite(os.path.join(out_dir, os.path.basename(frame_list[i])), styled)

    if gpu_video:
        for fname in sorted(jpegs):
            vw.write_jpeg(jpegs[fname])
        vw.release()
    elif save_video and frame_num:
        written = sorted(glob.glob(os.path.join(out_dir, "*.*")))
        demo = cv2.imread(written[0])
        writer = cv2.VideoWriter(os.path.join(result_videos_path, name + ".avi"), cv2.VideoWriter_fourcc(*"MJPG"), fps,
                                 (demo.shape[1], demo.shape[0]))
        for f in written:
            writer.write(cv2.imread(f))
        writer.release()
    return out_dir


if __name__ == "__main__":
    main()
