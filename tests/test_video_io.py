"""The optional video-output side library (include/rerevst_b200_io.h, csrc/mjpg_io.cpp, video_io.MjpgWriter): replaces
test/generate_real_video.py:175-186 (cv2.VideoWriter('MJPG') over frames re-read from disk) with a GPU JPEG encode of the frame
while it is still on the device.  CPU tests: the C-ABI surface and the AVI muxer; GPU tests: the nvJPEG path and the script entry."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rerevst_b200_io.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rrv_[a-z0-9_]+)\s*\(", src)))


def _pattern(h, w, i):
    """A smooth colour image (JPEG-friendly: the tests look at decoded PSNR), different for every i."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ch = [127 + 100 * np.sin(xx / 23.0 + 0.7 * i), 127 + 100 * np.cos(yy / 17.0 - 0.5 * i), 127 + 90 * np.sin((xx + yy) / 31.0 + i)]
    return np.clip(np.stack(ch, -1), 0, 255).astype(np.uint8)


def _read_video(path):
    import cv2
    cap = cv2.VideoCapture(str(path))
    props = (int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), int(round(cap.get(cv2.CAP_PROP_FPS))), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
             int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(fr)
    return props, frames


def test_io_header_declares_what_python_binds():
    from rerevst_code_b200 import video_io
    assert _header_functions() == sorted(video_io.SIGNATURES)


def test_io_library_loads_and_exports_every_declared_symbol():
    from rerevst_code_b200 import video_io
    assert os.path.exists(video_io.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(video_io.LIB_PATH)
    for name in _header_functions():
        assert hasattr(handle, name), name
    assert video_io.lib().rrv_io_abi_version() == video_io.ABI_VERSION == 1


def test_core_package_does_not_load_the_io_library():
    """Nothing on the stylization path needs libnvjpeg: importing the package and its core binding leaves the side library alone."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import rerevst_code_b200; from rerevst_code_b200 import _lib; _lib.lib(); "
            "maps = open('/proc/self/maps').read(); assert 'librerevst_b200.so' in maps; assert 'nvjpeg' not in maps and "
            "'librerevst_b200_io' not in maps" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])


def test_mjpg_muxer_round_trip_on_cpu(tmp_path):
    """rrv_mjpg_write_jpeg: JPEGs made by cv2, container by our muxer, read back by cv2.VideoCapture -- frame count, rate, size
    and content survive (the muxer writes what cv2.VideoWriter('MJPG') writes: RIFF AVI, '00dc' chunks, idx1)."""
    import cv2
    from rerevst_code_b200.video_io import MjpgWriter
    frames = [_pattern(120, 200, i) for i in range(7)]
    path = tmp_path / "t.avi"
    with MjpgWriter(str(path), 24, (200, 120), quality=90) as w:
        for f in frames:
            ok, enc = cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])
            assert ok
            w.write_jpeg(enc.tobytes())
        assert w.frames == 7 and w.bytes > 0
    props, got = _read_video(path)
    assert props == (7, 24, 200, 120) and len(got) == 7
    for a, b in zip(got, frames):
        direct = cv2.imdecode(cv2.imencode(".jpg", b, [cv2.IMWRITE_JPEG_QUALITY, 90])[1], cv2.IMREAD_COLOR)
        assert cv2.PSNR(a, direct) > 32 and cv2.PSNR(a, b) > 30         # the JPEG that went in (another decoder: not bit-equal)
    with pytest.raises(RuntimeError):
        MjpgWriter(str(tmp_path / "missing_dir" / "x.avi"), 24, (200, 120))
    with pytest.raises(RuntimeError):
        MjpgWriter(str(path), 0, (200, 120))


@pytest.mark.gpu
def test_mjpg_writer_encodes_device_frames(tmp_path):
    """video_io.MjpgWriter.write / encode + flush / retrieve on CUDA tensors: the video decodes to the frames (PSNR of a JPEG at
    quality 90), in the order of the flush calls; a retrieved bitstream is a JPEG of the frame."""
    import cv2
    from rerevst_code_b200.video_io import MjpgWriter
    dev = torch.device("cuda", 0)
    h, w = 136, 248
    frames = [_pattern(h, w, i) for i in range(6)]
    path = tmp_path / "g.avi"
    with MjpgWriter(str(path), 30, (w, h), quality=90, states=3) as vw:
        vw.write(torch.from_numpy(frames[0]).to(dev))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        devs = [torch.from_numpy(f).to(dev) for f in frames[1:4]]
        torch.cuda.current_stream(dev).synchronize()
        for k, d in enumerate(devs):                       # three frames in flight on a side stream
            vw.encode(d, k, side)
        for k in range(3):
            vw.flush(k)
        vw.encode(torch.from_numpy(frames[4]).to(dev), 1)
        jpg = vw.retrieve(1)                               # not appended
        assert jpg[:2] == b"\xff\xd8" and cv2.PSNR(cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR), frames[4]) > 30
        vw.write_jpeg(jpg)
        vw.write(torch.from_numpy(frames[5]).to(dev))
        assert vw.frames == 6
        with pytest.raises(ValueError):
            vw.write(torch.zeros((h, w, 3), dtype=torch.float32, device=dev))
    props, got = _read_video(path)
    assert props == (6, 30, w, h) and len(got) == 6
    for a, b in zip(got, frames):
        assert cv2.PSNR(a, b) > 30


@pytest.mark.gpu
def test_generate_real_video_gpu_video_writer(tmp_path, state_dict):
    """generate_real_video.main(video_writer="nvjpeg"): the .avi holds the stylized frames in FILE-NAME order like the reference's
    (generate_real_video.py:176-186), next to the cv2 writer's on the same clip."""
    import cv2
    from rerevst_code_b200 import generate_real_video as grv
    rng = np.random.RandomState(3)
    smooth = lambda h, w: np.clip(rng.rand(h // 8 + 1, w // 8 + 1, 3).repeat(8, 0).repeat(8, 1)[:h, :w] * 255, 0, 255).astype(np.uint8)
    vid = tmp_path / "inputs" / "clip"
    vid.mkdir(parents=True)
    for i in range(9):
        cv2.imwrite(str(vid / f"frame_{i:04d}.png"), smooth(48, 64))
    cv2.imwrite(str(tmp_path / "style.png"), smooth(64, 72))
    ckpt = tmp_path / "net.pth"
    torch.save(state_dict, str(ckpt))
    outs = {}
    for kind in ("cv2", "nvjpeg"):
        out_dir = grv.main(style_img=str(tmp_path / "style.png"), content_video=str(vid / "*.png"), checkpoint_path=str(ckpt),
                           result_frames_path=str(tmp_path / ("rf_" + kind)), result_videos_path=str(tmp_path / ("rv_" + kind)),
                           save_video=True, verbose=False, video_writer=kind, video_quality=95)
        avi = [f for f in os.listdir(tmp_path / ("rv_" + kind)) if f.endswith(".avi")]
        assert len(avi) == 1
        outs[kind] = (out_dir, _read_video(tmp_path / ("rv_" + kind) / avi[0]))
    (dir_a, (props_a, fa)), (dir_b, (props_b, fb)) = outs["cv2"], outs["nvjpeg"]
    assert props_a == props_b == (9, 24, 64, 48)
    written = sorted(os.listdir(dir_b))
    assert written == sorted(os.listdir(dir_a)) and len(written) == 9
    for k, name in enumerate(written):                       # both videos follow the sorted frame files
        png = cv2.imread(os.path.join(dir_b, name))
        assert np.array_equal(png, cv2.imread(os.path.join(dir_a, name)))
        # (random-weight stylizations are noise-like: low PSNR for any JPEG; the GPU encoder at quality 95 must not be worse than
        #  OpenCV's MJPG writer at its default quality)
        assert cv2.PSNR(fb[k], png) > cv2.PSNR(fa[k], png) - 1.0 and cv2.PSNR(fb[k], png) > 20
