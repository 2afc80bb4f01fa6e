"""Where the time of the raw-YUV path goes (4K, 120 frames, 8-bit 4:2:0): frames resident in HBM, in pinned host memory,
in pageable host memory, and in a file mapping (what video_source_yuv_file hands to the library)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import colorvideovdp_b200 as cv  # noqa: E402
from colorvideovdp_b200 import _native as N  # noqa: E402

dev = torch.device("cuda:0")
H, W, F, fps = 2160, 3840, 120, 60.0
m = cv.cvvdp(display_name="standard_4k", device=dev)


class A:
    pass


args = A()
td = "/dev/shm/cvvdp_probe"
os.makedirs(td, exist_ok=True)
props = {"width": W, "height": H, "fps": fps, "bit_depth": 8, "color_space": "709", "chroma_ss": "420"}
names = [os.path.join(td, cv.create_yuv_fname(n, props)) for n in ("test", "ref")]
tst, ref = bench.make_clip(3, 0, F, H, W, "u8", dev)
for fn, clip in zip(names, (tst, ref)):
    with open(fn, "wb") as fh:
        for f in range(F):
            rgb = clip[0, :, f].float() / 255.0
            Y = 0.2126 * rgb[0] + 0.7152 * rgb[1] + 0.0722 * rgb[2]
            cb, cr = (rgb[2] - Y) / 1.8556, (rgb[0] - Y) / 1.5748
            planes = [(Y * 219 + 16).round().clamp(0, 255)]
            planes += [torch.nn.functional.avg_pool2d((c * 224 + 128)[None, None], 2)[0, 0].round().clamp(0, 255) for c in (cb, cr)]
            fh.write(torch.cat([p.reshape(-1) for p in planes]).to(torch.uint8).cpu().numpy().tobytes())
del tst, ref
vs = cv.video_source_yuv_file(names[0], names[1], display_photometry="standard_4k")
tr, rr = vs.test_vidr, vs.reference_vidr
info = m._plan(1, H, W, F, fps, 3, tr.native_dtype(), vs.dm_photometry, tr.native_yuv())
fpix = tr.frame_pixels


def clip_of(ptr):
    c = N.Clip()
    c.data = ptr
    c.stride[0], c.stride[2] = 0, fpix
    c.frame0, c.n_frames = 0, F
    return c


def timed(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


host_t = np.fromfile(names[0], dtype=np.uint8)
host_r = np.fromfile(names[1], dtype=np.uint8)
dt, dr = torch.from_numpy(host_t).to(dev), torch.from_numpy(host_r).to(dev)
Q = torch.zeros((1, 4, F, info.n_bands), device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
m._ctx.profile_enable(True)
m._ctx.profile_read()
print("resident in HBM      %.1f ms" % timed(lambda: m._ctx.process_device(clip_of(dt.data_ptr()), clip_of(dr.data_ptr()), 0, F, Q.data_ptr(), None, st)))
for k in m._ctx.profile_read():
    if k["kind"] == "temporal":
        print("   temporal kernel: %d launches, %.2f ms total" % (k["launches"], k["total_ms"]))
m._ctx.profile_enable(False)
del dt, dr
Qh = torch.zeros((1, 4, F, info.n_bands), pin_memory=True)
pt, pr = torch.from_numpy(host_t).pin_memory(), torch.from_numpy(host_r).pin_memory()
print("pinned host          %.1f ms" % timed(lambda: m._ctx.process_host(clip_of(pt.data_ptr()), clip_of(pr.data_ptr()), 0, F, Qh.data_ptr(), None)))
del pt, pr
print("pageable host        %.1f ms" % timed(lambda: m._ctx.process_host(clip_of(host_t.ctypes.data), clip_of(host_r.ctypes.data), 0, F, Qh.data_ptr(), None)))


def mapped():
    a, b = np.memmap(names[0], np.uint8, mode="r"), np.memmap(names[1], np.uint8, mode="r")
    m._ctx.process_host(clip_of(a.ctypes.data), clip_of(b.ctypes.data), 0, F, Qh.data_ptr(), None)


print("fresh file mapping   %.1f ms" % timed(mapped))
fdt, fdr = os.open(names[0], os.O_RDONLY), os.open(names[1], os.O_RDONLY)
null_clip = lambda: clip_of(None)
print("file descriptors     %.1f ms" % timed(lambda: m._ctx.process_files(null_clip(), null_clip(), fdt, fdr, 0, 0, 0, F, Qh.data_ptr(), None)))
print("predict_video_source %.1f ms" % timed(lambda: m.predict_video_source(cv.video_source_yuv_file(names[0], names[1], display_photometry="standard_4k"))))
import shutil
shutil.rmtree(td, ignore_errors=True)
