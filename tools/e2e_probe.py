"""Break down the host-buffer path: upload-only pipeline vs full pipeline vs device-resident compute."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import colorvideovdp_b200 as cv
F, H, W, fps = 120, 2160, 3840, 60.0
dev = torch.device("cuda:0")
tst, ref = bench.make_clip(3, 0, F, H, W, "u8", dev)
th = torch.empty(tst.shape, dtype=tst.dtype, pin_memory=True).copy_(tst)
rh = torch.empty(ref.shape, dtype=ref.dtype, pin_memory=True).copy_(ref)
m = cv.cvvdp(display_name="standard_4k", device=dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("device-resident ms", t(lambda: m.q_per_ch_from_tensors(tst, ref, F, fps)))
print("host path ms      ", t(lambda: m.q_per_ch_from_tensors(th, rh, F, fps)))
print("plain H2D of both clips ms", t(lambda: (th.to(dev, non_blocking=True), rh.to(dev, non_blocking=True))))
