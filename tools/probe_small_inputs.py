"""Latency of predict() on small inputs (BASELINE configs[0]: a 256x256 image) -- launch- and host-bound territory."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import colorvideovdp_b200 as cv  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")
m = cv.cvvdp(display_name="standard_fhd", device=dev)


def timed(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for (F, H, W, fps) in [(1, 256, 256, 0), (1, 1080, 1920, 0), (8, 256, 256, 30), (30, 540, 960, 30)]:
    tst, ref = synth.make_pair_u8(5, F, H, W)
    td, rd = torch.from_numpy(tst).to(dev), torch.from_numpy(ref).to(dev)
    n0 = m._ctx.launch_count()
    ms_dev = timed(lambda: m.predict(td, rd, frames_per_second=fps))
    launches = (m._ctx.launch_count() - n0) / 220
    ms_sync = timed(lambda: float(m.predict(td, rd, frames_per_second=fps)[0]))
    ms_host = timed(lambda: float(m.predict(tst, ref, frames_per_second=fps)[0]), n=100, warm=10)
    Q = torch.zeros((1, 4 if F > 1 else 3, F, 8), device=dev)
    print(f"{F}x{H}x{W}: predict(device tensors) {ms_dev:.3f} ms/call ({launches:.0f} launches), with .item() {ms_sync:.3f} ms, "
          f"predict(numpy) {ms_host:.3f} ms")

# per-kernel breakdown of the 1080p image
tst, ref = synth.make_pair_u8(5, 1, 1080, 1920)
td, rd = torch.from_numpy(tst).to(dev), torch.from_numpy(ref).to(dev)
m.predict(td, rd)
m._ctx.profile_enable(True)
m._ctx.profile_read()
for _ in range(10):
    m.predict(td, rd)
for k in sorted(m._ctx.profile_read(), key=lambda k: -k["total_ms"])[:8]:
    print(f"   {k['kind']}_l{k['level']}: {k['total_ms'] / 10:.3f} ms per call ({k['launches'] // 10} launches)")
m._ctx.profile_enable(False)
