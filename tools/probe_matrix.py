"""Throughput across the input space (device-resident tensors): looks for paths that fall off the fast kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import colorvideovdp_b200 as cv  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def run(name, *a, **k):
    try:
        run_case(name, *a, **k)
    except Exception as e:  # keep going: the point is to find the odd one out
        print(f"{name:44s} FAILED: {type(e).__name__}: {str(e)[:160]}")
        torch.cuda.empty_cache()


def run_case(name, F, H, W, fps, display="standard_fhd", dtype="u8", heatmap=None, padding="replicate", B=1, gray=False, dim_order="BCFHW"):
    tst, ref = synth.make_pair_u8(5, min(F, 8), H, W)
    reps = -(-F // tst.shape[2])
    td = torch.from_numpy(tst).to(dev).repeat(B, 1, reps, 1, 1)[:, :, :F].contiguous()
    rd = torch.from_numpy(ref).to(dev).repeat(B, 1, reps, 1, 1)[:, :, :F].contiguous()
    if gray:
        td, rd = td[:, :1].contiguous(), rd[:, :1].contiguous()
    if dtype == "f32":
        td, rd = td.float() / 255, rd.float() / 255
    elif dtype == "f16":
        td, rd = (td.float() / 255).half(), (rd.float() / 255).half()
    elif dtype == "u16":
        td, rd = (td.to(torch.int32) * 257).to(torch.int16), (rd.to(torch.int32) * 257).to(torch.int16)
    if dim_order != "BCFHW":
        perm = ["BCFHW".index(c) for c in dim_order]
        td, rd = td.permute(perm).contiguous(), rd.permute(perm).contiguous()
    m = cv.cvvdp(display_name=display, device=dev, heatmap=heatmap, temp_padding=padding)
    ms = timed(lambda: m.predict(td, rd, dim_order=dim_order, frames_per_second=fps))
    m._ctx.profile_enable(True)
    m._ctx.profile_read()
    m.predict(td, rd, dim_order=dim_order, frames_per_second=fps)
    ks = sorted(m._ctx.profile_read(), key=lambda k: -k["total_ms"])[:2]
    m._ctx.profile_enable(False)
    top = ", ".join(f"{k['kind']}_l{k['level']} {k['total_ms']:.2f}" for k in ks)
    print(f"{name:44s} {B * F * H * W / 1e6 / (ms / 1e3) / 1e3:7.2f} Gpix/s  {ms:8.3f} ms   top: {top}")
    del td, rd, m
    torch.cuda.empty_cache()


run("1080p 60f @30 u8 (config 2)", 60, 1080, 1920, 30)
run("1080p 60f @24 u8 (7 taps)", 60, 1080, 1920, 24)
run("1080p 60f @120 u8 (31 taps: generic)", 60, 1080, 1920, 120)
run("1080p 60f @30 u8 symmetric padding", 60, 1080, 1920, 30, padding="symmetric")
run("1000x600 60f @30 u8 (width not /64)", 60, 600, 1000, 30)
run("1080p 60f @30 f16", 60, 1080, 1920, 30, dtype="f16")
run("1080p 60f @30 u16 PQ", 60, 1080, 1920, 30, dtype="u16", display="standard_hdr_pq")
run("1080p 60f @30 f32 HLG (generic body)", 60, 1080, 1920, 30, dtype="f32", display="standard_hdr_hlg")
run("1080p 60f @30 f32 linear HDR", 60, 1080, 1920, 30, dtype="f32", display="standard_hdr_linear")
run("1080p 60f @30 gray u8", 60, 1080, 1920, 30, gray=True)
run("1080p 60f @30 gray f32", 60, 1080, 1920, 30, gray=True, dtype="f32")
run("1080p 60f @30 u8 FHWC layout (permuted)", 60, 1080, 1920, 30, dim_order="BFHWC")
run("1080p 60f @30 u8 raw heat map", 60, 1080, 1920, 30, heatmap="raw")
run("1080p 60f @30 u8 supra-threshold heat map", 60, 1080, 1920, 30, heatmap="supra-threshold")
run("1080p batch of 4 x 30f @30 u8", 30, 1080, 1920, 30, B=4)
run("1080p batch of 32 images u8", 1, 1080, 1920, 0, B=32)
run("4K image u8", 1, 2160, 3840, 0, display="standard_4k")
run("4K image u8 raw heat map", 1, 2160, 3840, 0, display="standard_4k", heatmap="raw")
run("480p 300f @30 u8", 300, 480, 832, 30)
