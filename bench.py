#!/usr/bin/env python
"""bench.py -- Mpixels/s of the ColorVideoVDP hot path on synthetic 4K@60fps video (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full prediction (front end, temporal filter, pyramid, CSF, masking, pooling, JOD) of the
workload: N=1 -> one 3840x2160x120-frame pair at 60 fps on standard_4k (BASELINE.json configs[2], the
configuration the metric is quoted on); N>1 -> a batch of N such pairs (configs[4] at N=8).  The flat
(item, frame) sequence is cut into one contiguous run per rank (distributed.work_shard): with N items on N
GPUs every rank owns one whole clip and reads no temporal halo; ONE NCCL all-reduce of Q_per_ch, identical
pooling on every rank.  Per-GPU work is constant => weak scaling.  (`--shard frames` keeps round 1's pure
frame sharding of every item, 2.07x halo at N=8, for comparison.)

One JSON line on stdout (rank 0):
  value         device-resident throughput (inputs in HBM when the timed region starts)
  e2e           the same through the public API -- N=1: literally `cvvdp.predict(test, ref, "BCFHW", fps)` with
                pinned host tensors; N>1: distributed.predict_sharded on pinned host clips -- H2D inside the timed region
  roofline      dominant kernel, CUDA events recorded around every launch of the timed steps
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, tools/stage_reference.sh) on the host cores on a bounded
                sample (first frames of the same clip); falls back to the numpy oracle port when it is not staged
  extra         fp32-input run of the same workload (SURVEY 8d canonical dtype), BASELINE configs[1] (1080p) and
                configs[3] (4K HDR PQ + raw heat map), predict() on pageable numpy arrays, reference on cuda
`--impl reference` times the CPU arm alone (reference if staged, else the oracle port).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

DISPLAY = "standard_4k"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--fps", type=float, default=60.0)
    ap.add_argument("--dtype", default="u8", choices=["u8", "f32", "f16", "u16"])
    ap.add_argument("--shard", default="runs", choices=["runs", "frames"],
                    help="N>1: contiguous (item, frame) runs (whole items when batch >= ranks) or frame-shard every item")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--e2e-exchange", action="store_true",
                    help="--shard frames only: hold only the owned frames on each host, fetch the history over NVLink")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="abort the process after this many seconds (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU sample (0 = auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic workload (same construction as tests/synth.py, generated with torch so that it can be
# produced directly in HBM for the big configurations)
# ---------------------------------------------------------------------------------------------------
def make_clip(seed, F_lo, F_hi, H, W, dtype, device, hdr=False):
    """(test, ref) [1,3,F_hi-F_lo,H,W] holding clip frames [F_lo, F_hi) of the clip with seed `seed`.
    hdr: 10-bit PQ code values 64..800 stored as uint16 (code * 64), dtype must be "u16"."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    base = torch.rand((1, 3, max(H // 8, 2), max(W // 8, 2)), generator=g, device=device)
    base = torch.nn.functional.avg_pool2d(base, 3, 1, 1)
    base = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False)[0]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :]
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None]
    n = F_hi - F_lo
    tdt = {"u8": torch.uint8, "f32": torch.float32, "f16": torch.float16, "u16": torch.int16}[dtype]
    tst = torch.empty((1, 3, n, H, W), dtype=tdt, device=device)
    ref = torch.empty((1, 3, n, H, W), dtype=tdt, device=device)
    for i, f in enumerate(range(F_lo, F_hi)):
        gf = torch.Generator(device=device)
        gf.manual_seed(seed * 100003 + f)
        for c in range(3):
            grating = 0.5 + 0.5 * torch.sin((xx + 0.5 * yy - 2.0 * f) * (2 * math.pi / (16.0 + 8 * c)))
            img = 0.55 * base[c] + 0.35 * grating + 0.02 * torch.randn((H, W), generator=gf, device=device) + 0.05
            r = (16.0 + img * 219.0).clamp(16, 235).round()
            t = (r + 4.0 * torch.randn((H, W), generator=gf, device=device) + (3.0 if c == 0 else 0.0)).clamp(0, 255).round()
            for dst, v in ((ref, r), (tst, t)):
                if dtype == "u8":
                    dst[0, c, i] = v.to(torch.uint8)
                elif dtype == "u16" and hdr:
                    code = (64.0 + (v - 16.0) / 219.0 * (800.0 - 64.0)).round().clamp(0, 1023)
                    dst[0, c, i] = (code * 64).to(torch.int32).to(torch.int16)  # bit pattern of uint16
                elif dtype == "u16":
                    dst[0, c, i] = (v * 257).to(torch.int32).to(torch.int16)
                else:
                    dst[0, c, i] = (v / 255).to(tdt)
    return tst, ref


def elem_size(dtype):
    return {"u8": 1, "f32": 4, "f16": 2, "u16": 2}[dtype]


def pinned_like(t):
    return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)


# ---------------------------------------------------------------------------------------------------
# CPU arm.  Preferred: the unmodified reference from baseline/_ref on all host threads.  Fallback: the
# numpy oracle port with frames spread over host threads.
# ---------------------------------------------------------------------------------------------------
def reference_staged():
    from oracle import ref_loader as RL
    return RL.staged()


class ReferenceCPU:
    """pycvvdp.cvvdp(device='cpu') on torch's intra-op thread pool (= host cores)."""

    def __init__(self, display):
        from oracle import ref_loader as RL
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.metric = RL.reference_metric(display, "cpu")
        tiny = torch.zeros((1, 3, 2, 32, 32), dtype=torch.uint8)
        with torch.no_grad():
            self.metric.predict(tiny, tiny, dim_order="BCFHW", frames_per_second=30)  # lazy initialisation, untimed

    def predict(self, tst, ref, fps):
        """(seconds, JOD, Q_per_ch) for BCFHW torch CPU tensors."""
        t0 = time.perf_counter()
        with torch.no_grad():
            jod, stats = self.metric.predict(tst, ref, dim_order="BCFHW", frames_per_second=fps)
        return time.perf_counter() - t0, float(jod), np.asarray(stats["Q_per_ch"])


def oracle_sample(tst, ref, fps, first_frame, f_lo, f_hi, n_frames_total, threads):
    """Oracle-port Q_per_ch for frames [f_lo, f_hi) given numpy windows starting at clip frame `first_frame`.
    Returns (seconds, Q)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cvvdp_oracle as O
    dm, P = O.Display(DISPLAY), O.Params()
    H, W = tst.shape[-2:]
    rho, _ = O.band_frequencies(W, H, dm.ppd)
    filt = O.temporal_filters(fps, P)
    fl = len(filt[0])
    t0 = time.perf_counter()
    need = sorted({max(t, 0) for f in range(f_lo, f_hi) for t in range(f - fl + 1, f + 1)})
    with ThreadPoolExecutor(max_workers=threads) as ex:
        dk = list(ex.map(lambda t: (O.frontend(tst[:, :, t - first_frame], dm), O.frontend(ref[:, :, t - first_frame], dm)), need))
        cache = dict(zip(need, dk))

        def one(f):
            R = O.temporal_channels(lambda t: cache[t][0], lambda t: cache[t][1], f, filt, "replicate", n_frames_total)
            return O.process_frame(R, rho, P, False)[0]

        Q = list(ex.map(one, range(f_lo, f_hi)))
    return time.perf_counter() - t0, np.stack(Q, axis=2)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    H, W = args.height, args.width
    total_steps = args.warmup + args.steps
    if reference_staged():
        ref_cpu = ReferenceCPU(DISPLAY)
        # calibrate on 2 frames (counts as the first warm-up step), then size the sample so that the run ends
        # within a few minutes: each step = predict() of the first S frames of the clip as an S-frame video
        S = 2
        tst, ref = make_clip(3, 0, 16 if args.cpu_frames == 0 else max(args.cpu_frames, 2), H, W, args.dtype, torch.device("cpu"))
        dt, _, _ = ref_cpu.predict(tst[:, :, :S], ref[:, :, :S], args.fps)
        if args.cpu_frames > 0:
            S = max(args.cpu_frames, 2)
        else:
            S = int(max(2, min(16, (240.0 / max(total_steps, 1)) / (dt / 2))))
        times = []
        for i in range(total_steps):
            if i == 0 and S == 2:
                t = dt
            else:
                t, _, _ = ref_cpu.predict(tst[:, :, :S], ref[:, :, :S], args.fps)
            if i >= args.warmup:
                times.append(t)
        kind, threads = "reference", ref_cpu.threads
        sample = (f"pycvvdp.cvvdp(device='cpu').predict on the first {S} frames of one {W}x{H} pair as a {S}-frame "
                  f"{args.fps:g} fps video, torch intra-op threads = {threads} of {cores} host cores")
    else:
        from oracle import cvvdp_oracle as O
        S = args.cpu_frames if args.cpu_frames > 0 else int(max(2, min(8, cores, args.frames)))
        fl = len(O.temporal_filters(args.fps, O.Params())[0])
        f_lo = max(min(fl - 1, args.frames - S), 0)
        w_lo = max(f_lo - (fl - 1), 0)
        tst, ref = make_clip(3, w_lo, f_lo + S, H, W, args.dtype, torch.device("cpu"))
        tst, ref = tst.numpy(), ref.numpy()
        if args.dtype == "u16":
            tst, ref = tst.view(np.uint16), ref.view(np.uint16)
        threads = min(cores, S)
        times = []
        for i in range(total_steps):
            dt, _ = oracle_sample(tst, ref, args.fps, w_lo, f_lo, f_lo + S, args.frames, threads)
            if i >= args.warmup:
                times.append(dt)
        kind = "port"
        sample = (f"frames [{f_lo},{f_lo + S}) of one {W}x{H} pair incl. the front end of their {fl - 1} history "
                  f"frames, numpy oracle port (reference not staged), {threads} threads of {cores} host cores")
    ms = 1e3 * float(np.mean(times))
    mpix = S * H * W / 1e6 / (ms / 1e3)
    line = {"impl": "reference", "metric": "Mpixels/s", "value": round(mpix, 4), "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": round(mpix, 4), "unit": "Mpix/s", "cores": threads, "host_cores": cores,
                             "kind": kind, "sample": sample},
            "e2e": {"value": round(mpix, 4), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n):
    name = (f"{n}x " if n > 1 else "") + f"{args.width}x{args.height}x{args.frames}f test/ref pair @{args.fps:g}fps, {DISPLAY}"
    if n > 1 and args.shard == "runs":
        name += f", batch of {n} over {n} GPUs (one whole clip per rank), one all-reduce of Q_per_ch"
    elif n > 1:
        name += f", batch of {n} frame-sharded over {n} GPUs, one all-reduce of Q_per_ch"
    return {"workload": name, "batch": n, "frames": args.frames, "height": args.height, "width": args.width,
            "fps": args.fps, "display": DISPLAY, "input_dtype": args.dtype,
            "parallelism": (f"(item, frame) runs x{n}" if args.shard == "runs" else f"frame-shard x{n}") if n > 1 else "single GPU",
            "l2": "inputs and every level-0 intermediate are far larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        """Start polling (nvidia-smi needs a few hundred ms before its first line: call this before the warm-up)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t_begin = time.monotonic()

    def end(self):
        self.t_end = time.monotonic()

    def stop(self, load_fn=None):
        """Samples taken between begin() and end().  A timed region shorter than the polling period can miss them all:
        then `load_fn` (the same step, untimed) keeps the GPU under the same load until two samples have arrived."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = lambda: [r for (t, r) in self.rows if self.t_begin is not None and self.t_begin <= t <= (self.t_end or t)]  # noqa: E731
        note = "timed region"
        if len(inside()) < 2 and load_fn is not None:
            note = "timed region + untimed repeats of the same step"
            t_stop = time.monotonic() + 4.0
            self.t_end = None
            while len(inside()) < 2 and time.monotonic() < t_stop:
                load_fn()
            self.t_end = time.monotonic()
        rows = inside()
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[7]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None, "sampled_during": note}


def start_watchdog(seconds):
    """A hung collective must not take the whole GPU box with it: leave after `seconds`, loudly."""
    if seconds <= 0:
        return

    def _bark():
        sys.stderr.write(f"bench.py watchdog: no result after {seconds:.0f} s, aborting\n")
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, _bark)
    t.daemon = True
    t.start()


def kernel_breakdown(prof, steps):
    kernels = {}
    for p in prof:
        key = p["kind"] + (f"_l{p['level']}" if p["kind"] in ("band", "reduce") else "")
        k = kernels.setdefault(key, {"launches": 0, "ms": 0.0, "bytes": 0.0})
        k["launches"] += p["launches"]; k["ms"] += p["total_ms"]; k["bytes"] += p["algo_bytes"]
    tot_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    breakdown = {name: {"launches": k["launches"], "ms_per_step": round(k["ms"] / steps, 4),
                        "share": round(k["ms"] / tot_ms, 4),
                        "algo_gbs": round(k["bytes"] / 1e9 / (k["ms"] / 1e3), 1) if k["ms"] > 0 else None}
                 for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}
    return kernels, breakdown


def time_steps(fn, steps, warmup, dev):
    """Mean ms per call of fn over `steps` calls (CUDA events on the current stream, sync on both sides)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, out


# ---------------------------------------------------------------------------------------------------
# extra lines (rank 0, N = 1): the other BASELINE configurations and input types
# ---------------------------------------------------------------------------------------------------
def yuv_file_run(cv, dev, args, H, W, F, fps):
    import shutil
    import tempfile
    tst, ref = make_clip(3, 0, F, H, W, "u8", dev)
    root = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 8 * F * H * W else None
    td = tempfile.mkdtemp(prefix="cvvdp_b200_bench_", dir=root)
    try:
        props = {"width": W, "height": H, "fps": fps, "bit_depth": 8, "color_space": "709", "chroma_ss": "420"}
        names = [os.path.join(td, cv.create_yuv_fname(n, props)) for n in ("test", "ref")]
        for fn, clip in zip(names, (tst, ref)):
            with open(fn, "wb") as fh:
                for f in range(F):  # BT.709 forward matrix, limited range, 2x2 chroma average
                    rgb = clip[0, :, f].float() / 255.0
                    Y = 0.2126 * rgb[0] + 0.7152 * rgb[1] + 0.0722 * rgb[2]
                    cb, cr = (rgb[2] - Y) / 1.8556, (rgb[0] - Y) / 1.5748
                    planes = [(Y * 219 + 16).round().clamp(0, 255)]
                    planes += [torch.nn.functional.avg_pool2d((c * 224 + 128)[None, None], 2)[0, 0].round().clamp(0, 255) for c in (cb, cr)]
                    fh.write(torch.cat([p.reshape(-1) for p in planes]).to(torch.uint8).cpu().numpy().tobytes())
        del tst, ref
        m = cv.cvvdp(display_name=DISPLAY, device=dev)

        def step():
            vs = cv.video_source_yuv_file(names[0], names[1], display_photometry=DISPLAY)
            return float(m.predict_video_source(vs)[0])

        ms, jod = time_steps(step, 2, 1, dev)
        return {"workload": f"{W}x{H}x{F}f @{fps:g}fps, {DISPLAY}, yuv420p 8-bit .yuv pair in {'/dev/shm' if root else 'tmp'}",
                "e2e": {"value": round(F * H * W / 1e6 / (ms / 1e3), 2), "unit": "Mpix/s", "ms_per_step": round(ms, 3),
                        "h2d_bytes_per_step": int(2 * F * H * W * 3 // 2)}, "jod": round(jod, 5)}
    finally:
        shutil.rmtree(td, ignore_errors=True)


def extra_runs(args, dev, peak, ref_sample):
    import colorvideovdp_b200 as cv
    out = {}
    H, W, F, fps = args.height, args.width, args.frames, args.fps

    def run(name, display, h, w, f, rate, dtype, heatmap=None, hdr=False, e2e=True, steps=3):
        m = cv.cvvdp(display_name=display, device=dev, heatmap=heatmap)
        tst, ref = make_clip(3, 0, f, h, w, dtype, dev, hdr=hdr)
        pix = f * h * w

        def step():
            return m.predict(tst, ref, dim_order="BCFHW", frames_per_second=rate)[0]

        m._ctx.profile_enable(False)
        ms, jod = time_steps(step, steps, 2, dev)
        m._ctx.profile_enable(True)
        m._ctx.profile_read()
        step()
        kernels, _ = kernel_breakdown(m._ctx.profile_read(), 1)
        m._ctx.profile_enable(False)
        ent = {"workload": f"{w}x{h}x{f}f @{rate:g}fps, {display}, input {dtype}" + (", raw heat map" if heatmap else ""),
               "value": round(pix / 1e6 / (ms / 1e3), 2), "unit": "Mpix/s", "ms_per_step": round(ms, 3),
               "jod": round(float(jod), 5)}
        if "temporal" in kernels and kernels["temporal"]["ms"] > 0:
            k = kernels["temporal"]
            ent["temporal_kernel"] = {"ms": round(k["ms"], 3), "algo_gbs": round(k["bytes"] / 1e9 / (k["ms"] / 1e3), 1),
                                      "frac_of_peak": round(k["bytes"] / 1e9 / (k["ms"] / 1e3) / peak, 4)}
        if e2e:
            th, rh = pinned_like(tst), pinned_like(ref)

            def step_h():  # returns sizes only: holding on to stats would keep a 1 GB pinned heat map alive and make
                jod_h, stats = m.predict(th, rh, dim_order="BCFHW", frames_per_second=rate)  # the next call allocate afresh
                return float(jod_h), stats["Q_per_ch"].nbytes + 4 + (stats["heatmap"].numel() * 2 if heatmap else 0)

            ms_h, (jod_h, d2h) = time_steps(step_h, 2, 2, dev)
            ent["e2e"] = {"value": round(pix / 1e6 / (ms_h / 1e3), 2), "unit": "Mpix/s", "ms_per_step": round(ms_h, 3),
                          "h2d_bytes_per_step": int(2 * th.numel() * th.element_size()), "d2h_bytes_per_step": int(d2h)}
            assert abs(jod_h - float(jod)) < 1e-4, (jod_h, float(jod))
            del th, rh
        out[name] = ent
        del tst, ref, m
        torch.cuda.empty_cache()

    # canonical reporting dtype fp32 [B,C,F,H,W] (24 algorithmic bytes per pixel-frame)
    run("config3_f32_input", DISPLAY, H, W, F, fps, "f32", e2e=False)
    # BASELINE configs[1]: 1920x1080, 60 frames, 30 fps, standard_fhd
    run("config2_1080p", "standard_fhd", 1080, 1920, 60, 30.0, "u8")
    # BASELINE configs[3]: 3840x2160 HDR (PQ), 60 frames, standard_hdr_pq, with the raw heat map (fp16, 1 GB D2H)
    run("config4_hdr_pq_heatmap", "standard_hdr_pq", 2160, 3840, 60, 60.0, "u16", heatmap="raw", hdr=True)

    # BASELINE configs[0]: one 256x256 image pair (launch- and host-bound: 18 launches per call), and a 1080p image
    for name, h, w in (("config1_image_256", 256, 256), ("image_1080p", 1080, 1920)):
        m = cv.cvvdp(display_name="standard_fhd", device=dev)
        ti, ri = make_clip(3, 0, 1, h, w, "u8", dev)
        ms_i, jod_i = time_steps(lambda: float(m.predict(ti, ri, dim_order="BCFHW")[0]), 200, 20, dev)
        out[name] = {"workload": f"{w}x{h} image pair, standard_fhd, u8, predict() on device tensors incl. the JOD read-back",
                     "ms_per_call": round(ms_i, 4), "value": round(h * w / 1e6 / (ms_i / 1e3), 2), "unit": "Mpix/s", "jod": round(jod_i, 5)}
        del m, ti, ri

    # predict() on PAGEABLE numpy arrays (what a user who just loaded a clip passes)
    m = cv.cvvdp(display_name=DISPLAY, device=dev)
    tst, ref = make_clip(3, 0, F, H, W, args.dtype, dev)
    tn, rn = tst.cpu().numpy().copy(), ref.cpu().numpy().copy()
    del tst, ref

    def step_p():
        return float(m.predict(tn, rn, dim_order="BCFHW", frames_per_second=fps)[0])

    ms_p, jod_p = time_steps(step_p, 2, 1, dev)
    out["config3_predict_pageable_numpy"] = {"value": round(F * H * W / 1e6 / (ms_p / 1e3), 2), "unit": "Mpix/s",
                                             "ms_per_step": round(ms_p, 3), "jod": round(jod_p, 5),
                                             "h2d_bytes_per_step": int(tn.nbytes + rn.nbytes)}
    del m, tn, rn

    # raw planar YUV files: predict_video_source(video_source_yuv_file) on an 8-bit 4:2:0 pair of the
    # headline shape; the file mapping (page cache) is read by the library's upload threads, 1.5 bytes per pixel
    try:
        out["config3_yuv420p_files"] = yuv_file_run(cv, dev, args, H, W, F, fps)
    except Exception as e:  # informational only
        out["config3_yuv420p_files"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    # the reference itself on this GPU (informational; tests/test_gpu_reference.py holds the parity checks)
    if ref_sample is not None:
        try:
            from oracle import ref_loader as RL
            S = 8
            refm = RL.reference_metric(DISPLAY, dev)
            tst, ref = make_clip(3, 0, S, H, W, args.dtype, dev)
            with torch.no_grad():
                refm.predict(tst[:, :, :2], ref[:, :, :2], dim_order="BCFHW", frames_per_second=fps)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                jr, _ = refm.predict(tst, ref, dim_order="BCFHW", frames_per_second=fps)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
            m = cv.cvvdp(display_name=DISPLAY, device=dev)
            jo, _ = m.predict(tst, ref, dim_order="BCFHW", frames_per_second=fps)
            out["reference_cuda"] = {"value": round(S * H * W / 1e6 / dt, 2), "unit": "Mpix/s",
                                     "sample": f"pycvvdp.cvvdp(device='cuda').predict on the first {S} frames, TF32 off",
                                     "jod_reference": round(float(jr), 5), "jod_b200": round(float(jo), 5)}
            del refm, m, tst, ref
        except Exception as e:  # informational only
            out["reference_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()
    return out


def main():
    args = parse_args()
    start_watchdog(args.watchdog)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from colorvideovdp_b200 import distributed as D
    numa = D.bind_to_gpu_numa_node(local) if world > 1 else "single process"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n = world
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    import colorvideovdp_b200 as cv
    metric = cv.cvvdp(display_name=DISPLAY, device=dev)
    F, H, W, fps = args.frames, args.height, args.width, args.fps

    # ---- this rank's share of the batch of n clips (item b has seed 3 + 17 b) ----
    if args.shard == "runs":
        shard = D.work_shard(n, F, rank, n)
    else:
        lo, hi = D.frame_shard(F, rank, n)
        shard = [(b, lo, hi) for b in range(n)]
    pieces = []
    for item, f_lo, f_hi in shard:
        wlo, whi = D.needed_window(metric, F, fps, f_lo, f_hi)
        tst, ref = make_clip(3 + 17 * item, wlo, whi, H, W, args.dtype, dev)
        pieces.append((item, f_lo, f_hi, wlo, tst, ref))

    def step_device():
        jod, Q = D.predict_sharded(metric, pieces, n, F, fps)
        return jod

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    pix_per_step = n * F * H * W
    clocks = ClockSampler(local)
    metric._ctx.profile_enable(False)
    if rank == 0:
        clocks.start()
    # ---- device-resident arm ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = metric._ctx.launch_count()
    metric._ctx.profile_enable(True)
    metric._ctx.profile_read()
    clocks.begin()
    ms_dev, jod = timed(step_device, args.steps, 0)
    clocks.end()
    prof = metric._ctx.profile_read()
    metric._ctx.profile_enable(False)
    launches = metric._ctx.launch_count() - launches0

    def load_again():  # rank 0 only, after the timed region: no collective inside step_device
        step_device()
        torch.cuda.synchronize(dev)

    clk = clocks.stop(load_again if world == 1 else None) if rank == 0 else None
    value = pix_per_step / 1e6 / (ms_dev / 1e3)

    # ---- end-to-end arm: pinned host clips through the public API ----
    e2e = None
    do_e2e = not args.no_e2e
    own_bytes = sum(2 * p[4].numel() * p[4].element_size() for p in pieces)
    if do_e2e:  # never pin more than half of the host's available memory (all ranks together)
        try:
            avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:
            avail = 1 << 62
        flag = torch.tensor([1 if world * own_bytes < 0.5 * avail else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        do_e2e = bool(flag.item())
    if do_e2e:
        exchange = args.e2e_exchange and world > 1 and args.shard == "frames"
        if world == 1:
            # N = 1: literally the public call a user makes -- cvvdp.predict() on (pinned) host tensors; its stats
            # dictionary brings Q_per_ch back to the host, float(jod) reads the score
            tst_h, ref_h = pinned_like(pieces[0][4]), pinned_like(pieces[0][5])
            h2d = int(2 * tst_h.numel() * tst_h.element_size())

            def step_host():
                jod_h, stats = metric.predict(tst_h, ref_h, dim_order="BCFHW", frames_per_second=fps)
                return torch.tensor([float(jod_h)])

            how = "cvvdp.predict(test, ref, 'BCFHW', fps) on pinned host tensors (streamed upload overlapped with compute)"
        elif exchange:
            # round 1's frame sharding with every rank holding only the frames it owns; history over NVLink
            lo, hi = D.frame_shard(F, rank, n)
            wlo = pieces[0][3]
            tst_h = pinned_like(torch.cat([p[4][:, :, lo - wlo:hi - wlo] for p in pieces], 0))
            ref_h = pinned_like(torch.cat([p[5][:, :, lo - wlo:hi - wlo] for p in pieces], 0))
            h2d = int(2 * tst_h.numel() * tst_h.element_size())

            def step_host():
                jod_h, Qd = D.predict_frame_sharded_exchange(metric, tst_h, ref_h, F, fps)
                return jod_h.cpu()

            how = "owned frames uploaded once from pinned host memory; history frames by NCCL send/recv over NVLink"
        else:
            host_pieces = [(p[0], p[1], p[2], p[3], pinned_like(p[4]), pinned_like(p[5])) for p in pieces]
            h2d = int(sum(2 * p[4].numel() * p[4].element_size() for p in host_pieces))

            def step_host():
                jod_h, Qd = D.predict_sharded(metric, host_pieces, n, F, fps)
                return jod_h.cpu()  # D2H of the result

            how = ("distributed.predict_sharded on pinned host clips (each rank uploads exactly the frames of its run"
                   + ("" if args.shard == "runs" else " plus their temporal halo") + f"); host memory {numa}")
        ms_e2e, jod_h = timed(step_host, max(2, min(args.steps, 3)), 1)
        assert torch.allclose(jod_h.flatten().float(), jod.flatten().cpu().float(), atol=1e-5), "host and device arms disagree"
        h2d_t = torch.tensor([h2d], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)  # the real bytes of every rank
        info = metric._info
        q_bytes = n * info.n_channels * F * info.n_bands * 4
        e2e = {"value": round(pix_per_step / 1e6 / (ms_e2e / 1e3), 2), "unit": "Mpix/s",
               "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": int(h2d_t.item()),
               "d2h_bytes_per_step": int(q_bytes + 4 * n), "input": how}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_file):
        peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kernels, breakdown = kernel_breakdown(prof, args.steps)
    dom_name = max(kernels, key=lambda name: kernels[name]["ms"]) if kernels else None
    roof = None
    if dom_name:
        k = kernels[dom_name]
        achieved = k["bytes"] / 1e9 / (k["ms"] / 1e3)
        # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture of this same
        # command (profiles/ncu_traffic.json), rescaled if this run's launch covers a different frame count
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tfile):
            ent = json.load(open(tfile)).get(dom_name)
            if ent:
                traffic = round(ent["dram_bytes_per_launch"] * (k["bytes"] / k["launches"]) / ent["algo_bytes_per_launch"], 0)
        roof = {"kernel": dom_name, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": round(k["ms"] / k["launches"], 4),
                "algo_bytes_per_launch": k["bytes"] / k["launches"],
                "note": "fp32 FMA / MUFU / shared-memory bound kernel: see DESIGN.md for why it sits below the HBM roofline"}
        whole = args.steps * pix_per_step / n * 2 * 3 * elem_size(args.dtype)  # rank-0 share of B_alg
        roof["whole_path_algo_gbs"] = round(whole / 1e9 / (args.steps * ms_dev / 1e3), 1)

    # ---- CPU baseline on a bounded sample + parity of this engine against it on that sample ----
    cpu, ref_sample = None, None
    if not args.no_cpu_baseline and n == 1:
        cores = os.cpu_count() or 1
        tst0, ref0, wlo0 = pieces[0][4], pieces[0][5], pieces[0][3]
        if reference_staged():
            S = args.cpu_frames if args.cpu_frames > 0 else 2
            ref_cpu = ReferenceCPU(DISPLAY)
            ts, rs = tst0[:, :, :S].cpu(), ref0[:, :, :S].cpu()
            dt, jod_r, Q_r = ref_cpu.predict(ts, rs, fps)
            jod_g, stats_g = metric.predict(ts, rs, dim_order="BCFHW", frames_per_second=fps)
            gate = float(np.max(np.abs(stats_g["Q_per_ch"] - Q_r) / (1e-3 * np.abs(Q_r) + 1e-5)))
            cpu = {"value": round(S * H * W / 1e6 / dt, 4), "unit": "Mpix/s", "cores": ref_cpu.threads, "host_cores": cores,
                   "kind": "reference",
                   "sample": f"pycvvdp.cvvdp(device='cpu').predict on the first {S} frames of the same clip as a {S}-frame "
                             f"video, torch threads = {ref_cpu.threads}, {dt:.1f} s",
                   "parity_on_sample": {"jod_reference": round(jod_r, 6), "jod_b200": round(float(jod_g), 6),
                                        "q_err_over_gate": round(gate, 4)}}
            ref_sample = True
            del ref_cpu
        else:
            S = args.cpu_frames if args.cpu_frames > 0 else int(max(2, min(8, cores, F)))
            info = metric._info
            fl = info.filter_len
            f_lo = max(min(fl - 1, F - S), 0)
            w0 = max(f_lo - (fl - 1), 0)
            t_np = tst0[:1, :, w0 - wlo0:f_lo + S - wlo0].cpu().numpy()
            r_np = ref0[:1, :, w0 - wlo0:f_lo + S - wlo0].cpu().numpy()
            if args.dtype == "u16":
                t_np, r_np = t_np.view(np.uint16), r_np.view(np.uint16)
            threads = min(cores, S)
            dt, Qo = oracle_sample(t_np, r_np, fps, w0, f_lo, f_lo + S, F, threads)
            Qg, _ = metric.q_per_ch_from_tensors(tst0[:1], ref0[:1], F, fps, (f_lo, f_lo + S), wlo0)
            Qg = Qg[:, :, f_lo:f_lo + S].cpu().numpy()
            gate = float(np.max(np.abs(Qg - Qo) / (1e-3 * np.abs(Qo) + 1e-5)))
            cpu = {"value": round(S * H * W / 1e6 / dt, 4), "unit": "Mpix/s", "cores": threads, "host_cores": cores,
                   "kind": "port",
                   "sample": f"frames [{f_lo},{f_lo + S}) of the same clip (incl. front end of {fl - 1} history frames), "
                             f"numpy oracle port (reference not staged), {threads} threads, {dt:.1f} s",
                   "parity_on_sample": {"q_err_over_gate": round(gate, 4)}}

    extra = None
    if n == 1 and not args.no_extras:
        del pieces, metric
        torch.cuda.empty_cache()
        extra = extra_runs(args, dev, peak, ref_sample)

    line = {"metric": "Mpixels/s", "value": round(value, 2), "unit": "Mpix/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "jod": [round(float(v), 5) for v in jod.flatten().cpu()], "clocks": clk, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": breakdown, "extra": extra}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
