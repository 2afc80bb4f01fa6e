#!/usr/bin/env python
"""bench.py -- Mpixels/s of the ColorVideoVDP hot path on synthetic 4K@60fps video (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full prediction (front end, temporal filter, pyramid, CSF, masking, pooling, JOD) of the
workload: N=1 -> one 3840x2160x120-frame pair at 60 fps on standard_4k (BASELINE.json configs[2], the
configuration the metric is quoted on); N>1 -> a batch of N such pairs frame-sharded over the N GPUs
with one NCCL all-reduce of Q_per_ch (configs[4] at N=8): per-GPU work is constant => weak scaling.

One JSON line on stdout (rank 0): `value` = device-resident throughput, `e2e` = the same through
cvvdp's public tensor API with pinned HOST clips (H2D inside the timed region), `roofline` for the
dominant kernel (band level 0) from CUDA events recorded around every launch of the timed steps,
`cpu_baseline` = the numpy oracle timed on a bounded sample on the host cores.
`--impl reference` times the CPU arm alone (oracle port; the reference is Python and cannot travel).
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-exchange", action="store_true",
                    help="N>1: hold only the owned frames on each host and fetch the temporal history over NVLink")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="abort the process after this many seconds (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU sample (0 = auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic workload (same construction as tests/synth.py, generated with torch so that it can be
# produced directly in HBM for the big configurations)
# ---------------------------------------------------------------------------------------------------
def make_clip(seed, F_lo, F_hi, H, W, dtype, device):
    """(test, ref) [1,3,F_hi-F_lo,H,W] holding clip frames [F_lo, F_hi) of the clip with seed `seed`."""
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
                elif dtype == "u16":
                    dst[0, c, i] = (v * 257).to(torch.int32).to(torch.int16)  # bit pattern of uint16
                else:
                    dst[0, c, i] = (v / 255).to(tdt)
    return tst, ref


def elem_size(dtype):
    return {"u8": 1, "f32": 4, "f16": 2, "u16": 2}[dtype]


# ---------------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle on a bounded sample, frames spread over host threads
# ---------------------------------------------------------------------------------------------------
def oracle_sample(tst, ref, fps, first_frame, f_lo, f_hi, n_frames_total, threads):
    """Oracle Q_per_ch for frames [f_lo, f_hi) given numpy windows starting at clip frame `first_frame`.
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


def cpu_sample_frames(args, cores):
    if args.cpu_frames > 0:
        return args.cpu_frames
    # ~7 s of single-thread numpy per 4K frame; keep a step to a handful of seconds per thread
    return int(max(2, min(8, cores, args.frames)))  # one frame per thread, ~2.5 GB of numpy temporaries each


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    S = cpu_sample_frames(args, cores)
    from oracle import cvvdp_oracle as O
    fl = len(O.temporal_filters(args.fps, O.Params())[0])
    f_lo = min(fl - 1, args.frames - S)
    f_lo = max(f_lo, 0)
    w_lo = max(f_lo - (fl - 1), 0)
    tst, ref = make_clip(3, w_lo, f_lo + S, args.height, args.width, args.dtype, torch.device("cpu"))
    tst, ref = tst.numpy(), ref.numpy()
    if args.dtype == "u16":
        tst, ref = tst.view(np.uint16), ref.view(np.uint16)
    threads = min(cores, S)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = oracle_sample(tst, ref, args.fps, w_lo, f_lo, f_lo + S, args.frames, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    mpix = S * args.height * args.width / 1e6 / (ms / 1e3)
    sample = (f"frames [{f_lo},{f_lo + S}) of one {args.width}x{args.height} pair incl. the front end of their "
              f"{fl - 1} history frames, numpy oracle port, {threads} threads")
    line = {"impl": "reference", "metric": "Mpixels/s", "value": round(mpix, 4), "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": round(mpix, 4), "unit": "Mpix/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(mpix, 4), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n):
    name = (f"{n}x " if n > 1 else "") + f"{args.width}x{args.height}x{args.frames}f test/ref pair @{args.fps:g}fps, {DISPLAY}"
    if n > 1:
        name += f", batch of {n} frame-sharded over {n} GPUs, one all-reduce of Q_per_ch"
    return {"workload": name, "batch": n, "frames": args.frames, "height": args.height, "width": args.width,
            "fps": args.fps, "display": DISPLAY, "input_dtype": args.dtype,
            "parallelism": f"frame-shard x{n}" if n > 1 else "single GPU",
            "l2": "inputs and every level-0 intermediate are far larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[7]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


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


def main():
    args = parse_args()
    start_watchdog(args.watchdog)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n = world
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    import colorvideovdp_b200 as cv
    from colorvideovdp_b200 import distributed as D
    metric = cv.cvvdp(display_name=DISPLAY, device=dev)
    F, H, W, fps = args.frames, args.height, args.width, args.fps

    # this rank's frame shard of every batch item, plus its temporal halo
    lo, hi = D.frame_shard(F, rank, n)
    wlo, whi = D.needed_window(metric, F, fps, lo, hi)
    items = [make_clip(3 + 17 * b, wlo, whi, H, W, args.dtype, dev) for b in range(n)]
    tst = torch.cat([it[0] for it in items], 0)
    ref = torch.cat([it[1] for it in items], 0)
    del items

    def step_device():
        jod, Q = D.predict_frame_sharded(metric, tst, ref, wlo, F, fps)
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
    # ---- device-resident arm ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = metric._ctx.launch_count()
    metric._ctx.profile_enable(True)
    metric._ctx.profile_read()
    if rank == 0:
        clocks.start()
    ms_dev, jod = timed(step_device, args.steps, 0)
    prof = metric._ctx.profile_read()
    metric._ctx.profile_enable(False)
    clk = clocks.stop() if rank == 0 else None
    launches = metric._ctx.launch_count() - launches0
    value = pix_per_step / 1e6 / (ms_dev / 1e3)

    # ---- end-to-end arm: pinned host clips through the public tensor API ----
    e2e = None
    do_e2e = not args.no_e2e
    if do_e2e:  # never pin more than half of the host's available memory (all ranks together)
        need = world * 2 * tst.numel() * tst.element_size()
        try:
            avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        except Exception:
            avail = 1 << 62
        flag = torch.tensor([1 if need < 0.5 * avail else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        do_e2e = bool(flag.item())
    if do_e2e:
        # Host-side data placement of the sharded job.  Default: every rank uploads its whole window (shard + fl-1
        # history frames) through the streaming C-ABI path, which overlaps the upload with compute.
        # --e2e-exchange (opt-in, for short shards, i.e. 4 or more ranks at 60 fps): every rank's host memory holds
        # only the frames it owns and the history frames come from their owners over NVLink (NCCL send/recv), so
        # each byte crosses PCIe once.  It ran at N=2 (171 ms vs 160 ms streamed); the one N=8 attempt of round 1
        # did not finish within the GPU budget, so it is not the default until it has been seen to work there.
        exchange = args.e2e_exchange and world > 1 and (whi - wlo) >= 1.5 * (hi - lo)
        if exchange:
            own_t, own_r = tst[:, :, lo - wlo:hi - wlo], ref[:, :, lo - wlo:hi - wlo]
            tst_h = torch.empty(own_t.shape, dtype=tst.dtype, pin_memory=True).copy_(own_t)
            ref_h = torch.empty(own_r.shape, dtype=ref.dtype, pin_memory=True).copy_(own_r)

            def step_host():
                jod_h, Qd = D.predict_frame_sharded_exchange(metric, tst_h, ref_h, F, fps)
                return jod_h.cpu()  # D2H of the result
        else:
            tst_h = torch.empty(tst.shape, dtype=tst.dtype, pin_memory=True).copy_(tst)
            ref_h = torch.empty(ref.shape, dtype=ref.dtype, pin_memory=True).copy_(ref)

            def step_host():
                jod_h, Qd = D.predict_frame_sharded(metric, tst_h, ref_h, wlo, F, fps)
                return jod_h.cpu()  # D2H of the result

        ms_e2e, jod_h = timed(step_host, max(2, min(args.steps, 3)), 1)
        assert torch.equal(jod_h, jod.cpu()), "host and device arms disagree"
        info = metric._info
        q_bytes = n * info.n_channels * F * info.n_bands * 4
        e2e = {"value": round(pix_per_step / 1e6 / (ms_e2e / 1e3), 2), "unit": "Mpix/s",
               "ms_per_step": round(ms_e2e, 3),
               "h2d_bytes_per_step": int(n * (tst_h.numel() * tst_h.element_size() + ref_h.numel() * ref_h.element_size())),
               "d2h_bytes_per_step": int(q_bytes + 4 * n),
               "input": ("every rank's pinned host memory holds the frames it owns (uploaded once); the fl-1 history "
                         "frames of a shard arrive from their owners by NCCL send/recv over NVLink") if exchange else
                        "pinned host clips streamed through cvvdp_b200_process_host (upload overlapped with compute)"}
        del tst_h, ref_h

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
    kernels = {}
    for p in prof:
        key = p["kind"] + (f"_l{p['level']}" if p["kind"] in ("band", "reduce") else "")
        k = kernels.setdefault(key, {"launches": 0, "ms": 0.0, "bytes": 0.0})
        k["launches"] += p["launches"]; k["ms"] += p["total_ms"]; k["bytes"] += p["algo_bytes"]
    tot_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    breakdown = {name: {"launches": k["launches"], "ms_per_step": round(k["ms"] / args.steps, 4),
                        "share": round(k["ms"] / tot_ms, 4),
                        "algo_gbs": round(k["bytes"] / 1e9 / (k["ms"] / 1e3), 1) if k["ms"] > 0 else None}
                 for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}
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
                "note": "fp32 MUFU/FMA-bound kernel: see DESIGN.md for why it sits below the HBM roofline"}
        whole = args.steps * pix_per_step / n * 2 * 3 * elem_size(args.dtype)  # rank-0 share of B_alg
        roof["whole_path_algo_gbs"] = round(whole / 1e9 / (args.steps * ms_dev / 1e3), 1)

    # ---- CPU baseline (oracle port on a bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline and n == 1:
        cores = os.cpu_count() or 1
        S = cpu_sample_frames(args, cores)
        info = metric._info
        fl = info.filter_len
        f_lo = max(min(fl - 1, F - S), 0)
        w0 = max(f_lo - (fl - 1), 0)
        t_np = tst[:1, :, w0 - wlo:f_lo + S - wlo].cpu().numpy()
        r_np = ref[:1, :, w0 - wlo:f_lo + S - wlo].cpu().numpy()
        if args.dtype == "u16":
            t_np, r_np = t_np.view(np.uint16), r_np.view(np.uint16)
        threads = min(cores, S)
        dt, Qo = oracle_sample(t_np, r_np, fps, w0, f_lo, f_lo + S, F, threads)
        Qg, _ = metric.q_per_ch_from_tensors(tst[:1], ref[:1], F, fps, (f_lo, f_lo + S), wlo)
        Qg = Qg[:, :, f_lo:f_lo + S].cpu().numpy()
        gate = float(np.max(np.abs(Qg - Qo) / (1e-3 * np.abs(Qo) + 1e-5)))
        cpu = {"value": round(S * H * W / 1e6 / dt, 4), "unit": "Mpix/s", "cores": threads, "kind": "port",
               "sample": f"frames [{f_lo},{f_lo + S}) of the same clip (incl. front end of {fl - 1} history frames), "
                         f"numpy oracle port, {threads} threads, {dt:.1f} s",
               "parity_err_over_gate_on_sample": round(gate, 4)}

    line = {"metric": "Mpixels/s", "value": round(value, 2), "unit": "Mpix/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "jod": [round(float(v), 5) for v in jod.flatten().cpu()], "clocks": clk, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "kernels": breakdown}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
