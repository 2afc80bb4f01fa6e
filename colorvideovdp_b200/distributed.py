"""Sharding a batch of clips across the GPUs of one box.

Every Q_per_ch[b, c, f, band] depends only on frames f-(fl-1)..f of item b (causal FIR,
cvvdp_metric.py:554-560), so the work is a flat sequence of (item, frame) units.  It is cut into one
CONTIGUOUS run per rank (`work_shard`): with at least as many items as ranks a rank owns whole items and
reads no temporal halo at all; a single clip degenerates to plain frame sharding, where a rank reads
the fl-1 frames before its range as a halo through the front end only.  Every rank writes its units
into a zero-initialised full-size Q_per_ch buffer; ONE all-reduce (sum) of that small buffer -- each
element is x + 0 + ... + 0, hence bit-exact -- gives every rank the whole tensor, and every rank runs the
identical final pooling.  One process per GPU, launched with torchrun; `torch.distributed` (NCCL over
NVLink on GPUs, gloo in the CPU tests) is the plumbing.
"""
import os

import torch
import torch.distributed as dist


def frame_shard(n_frames, rank, world_size):
    """Contiguous, balanced frame range [lo, hi) of `rank`."""
    return (rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size


def needed_window(metric, n_frames, fps, lo, hi):
    """Clip frames [w_lo, w_hi) that must be present to evaluate frames [lo, hi)."""
    import math
    fl = 1 if n_frames == 1 else int(math.ceil(0.250 * fps / 2) * 2) + 1
    return metric._needed_frames(lo, hi, fl, n_frames)


def predict_frame_sharded(metric, test_win, ref_win, first_frame, n_frames, frames_per_second, group=None):
    """Evaluate this rank's frame shard from its window tensors and combine across ranks.

    test_win / ref_win: BCFHW tensors holding clip frames [first_frame, first_frame + win) -- at least
    `needed_window(...)` of this rank's shard.  Returns (JOD [B] identical on every rank, Q_per_ch
    [B,C,F,L] on the metric's device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = frame_shard(n_frames, rank, world)
    if hi <= lo:
        raise RuntimeError(f"frame sharding needs at least one frame per rank ({n_frames} frames, {world} ranks)")
    Q, _ = metric.q_per_ch_from_tensors(test_win, ref_win, n_frames, frames_per_second, (lo, hi), first_frame)
    Q = Q.to(metric.device)
    if world > 1:
        dist.all_reduce(Q, op=dist.ReduceOp.SUM, group=group)
    jod = metric.do_pooling_and_jods(Q)
    return jod, Q


def work_shard(batch, n_frames, rank, world_size):
    """This rank's contiguous run of the flattened (item, frame) sequence as a list of pieces
    (item, f_lo, f_hi).  batch >= world_size and divisible: whole items only (no temporal halo);
    batch == 1: plain frame sharding."""
    total = batch * n_frames
    lo, hi = (rank * total) // world_size, ((rank + 1) * total) // world_size
    pieces = []
    b = lo // n_frames if n_frames else 0
    while lo < hi:
        f_lo = lo - b * n_frames
        f_hi = min(hi - b * n_frames, n_frames)
        pieces.append((b, f_lo, f_hi))
        lo = b * n_frames + f_hi
        b += 1
    return pieces


def predict_sharded(metric, pieces, batch, n_frames, frames_per_second, group=None):
    """Evaluate this rank's pieces and combine across ranks.

    pieces: list of (item, f_lo, f_hi, first_frame, test_win, ref_win); the window tensors are
    [1,C,win,H,W] (host or device) holding clip frames [first_frame, first_frame + win) of that item --
    at least `needed_window(...)` of [f_lo, f_hi).  Returns (JOD [batch] identical on every rank,
    Q_per_ch [batch,C,F,L] on the metric's device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    Q = None
    for item, f_lo, f_hi, first_frame, test_win, ref_win in pieces:
        Qi, _ = metric.q_per_ch_from_tensors(test_win, ref_win, n_frames, frames_per_second, (f_lo, f_hi), first_frame)
        Qi = Qi.to(metric.device, non_blocking=True)
        if Q is None:
            Q = torch.zeros((batch,) + tuple(Qi.shape[1:]), dtype=torch.float32, device=metric.device)
        Q[item] += Qi[0]  # pieces of one item cover disjoint frames; the rest of Qi is zero
    if Q is None:
        raise RuntimeError(f"no work for this rank ({batch} items x {n_frames} frames over {world} ranks)")
    if world > 1:
        dist.all_reduce(Q, op=dist.ReduceOp.SUM, group=group)
    jod = metric.do_pooling_and_jods(Q)
    # range / NaN warnings and the reference's failure on NaN input, from the front end's device counters
    H, W = pieces[0][4].shape[3], pieces[0][4].shape[4]
    n0 = sum(1 for p in pieces if p[1] == 0)
    metric.report_input_problems(None, n0 * H * W, n0 > 0)
    return jod, Q


def bind_to_gpu_numa_node(device_index):
    """Pin this process (CPU affinity + preferred memory node) to the NUMA node of its GPU, so that the pinned
    host buffers it allocates afterwards are local to the GPU's PCIe root port: with one process per GPU on a
    two-socket host, remote pinned memory makes every upload cross the socket interconnect.  Best effort --
    returns a short description of what was done (for the bench line)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpulist = open(f"{base}/local_cpulist").read().strip()
    except Exception as e:  # no sysfs entry, attribute missing in this torch build, ...
        return f"unbound ({type(e).__name__})"
    cpus = set()
    for part in cpulist.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    allowed = os.sched_getaffinity(0)
    cpus &= allowed
    if not cpus or cpus == allowed and node < 0:
        return f"unbound (node {node}, one NUMA domain)"
    os.sched_setaffinity(0, cpus)
    how = f"cpus {cpulist}"
    if node >= 0:
        try:  # set_mempolicy(MPOL_PREFERRED, {node}): x86-64 syscall 238
            import ctypes
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            how += f", memory node {node}" + ("" if rc == 0 else " (set_mempolicy refused)")
        except Exception:
            pass
    return "bound to " + how


def _overlap(a_lo, a_hi, b_lo, b_hi):
    lo, hi = max(a_lo, b_lo), min(a_hi, b_hi)
    return (lo, hi) if hi > lo else None


def exchange_plan(metric, n_frames, frames_per_second, world_size):
    """For every rank: its shard [lo, hi) and needed window [w_lo, w_hi); and the list of transfers
    (src_rank, dst_rank, f_lo, f_hi): clip frames [f_lo, f_hi) owned by src that dst needs as temporal
    history (or, with symmetric padding of a short first shard, as mirrored future frames).  Pure
    function of the arguments: every rank derives the identical plan."""
    shards = [frame_shard(n_frames, r, world_size) for r in range(world_size)]
    windows = [needed_window(metric, n_frames, frames_per_second, lo, hi) for lo, hi in shards]
    transfers = []
    for dst, ((lo, hi), (wlo, whi)) in enumerate(zip(shards, windows)):
        for src, (slo, shi) in enumerate(shards):
            if src == dst:
                continue
            for part in (_overlap(wlo, min(lo, whi), slo, shi), _overlap(max(hi, wlo), whi, slo, shi)):
                if part:
                    transfers.append((src, dst, part[0], part[1]))
    return shards, windows, transfers


def predict_frame_sharded_exchange(metric, test_own, ref_own, n_frames, frames_per_second, group=None):
    """Frame-sharded prediction where every rank holds ONLY its own frames [lo, hi) of the batch
    (host or device tensors, BCFHW): the fl-1 history frames a shard needs are fetched from the ranks
    that own them, device to device (NCCL send/recv over NVLink; gloo in the CPU tests), instead of
    being uploaded a second time over PCIe.  Each input byte crosses PCIe exactly once; with 8 ranks
    and 15-frame shards at 60 fps that halves the host->device volume of `predict_frame_sharded`.
    Returns (JOD [B] identical on every rank, Q_per_ch [B,C,F,L] on the metric's device)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shards, windows, transfers = exchange_plan(metric, n_frames, frames_per_second, world)
    lo, hi = shards[rank]
    wlo, whi = windows[rank]
    if hi <= lo:
        raise RuntimeError(f"frame sharding needs at least one frame per rank ({n_frames} frames, {world} ranks)")
    if test_own.shape[2] != hi - lo or ref_own.shape[2] != hi - lo:
        raise RuntimeError(f"rank {rank} must hold exactly its shard, frames [{lo}, {hi})")
    dev = metric.device if metric.device.type == "cuda" else test_own.device
    wins = []
    for own in (test_own, ref_own):
        win = torch.empty(own.shape[:2] + (whi - wlo,) + own.shape[3:], dtype=own.dtype, device=dev)
        win[:, :, lo - wlo:hi - wlo].copy_(own, non_blocking=True)  # the only host->device traffic
        wins.append(win)
    if world > 1 and transfers:
        # One message per (transfer, video, batch item, channel): win[b, c, f_lo:f_hi] is a contiguous run of
        # whole frames, so it goes on the wire and lands in place without staging copies, and no message
        # comes near 2 GiB (8 items x 15 4K frames of one video in one piece would be 3 GB).
        ops = []
        for src, dst, flo, fhi in transfers:
            if rank not in (src, dst):
                continue
            for win in wins:
                for b in range(win.shape[0]):
                    for c in range(win.shape[1]):
                        piece = win[b, c, flo - wlo:fhi - wlo]
                        assert piece.is_contiguous()
                        if src == rank:
                            ops.append(dist.P2POp(dist.isend, piece, dst, group=group))
                        else:
                            ops.append(dist.P2POp(dist.irecv, piece, src, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    resident = True if dev.type != "cuda" else None  # mock device in the CPU tests: keep the resident code path
    Q, _ = metric.q_per_ch_from_tensors(wins[0], wins[1], n_frames, frames_per_second, (lo, hi), wlo, _resident=resident)
    Q = Q.to(metric.device)
    if world > 1:
        dist.all_reduce(Q, op=dist.ReduceOp.SUM, group=group)
    jod = metric.do_pooling_and_jods(Q)
    return jod, Q
