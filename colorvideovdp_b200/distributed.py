"""Frame sharding across the GPUs of one box.

Every Q_per_ch[b, c, f, band] depends only on frames f-(fl-1)..f of item b (causal FIR,
cvvdp_metric.py:554-560), so rank r evaluates a contiguous frame range of every batch item (reading
fl-1 halo frames before it) into a zero-initialised full-size Q_per_ch buffer.  ONE all-reduce (sum)
of that small buffer -- each element is x + 0 + ... + 0, hence bit-exact -- gives every rank the
whole tensor, and every rank runs the identical final pooling.  One process per GPU, launched with
torchrun; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) is the plumbing.
"""
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
