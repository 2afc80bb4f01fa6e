"""Frame sharding with world_size 2 over gloo on the CPU (mock device): each rank evaluates its frame
range from a window of the clip, one all-reduce of Q_per_ch, identical pooling on both ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from emu_util import mock_device  # noqa: E402,F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colorvideovdp_b200 as cv
    from colorvideovdp_b200 import cvvdp_metric, distributed as D
    import emu_util
    import synth
    cvvdp_metric._set_mock_library_for_tests(emu_util.emu_library())
    F, fps = 9, 30
    tst, ref = synth.make_pair_u8(51, F, 36, 48)
    tb = np.concatenate([tst, np.clip(tst.astype(np.int16) + 9, 0, 255).astype(np.uint8)], 0)  # batch of 2
    m = cv.cvvdp(display_name="standard_fhd", temp_padding="symmetric")
    lo, hi = D.frame_shard(F, rank, world)
    wlo, whi = D.needed_window(m, F, fps, lo, hi)
    tw, rw = torch.from_numpy(tb[:, :, wlo:whi].copy()), torch.from_numpy(ref[:, :, wlo:whi].copy())
    jod, Q = D.predict_frame_sharded(m, tw, rw, wlo, F, fps)
    # same shard, but every rank holds only its own frames: history frames come from their owners
    to, ro = torch.from_numpy(tb[:, :, lo:hi].copy()), torch.from_numpy(ref[:, :, lo:hi].copy())
    jod_x, Q_x = D.predict_frame_sharded_exchange(m, to, ro, F, fps)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), jod=jod.numpy(), Q=Q.numpy(), shard=np.asarray([lo, hi, wlo, whi]),
             jod_x=jod_x.numpy(), Q_x=Q_x.numpy())
    if rank == 0:
        j_full, s_full = m.predict(tb, ref, frames_per_second=fps)
        np.savez(os.path.join(out_dir, "full.npz"), jod=j_full.numpy(), Q=s_full["Q_per_ch"])
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1, full = (np.load(tmp_path / n) for n in ("r0.npz", "r1.npz", "full.npz"))
    assert list(r0["shard"][:2]) == [0, 4] and list(r1["shard"][:2]) == [4, 9]
    assert r1["shard"][2] == 0  # fl = 9 at 30 fps: the second shard needs the whole history
    assert np.array_equal(r0["Q"], r1["Q"]) and np.array_equal(r0["jod"], r1["jod"])
    assert np.array_equal(r0["Q"], full["Q"])  # sharded == single process, bit for bit
    assert np.array_equal(r0["jod"], full["jod"])
    # halo exchange instead of double upload: identical bits again
    assert np.array_equal(r0["Q_x"], full["Q"]) and np.array_equal(r1["Q_x"], full["Q"])
    assert np.array_equal(r0["jod_x"], full["jod"]) and np.array_equal(r1["jod_x"], full["jod"])


def test_exchange_plan_covers_every_window(mock_device):
    """Every frame a rank needs and does not own is delivered exactly once, by its owner."""
    import colorvideovdp_b200 as cv
    from colorvideovdp_b200 import distributed as D
    for padding in ("replicate", "symmetric"):
        m = cv.cvvdp(display_name="standard_fhd", temp_padding=padding)
        for F, fps, world in ((120, 60, 8), (9, 30, 2), (20, 60, 4), (7, 24, 3)):
            shards, windows, transfers = D.exchange_plan(m, F, fps, world)
            for r, ((lo, hi), (wlo, whi)) in enumerate(zip(shards, windows)):
                have = set(range(lo, hi))
                for src, dst, flo, fhi in transfers:
                    if dst == r:
                        assert shards[src][0] <= flo and fhi <= shards[src][1]
                        got = set(range(flo, fhi))
                        assert not (got & have)
                        have |= got
                assert have == set(range(wlo, whi)), (padding, F, fps, world, r)


def test_frame_shard_partition():
    from colorvideovdp_b200.distributed import frame_shard
    for F in (1, 7, 120, 121):
        for world in (1, 2, 3, 8):
            parts = [frame_shard(F, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == F
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def _worker_exchange4(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colorvideovdp_b200 as cv
    from colorvideovdp_b200 import cvvdp_metric, distributed as D
    import emu_util
    import synth
    cvvdp_metric._set_mock_library_for_tests(emu_util.emu_library())
    F, fps = 20, 60  # fl = 17: every shard of 5 frames needs history from up to four earlier ranks
    tst, ref = synth.make_pair_u8(52, F, 24, 32)
    m = cv.cvvdp(display_name="standard_fhd")
    lo, hi = D.frame_shard(F, rank, world)
    jod, Q = D.predict_frame_sharded_exchange(m, torch.from_numpy(tst[:, :, lo:hi].copy()), torch.from_numpy(ref[:, :, lo:hi].copy()), F, fps)
    np.savez(os.path.join(out_dir, f"x{rank}.npz"), jod=jod.numpy(), Q=Q.numpy())
    if rank == 0:
        j_full, s_full = m.predict(tst, ref, frames_per_second=fps)
        np.savez(os.path.join(out_dir, "xfull.npz"), jod=j_full.numpy(), Q=s_full["Q_per_ch"])
    dist.destroy_process_group()


def test_halo_exchange_world4_gloo(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_exchange4, args=(4, port, str(tmp_path)), nprocs=4, join=True)
    full = np.load(tmp_path / "xfull.npz")
    for r in range(4):
        z = np.load(tmp_path / f"x{r}.npz")
        assert np.array_equal(z["Q"], full["Q"]) and np.array_equal(z["jod"], full["jod"])


def _worker_runs(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import colorvideovdp_b200 as cv
    from colorvideovdp_b200 import cvvdp_metric, distributed as D
    import emu_util
    import synth
    cvvdp_metric._set_mock_library_for_tests(emu_util.emu_library())
    B, F, fps = 3, 10, 30  # 30 units over 2 ranks: item 1 is split between them (halo on rank 1), items 0 and 2 are whole
    clips = [synth.make_pair_u8(70 + b, F, 36, 48) for b in range(B)]
    m = cv.cvvdp(display_name="standard_fhd")
    pieces = []
    for item, f_lo, f_hi in D.work_shard(B, F, rank, world):
        wlo, whi = D.needed_window(m, F, fps, f_lo, f_hi)
        tst, ref = clips[item]
        pieces.append((item, f_lo, f_hi, wlo, torch.from_numpy(tst[:, :, wlo:whi].copy()), torch.from_numpy(ref[:, :, wlo:whi].copy())))
    jod, Q = D.predict_sharded(m, pieces, B, F, fps)
    np.savez(os.path.join(out_dir, f"w{rank}.npz"), jod=jod.numpy(), Q=Q.numpy(),
             pieces=np.asarray([p[:3] for p in pieces]))
    if rank == 0:
        tb = np.concatenate([c[0] for c in clips], 0)
        rb = np.concatenate([c[1] for c in clips], 0)
        j_full, s_full = m.predict(tb, rb, frames_per_second=fps)
        np.savez(os.path.join(out_dir, "wfull.npz"), jod=j_full.numpy(), Q=s_full["Q_per_ch"])
    dist.destroy_process_group()


def test_run_sharding_world2_gloo(tmp_path):
    """Contiguous runs of the flattened (item, frame) sequence: whole items and a split item, bit-identical to
    the single-process batch prediction."""
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_runs, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0, w1, full = (np.load(tmp_path / n) for n in ("w0.npz", "w1.npz", "wfull.npz"))
    assert w0["pieces"].tolist() == [[0, 0, 10], [1, 0, 5]] and w1["pieces"].tolist() == [[1, 5, 10], [2, 0, 10]]
    assert np.array_equal(w0["Q"], w1["Q"]) and np.array_equal(w0["jod"], w1["jod"])
    assert np.array_equal(w0["Q"], full["Q"]) and np.array_equal(w0["jod"], full["jod"])


def test_work_shard_covers_every_unit_once():
    from colorvideovdp_b200.distributed import work_shard
    for B, F, world in ((8, 120, 8), (8, 120, 2), (1, 120, 8), (3, 10, 2), (2, 9, 4), (5, 7, 3), (1, 1, 1)):
        seen = []
        for r in range(world):
            for item, lo, hi in work_shard(B, F, r, world):
                assert 0 <= lo < hi <= F
                seen += [(item, f) for f in range(lo, hi)]
        assert seen == [(b, f) for b in range(B) for f in range(F)], (B, F, world)
    assert work_shard(8, 120, 3, 8) == [(3, 0, 120)]  # as many items as ranks: whole items, no halo
