import os, sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import colorvideovdp_b200 as cv, synth
dev = torch.device('cuda:0')
tst, ref = synth.make_pair_u8(5, 8, 1080, 1920)
td = torch.from_numpy(tst).to(dev).repeat(1, 1, 8, 1, 1)[:, :, :60].contiguous(); rd = torch.from_numpy(ref).to(dev).repeat(1, 1, 8, 1, 1)[:, :, :60].contiguous()
for hm in (None, "raw", "supra-threshold"):
    m = cv.cvvdp(display_name="standard_fhd", device=dev, heatmap=hm)
    for _ in range(2): m.predict(td, rd, frames_per_second=30)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): m.predict(td, rd, frames_per_second=30)
    torch.cuda.synchronize(); print(hm, "predict %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
    Q, h = m.compute_q_per_ch(cv.video_source_array(td, rd, 30, display_photometry=m.display_photometry))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): m.compute_q_per_ch(cv.video_source_array(td, rd, 30, display_photometry=m.display_photometry))
    torch.cuda.synchronize(); print(hm, "compute_q_per_ch (device only) %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
    m._ctx.profile_enable(True); m._ctx.profile_read()
    m.compute_q_per_ch(cv.video_source_array(td, rd, 30, display_photometry=m.display_photometry))
    tot = {}
    for k in m._ctx.profile_read(): tot[k["kind"]] = tot.get(k["kind"], 0) + k["total_ms"]
    print("   ", {k: round(v, 2) for k, v in tot.items()})
