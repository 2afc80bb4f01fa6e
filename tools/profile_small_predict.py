import cProfile, pstats, sys, os, torch, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import colorvideovdp_b200 as cv, synth
dev = torch.device('cuda:0')
m = cv.cvvdp(display_name='standard_fhd', device=dev)
tst, ref = synth.make_pair_u8(5, 1, 256, 256)
td, rd = torch.from_numpy(tst).to(dev), torch.from_numpy(ref).to(dev)
for _ in range(50): m.predict(td, rd)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): m.predict(td, rd)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
