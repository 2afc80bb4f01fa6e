"""Coloured heat maps of VIDEOS, generated from the UNMODIFIED reference.  Run in the build container only:

    python tests/golden/make_golden_colourmap.py

The reference colours each BLOCK of frames with tone-curve statistics of that block (cvvdp_metric.py:350-401,
visualize_diff_map.py:23-45); its block size is 1 on the CPU and "what fits in memory" on a GPU.  Two fixtures per
clip, both produced by the reference's own code:
  *_perframe  -- the CPU run as it is (one frame per block);
  *_oneblock  -- what the reference computes on a GPU for a clip that fits in one block: its own
                 process_block_of_frames outputs (raw map and R[:,0] of every frame, captured by a wrapper
                 around the unmodified method) concatenated and passed ONCE through its own visualize_diff_map.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402
import synth  # noqa: E402

pycvvdp = ref_loader.load(prefer_staged=False)
from pycvvdp.visualize_diff_map import visualize_diff_map  # noqa: E402

DEV = torch.device("cpu")
torch.set_num_threads(8)


def run(name, test, ref, fps, display, kind):
    m = pycvvdp.cvvdp(display_name=display, device=DEV, heatmap=kind, quiet=True)
    captured = []
    inner = m.process_block_of_frames

    def spy(R, vid_sz, temp_ch, lpyr, is_image):
        out = inner(R, vid_sz, temp_ch, lpyr, is_image)
        captured.append((R[:, 0].clone(), out[1].clone()))
        return out

    m.process_block_of_frames = spy
    with torch.no_grad():
        q, s = m.predict(test, ref, dim_order="BCFHW", frames_per_second=fps)
    per_frame = s["heatmap"].numpy()
    ctx = torch.cat([c[0] for c in captured], dim=1)          # [1,F,H,W]
    raw = torch.cat([c[1] for c in captured], dim=2)          # [1,1,F,H,W]
    one_block = visualize_diff_map(raw, context_image=ctx, colormap_type=kind).detach().type(torch.float16).numpy()[None]
    for tag, hm, blk in (("perframe", per_frame, 1), ("oneblock", one_block, 0)):
        meta = {"dim_order": "BCFHW", "fps": fps, "display": display, "padding": "replicate", "heatmap": kind,
                "hm_block": blk,
                "reference": "gfxdisp/ColorVideoVDP pycvvdp 0.5.4 (params 0.5.6), torch %s CPU" % torch.__version__}
        np.savez_compressed(os.path.join(HERE, f"{name}_{tag}.npz"), test=test, ref=ref, meta=np.asarray(json.dumps(meta)),
                            jod=np.asarray(q.cpu().numpy(), dtype=np.float32), Q_per_ch=s["Q_per_ch"].astype(np.float32),
                            rho_band=np.asarray(s["rho_band"], dtype=np.float64), heatmap=hm)
        print(f"{name}_{tag}: JOD={float(q):.6f} heatmap{hm.shape} mean {hm.astype(np.float32).mean():.4f}")


if __name__ == "__main__":
    tst, ref = synth.make_pair_u8(91, 6, 64, 96)
    # a strong luminance gradient across the frames so that per-frame and per-block tone curves differ visibly
    ramp = np.linspace(0.35, 1.0, 6, dtype=np.float32)[None, None, :, None, None]
    tst = np.clip(tst.astype(np.float32) * ramp, 0, 255).astype(np.uint8)
    ref = np.clip(ref.astype(np.float32) * ramp, 0, 255).astype(np.uint8)
    run("cmap_vid_u8_6x64x96_fhd_thr", tst, ref, 30, "standard_fhd", "threshold")
    run("cmap_vid_u8_6x64x96_fhd_supra", tst, ref, 30, "standard_fhd", "supra-threshold")
