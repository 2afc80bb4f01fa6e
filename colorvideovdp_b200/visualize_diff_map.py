"""Coloured heat maps ('threshold' / 'supra-threshold'), pycvvdp/visualize_diff_map.py:23-106.

Not part of the hot path: the raw difference map comes from the CUDA kernels (fp16, 0..1); this module
only applies the reference's colour look-up table and its tone-mapped grey-scale context image with
plain torch ops on the metric's device.  The reference tone-maps each *block of frames* with global
statistics of that block (its block size depends on free GPU memory), so coloured maps are only
reproducible up to that partition; here the statistics are taken per call over all frames given, which
equals the reference whenever it processes the clip in one block (always true for images).
"""
import torch


def _interp1(x, v, x_q):
    """pycvvdp/interp.py:22-31, 81-89 (bucketize-based linear interpolation with edge clamping)."""
    shp = x_q.shape
    x_q = x_q.flatten()
    imax = torch.bucketize(x_q, x)
    imax[imax >= x.shape[0]] = x.shape[0] - 1
    imin = (imax - 1).clamp(0, x.shape[0] - 1)
    ifrc = (x_q - x[imin]) / (x[imax] - x[imin] + 0.000001)
    ifrc[imax == imin] = 0.0
    ifrc[ifrc < 0.0] = 0.0
    return (v[imin] * (1.0 - ifrc) + v[imax] * ifrc).reshape(shp)


def vis_tonemap(b, dr):
    """pycvvdp/visualize_diff_map.py:23-45: histogram-based tone mapping of a log-luminance image."""
    t = 3.0
    b_min, b_max = torch.min(b), torch.max(b)
    if b_max - b_min < dr:
        return (b - b_min) / (b_max - b_min + 1e-3) * dr + (1 - dr) / 2
    b_scale = torch.linspace(b_min, b_max, 1024, device=b.device)
    b_p = torch.histc(b, 1024, float(b_min), float(b_max))
    b_p = b_p / torch.sum(b_p)
    sum_b_p = torch.sum(torch.pow(b_p, 1.0 / t))
    dy = torch.pow(b_p, 1.0 / t) / sum_b_p
    v = torch.cumsum(dy, 0) * dr + (1.0 - dr) / 2.0
    return _interp1(b_scale, v, b)


def visualize_diff_map(diff_map, context_image=None, colormap_type="supra-threshold"):
    """diff_map [1,1,F,H,W] in 0..1, context_image [1,F,H,W] (absolute luminance) -> fp16 [3,F,H,W]."""
    diff_map = torch.clamp(diff_map.float(), 0.0, 1.0)
    if context_image is None:
        tmo_img = torch.ones_like(diff_map) * 0.5
    else:
        y = context_image.float()
        clampval = torch.min(y[y > 0.0])
        tmo_img = vis_tonemap(torch.log(torch.clamp(y, min=clampval)), 0.6)
    dev = diff_map.device
    if colormap_type == "threshold":
        color_map = torch.tensor([[0.2, 0.2, 1.0], [0.2, 1.0, 1.0], [0.2, 1.0, 0.2], [1.0, 1.0, 0.2], [1.0, 0.2, 0.2]], device=dev)
        color_map_in = torch.tensor([0.00, 0.25, 0.50, 0.75, 1.00], device=dev) * 0.1
    elif colormap_type == "supra-threshold":
        color_map = torch.tensor([[0.2, 1.0, 1.0], [1.0, 1.0, 1.0], [1.0, 1.0, 0.2]], device=dev)
        color_map_in = torch.tensor([0.0, 0.5, 1.0], device=dev) * 0.3
    elif colormap_type == "monochromatic":
        color_map = torch.tensor([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0]], device=dev)
        color_map_in = torch.tensor([0.0, 1.0], device=dev)
    else:
        raise RuntimeError(f"Unknown colormap: {colormap_type}")
    frame_count, h, w = diff_map.shape[-3], diff_map.shape[-2], diff_map.shape[-1]
    cmap = torch.empty([3, frame_count, h, w], device=dev, dtype=torch.float16)
    color_map_l = color_map[:, 0:1] * 0.212656 + color_map[:, 1:2] * 0.715158 + color_map[:, 2:3] * 0.072186
    color_map_ch = color_map / (torch.cat([color_map_l] * 3, 1) + 0.0001)
    for c in range(3):
        cmap[c:c + 1, ...] = _interp1(color_map_in, color_map_ch[:, c], diff_map).type(torch.float16)
    return (cmap * tmo_img).clip(0.0, 1.0)


def colorize_heatmap(raw_heatmap, vid_source, colormap_type, device, context=None):
    """raw fp16 heat map [1,1,F,H,W] -> coloured fp16 CPU tensor [1,3,F,H,W].  The context image is the
    TEST achromatic temporal channel in the reference (cvvdp_metric.py:400: R[:,0]); its sustained
    response is approximated here by the test frame's DKL achromatic plane, which is what the reference
    uses for images and differs for videos only by the temporal low-pass of the context picture."""
    F = raw_heatmap.shape[2]
    if context is None:
        frames = [vid_source.get_test_frame(f, device=device, colorspace="DKLd65")[:, 0:1] for f in range(F)]
        context = torch.cat(frames, dim=2)[:, 0]  # [1,F,H,W]
    out = visualize_diff_map(raw_heatmap.to(device), context_image=context, colormap_type=colormap_type)
    return out.unsqueeze(0).to(torch.float16).cpu()
