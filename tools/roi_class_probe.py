"""Where does the ROIAlign time go?  Splits the benchmark's RoIs by samples-per-bin class (gh, gw) and times the
forward / backward of each subset separately (CUDA events).  Diagnostic; evidence for profiles/."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cddmsl_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc"]
    dev = torch.device("cuda:0")
    g = synth.generator(cfg.seed + 1000)
    feat = synth.make_features(cfg, g).to(dev)
    rois = synth.make_rois(cfg, g)
    P, scale = cfg.pooled, 1.0 / cfg.stride
    N, C = cfg.n_images, cfg.channels
    Hf, Wf = cfg.feat_hw
    h = ((rois[:, 4] - rois[:, 2]) * scale).clamp(min=0)
    w = ((rois[:, 3] - rois[:, 1]) * scale).clamp(min=0)
    gh = torch.ceil(h / P).clamp(min=1).long()
    gw = torch.ceil(w / P).clamp(min=1).long()
    out = {"workload": cfg.name, "rois": int(rois.shape[0]), "classes": []}
    tot_f = tot_b = 0.0
    for lo, hi, name in [(1, 1, "gh=1"), (2, 2, "gh=2"), (3, 99, "gh>=3")]:
        for wlo, whi, wname in [(1, 1, "gw=1"), (2, 2, "gw=2"), (3, 99, "gw>=3")]:
            sel = ((gh >= lo) & (gh <= hi) & (gw >= wlo) & (gw <= whi)).nonzero()[:, 0]
            if sel.numel() == 0:
                continue
            r = rois[sel].to(dev)
            o = ops.roi_align(feat, r, scale, P, P, cfg.sampling_ratio, True)
            tf = timeit(lambda: ops.roi_align(feat, r, scale, P, P, cfg.sampling_ratio, True))
            tb = timeit(lambda: ops.roi_align_backward(o, r, scale, P, P, N, C, Hf, Wf, cfg.sampling_ratio, True))
            tot_f += tf
            tot_b += tb
            out["classes"].append({"class": f"{name},{wname}", "rois": int(sel.numel()), "fwd_ms": round(tf, 3),
                                   "bwd_ms": round(tb, 3), "fwd_us_per_roi": round(1e3 * tf / sel.numel(), 3),
                                   "bwd_us_per_roi": round(1e3 * tb / sel.numel(), 3)})
    out["sum_fwd_ms"], out["sum_bwd_ms"] = round(tot_f, 3), round(tot_b, 3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
