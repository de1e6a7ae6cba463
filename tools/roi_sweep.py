#!/usr/bin/env python
"""BASELINE.json configs[4]: ROIAlign microbench sweep — R in 1k..256k, 7x7 and 14x14 bins, sampling_ratio 0/2,
fwd and bwd, this repo vs torchvision-CUDA on the same inputs (CUDA events, median).  [16,C,38,63] map; C=1024 up to
16k RoIs, C=256 beyond (bounds the pooled tensor: 256k x 256 x 14 x 14 x 4 B = 51 GB)."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cddmsl_b200 import ops, synth  # noqa: E402


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-torchvision", action="store_true")
    a = ap.parse_args()
    from torchvision.ops import roi_align as tv

    dev = torch.device("cuda:0")
    peak = 6550.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for r in (1024, 4096, 16384, 65536, 262144):
        c = 1024 if r <= 16384 else 256
        cfg = synth.PathConfig("sweep", 16, 600, 1000, r // 16, 20, channels=c, seed=4)
        g = synth.generator(4)
        feat = synth.make_features(cfg, g).to(dev)
        rois = synth.make_rois(cfg, g).to(dev)
        for p in (14, 7):
            for sr in (0, 2):
                out = ops.roi_align(feat, rois, 1 / 16, p, p, sr, True)
                f = timeit(lambda: ops.roi_align(feat, rois, 1 / 16, p, p, sr, True), a.iters)
                b = timeit(lambda: ops.roi_align_backward(out, rois, 1 / 16, p, p, 16, c, 38, 63, sr, True), a.iters)
                byts = r * (4 * c * p * p + 20) + 16 * c * 38 * 63 * 4
                row = {"rois": r, "channels": c, "bins": p, "sampling_ratio": sr, "fwd_ms": f, "bwd_ms": b,
                       "fwd_GBps": byts / f / 1e6, "bwd_GBps": byts / b / 1e6, "fwd_frac_of_hbm": byts / f / 1e6 / peak,
                       "bwd_frac_of_hbm": byts / b / 1e6 / peak, "rois_per_s_fwd_bwd": r / ((f + b) * 1e-3)}
                if not a.no_torchvision and r <= 65536:
                    tf = timeit(lambda: tv(feat, rois, (p, p), 1 / 16, sr, True), max(2, a.iters // 2), warm=1)
                    tb = timeit(lambda: torch.ops.torchvision._roi_align_backward(out, rois, 1 / 16, p, p, 16, c, 38, 63, sr,
                                                                                  True), max(2, a.iters // 2), warm=1)
                    row.update(torchvision_fwd_ms=tf, torchvision_bwd_ms=tb)
                rows.append(row)
                print(json.dumps(row), flush=True)
                del out
        del feat, rois
        torch.cuda.empty_cache()
    if a.out:
        json.dump({"gpu": torch.cuda.get_device_name(0), "peak_GBps": peak, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
