#!/usr/bin/env python
"""BASELINE.json configs[3]: LVIS-scale open-vocabulary head, 1203 concept embeddings x 64k RoIs, fwd+bwd.
Times the tcgen05 3xTF32 path, the generic CUDA-core kernel and eager PyTorch (cuBLAS fp32, TF32 off) on the
same inputs; reports RoIs/s, achieved TFLOP/s (algorithmic 2*R*D*(K+1) per GEMM, two GEMMs) and parity."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cddmsl_b200 import _lib, ops  # noqa: E402


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
    ap.add_argument("--rois", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only-tc", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    r, d, k = a.rois, 1024, 1203
    x = torch.randn(r, d, generator=g).to(dev)
    w = torch.randn(k, d, generator=g).to(dev)
    w_bg = torch.zeros(1, d, device=dev)
    gt = torch.randint(0, k + 1, (r,), generator=g).to(dev)
    one = torch.ones(1, device=dev)
    flops = 2 * 2.0 * r * d * (k + 1)

    def ours():
        return ops.clip_head_loss(x, w, w_bg, gt, 0.01, ops.LOSS_FOCAL, 0.5, 0.2, one, False, True, True)

    res = {"rois": r, "D": d, "K": k, "algorithmic_flop": flops}
    _lib.tune("head_tc", 1)
    l_tc, s_tc, dx_tc, _ = ours()
    t_tc = timeit(ours, a.iters)
    res["tcgen05_3xTF32"] = {"ms": t_tc, "rois_per_s": r / (t_tc * 1e-3), "algorithmic_TFLOPs": flops / (t_tc * 1e-3) / 1e12,
                             "tensor_TFLOPs_issued": 3 * flops / (t_tc * 1e-3) / 1e12}
    if not a.only_tc:
        _lib.tune("head_tc", 0)
        l_cc, s_cc, dx_cc, _ = ours()
        t_cc = timeit(ours, max(2, a.iters // 2), warm=1)
        res["cuda_core_kernel"] = {"ms": t_cc, "rois_per_s": r / (t_cc * 1e-3)}
        _lib.tune("head_tc", 2)
        torch.backends.cuda.matmul.allow_tf32 = False

        def eager():
            xx = x.detach().requires_grad_(True)
            nx = F.normalize(xx, p=2.0, dim=1)
            s = torch.cat((nx @ F.normalize(w, p=2.0, dim=1).t(), F.linear(nx, w_bg)), dim=1) / 0.01
            ce = F.cross_entropy(s, gt, reduction="none")
            p = F.softmax(s, dim=-1)
            pt = p[torch.arange(p.size(0), device=dev), gt]
            loss = ce * ((1 - pt) ** 0.5)
            lw = torch.ones(loss.size(0), device=dev)
            lw[gt == k] = 0.2
            loss = (loss * lw).mean()
            loss.backward()
            return loss.detach(), s.detach(), xx.grad

        l_e, s_e, dx_e = eager()
        t_e = timeit(eager, max(2, a.iters // 2), warm=1)
        res["eager_pytorch_fp32"] = {"ms": t_e, "rois_per_s": r / (t_e * 1e-3)}
        sc = float(s_e.abs().max())
        res["parity_vs_eager_fp32"] = {
            "scores_max_abs_err_tc": float((s_tc - s_e).abs().max()), "scores_scale": sc,
            "scores_max_abs_err_cuda_core": float((s_cc - s_e).abs().max()),
            "loss_tc": float(l_tc), "loss_cuda_core": float(l_cc), "loss_eager": float(l_e),
            "dx_max_abs_err_tc": float((dx_tc - dx_e).abs().max()), "dx_scale": float(dx_e.abs().max())}
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
