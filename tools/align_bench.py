"""Alignment loss (forward + backward of the local rows) at the gathered sizes of BASELINE.json's configs:
the C-ABI kernels vs eager PyTorch (normalise, matmul, two cross-entropies, autograd) on the same B200.

    python tools/align_bench.py > profiles/align_rN.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cddmsl_b200 import ops, synth  # noqa: E402


def _time(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    out = []
    for world, n_local, dim in [(1, 16, 256), (1, 256, 256), (8, 16, 256), (8, 256, 256), (8, 256, 1024), (8, 1024, 256)]:
        g = synth.generator(world + n_local)
        a = torch.randn(world * n_local, dim, generator=g).to(dev)
        b = torch.randn(world * n_local, dim, generator=g).to(dev)
        packs = [ops.align_pack(a[r * n_local:(r + 1) * n_local], b[r * n_local:(r + 1) * n_local]) for r in range(world)]
        packed_all = torch.stack([p for p, _ in packs])
        norms0 = packs[0][1]

        def ours():
            ops.align_pack(a[:n_local], b[:n_local])
            return ops.align_loss(packed_all, norms0, 0, None, True)

        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)

        def eager():
            ar.grad = br.grad = None
            x = ar / ar.norm(dim=1, keepdim=True)
            y = br / br.norm(dim=1, keepdim=True)
            s = x @ y.t()
            t = torch.arange(s.shape[0], device=dev)
            loss = (torch.nn.functional.cross_entropy(s, t) + torch.nn.functional.cross_entropy(s.t(), t)) / 2
            loss.backward()
            return loss

        kern = None
        try:  # per-kernel device times (CUPTI); informational
            from torch.profiler import ProfilerActivity, profile
            ours()
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(5):
                    ours()
                torch.cuda.synchronize()
            kern = {e.key[:48]: round(e.device_time_total / e.count, 2) for e in prof.key_averages() if e.device_time_total > 0}
        except Exception as ex:  # noqa: BLE001
            kern = {"unavailable": str(ex)[:80]}
        l_o = ours()[0].item()
        l_e = eager().item()
        out.append({"world": world, "n_local": n_local, "n": world * n_local, "D": dim, "ours_ms": round(_time(ours), 4),
                    "eager_torch_ms": round(_time(eager), 4), "loss_ours": l_o, "loss_eager": l_e, "kernels_us": kern})
    print(json.dumps({"device": torch.cuda.get_device_name(0), "note": "one rank's work; all-gather excluded (emulated "
                      "by stacking); eager computes gradients of all n rows, the reference semantics keep n_local",
                      "cases": out}, indent=1))


if __name__ == "__main__":
    main()
