#!/usr/bin/env python
"""Times the same synthetic inputs through (a) this repo's kernels and (b) what the reference runs on a GPU today
— torchvision's CUDA roi_align / nms and eager PyTorch for the head and the alignment loss ("the
Blackwell kernel to beat").  CUDA events, warm-up, median of `--iters`.  Also sweeps NMS sizes (boxes/s, ms).
Writes one JSON document to stdout (and --out).  Not part of bench.py's contract; evidence for profiles/."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cddmsl_b200 import ops, synth  # noqa: E402
from cddmsl_b200.layers import batched_nms  # noqa: E402


def timeit(fn, iters, warm=3):
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
    ap.add_argument("--workload", default="voc")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torchvision
    from torchvision.ops import roi_align as tv_roi_align
    from torchvision.ops.boxes import batched_nms as tv_batched_nms

    dev = torch.device("cuda:0")
    cfg = synth.CONFIGS[args.workload]
    g = synth.generator(cfg.seed)
    feat = synth.make_features(cfg, g).to(dev)
    rois = synth.make_rois(cfg, g).to(dev)
    x, w, w_bg, gt = [t.to(dev) for t in synth.make_head_inputs(cfg, g)]
    a_is, a_it, a_rs, a_rt = [t.to(dev) for t in synth.make_align_inputs(cfg, g)]
    P, scale = cfg.pooled, 1.0 / cfg.stride
    N, C = cfg.n_images, cfg.channels
    Hf, Wf = cfg.feat_hw
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "torchvision": torchvision.__version__,
           "workload": cfg.name, "rois": cfg.n_rois, "iters": args.iters}

    # ---- ROIAlign
    out = ops.roi_align(feat, rois, scale, P, P, cfg.sampling_ratio, True)
    ours_f = timeit(lambda: ops.roi_align(feat, rois, scale, P, P, cfg.sampling_ratio, True), args.iters)
    ours_b = timeit(lambda: ops.roi_align_backward(out, rois, scale, P, P, N, C, Hf, Wf, cfg.sampling_ratio, True), args.iters)
    tv_f = timeit(lambda: tv_roi_align(feat, rois, (P, P), scale, cfg.sampling_ratio, True), args.iters)
    tv_b = timeit(lambda: torch.ops.torchvision._roi_align_backward(out, rois, scale, P, P, N, C, Hf, Wf,
                                                                    cfg.sampling_ratio, True), max(3, args.iters // 3))
    ref = tv_roi_align(feat, rois, (P, P), scale, cfg.sampling_ratio, True)
    res["roi_align"] = {"ours_fwd_ms": ours_f, "ours_bwd_ms": ours_b, "torchvision_cuda_fwd_ms": tv_f,
                        "torchvision_cuda_bwd_ms": tv_b, "fwd_speedup": tv_f / ours_f, "bwd_speedup": tv_b / ours_b,
                        "max_abs_diff_vs_torchvision_cuda": float((ref - out).abs().max())}
    del out, ref

    # ---- CLIP head + focal loss, fwd+bwd
    one = torch.ones(1, device=dev)

    def ours_head():
        ops.clip_head_loss(x, w, w_bg, gt, cfg.temperature, ops.LOSS_FOCAL, cfg.focal_gamma, cfg.bg_weight, one, False,
                           False, True)

    import torch.nn.functional as F

    def eager_head():
        xx = x.detach().requires_grad_(True)
        nx = F.normalize(xx, p=2.0, dim=1)
        s = torch.cat((nx @ F.normalize(w, p=2.0, dim=1).t(), F.linear(nx, w_bg)), dim=1) / cfg.temperature
        ce = F.cross_entropy(s, gt, reduction="none")
        p = F.softmax(s, dim=-1)
        pt = p[torch.arange(p.size(0), device=dev), gt]
        loss = ce * ((1 - pt) ** cfg.focal_gamma)
        lw = torch.ones(loss.size(0), device=dev)
        lw[gt == cfg.num_classes] = cfg.bg_weight
        (loss * lw).mean().backward()

    res["clip_head"] = {"ours_ms": timeit(ours_head, args.iters), "eager_ms": timeit(eager_head, args.iters)}

    # ---- alignment loss fwd+bwd (region level n = 16 * images)
    def ours_align():
        packed, norms = ops.align_pack(a_rs, a_rt)
        ops.align_loss(packed.unsqueeze(0), norms, 0, one, True)

    def eager_align():
        s_, t_ = a_rs.detach().requires_grad_(True), a_rt.detach().requires_grad_(True)
        a = s_ / s_.norm(dim=1, keepdim=True)
        b = t_ / t_.norm(dim=1, keepdim=True)
        j = a @ b.t()
        gt_ = torch.arange(len(j), device=dev)
        ((F.cross_entropy(j, gt_) + F.cross_entropy(j.t(), gt_)) / 2).backward()

    res["align_loss"] = {"n": a_rs.shape[0], "ours_ms": timeit(ours_align, args.iters), "eager_ms": timeit(eager_align, args.iters)}

    # ---- NMS sweep
    sweep = []
    for m, k in [(1000, 1), (4000, 1), (12000, 1), (12000, 20), (64000, 1), (262144, 1)]:
        gg = synth.generator(1000 + m + k)
        b, s, ids = [t.to(dev) for t in synth.make_nms_inputs(m, cfg.img_h, cfg.img_w, gg, num_classes=k)]
        ours = timeit(lambda: batched_nms(b, s, ids, 0.7), max(3, args.iters // 2))
        tv = timeit(lambda: tv_batched_nms(b, s, ids, 0.7), max(3, args.iters // 2))
        kept = int(batched_nms(b, s, ids, 0.7).numel())
        mask_bytes = 28 * m + 8 * m * ((m + 63) // 64)
        sweep.append({"boxes": m, "classes": k, "kept": kept, "ours_ms": ours, "torchvision_cuda_ms": tv,
                      "ours_boxes_per_s": m / (ours * 1e-3), "speedup": tv / ours,
                      "algorithmic_GBps": mask_bytes / (ours * 1e-3) / 1e9})
    res["nms"] = sweep

    # ---- the RPN call site: 16 images x 12000 pre-NMS proposals (proposal_utils.py:42-66), one level
    from cddmsl_b200.layers import batched_nms_images
    nb, m = cfg.n_images, 12000
    gg = synth.generator(777)
    per = [synth.make_nms_inputs(m, cfg.img_h, cfg.img_w, gg, num_classes=1) for _ in range(nb)]
    bb = torch.stack([p[0] for p in per]).to(dev)
    ss = torch.stack([p[1] for p in per]).to(dev)
    ii = torch.stack([p[2] for p in per]).to(dev)
    cnt = torch.full((nb,), m, dtype=torch.int32, device=dev)

    def ours_images():
        keep, nk = batched_nms_images(bb, ss, ii, cnt, 0.7)
        return nk.tolist()          # the one sync of the batch

    def ours_loop():
        return [batched_nms(bb[i], ss[i], ii[i], 0.7) for i in range(nb)]

    def tv_loop():
        return [tv_batched_nms(bb[i], ss[i], ii[i], 0.7) for i in range(nb)]

    res["nms_rpn_batch"] = {"images": nb, "boxes_per_image": m, "ours_one_call_ms": timeit(ours_images, 5),
                            "ours_per_image_loop_ms": timeit(ours_loop, 5), "torchvision_cuda_loop_ms": timeit(tv_loop, 3)}
    # ---- proposal labelling / sampling before ROIAlign (roi_heads.py:236-319): 16 images x 2000 proposals, 8 gt each
    import cddmsl_b200.modeling.roi_heads as rh
    from cddmsl_b200.structures import Boxes, Instances
    gg = synth.generator(778)
    props, tgts = [], []
    for _ in range(cfg.n_images):
        t = Instances((cfg.img_h, cfg.img_w))
        t.gt_boxes = Boxes(synth.make_boxes(8, cfg.img_h, cfg.img_w, gg, min_side=48.0, degenerate_frac=0.0).to(dev))
        t.gt_classes = torch.randint(0, cfg.num_classes, (8,), generator=gg).to(dev)
        p = Instances((cfg.img_h, cfg.img_w))
        p.proposal_boxes = Boxes(synth.make_boxes(2000, cfg.img_h, cfg.img_w, gg, degenerate_frac=0.0).to(dev))
        p.objectness_logits = torch.randn(2000, generator=gg).to(dev)
        props.append(p)
        tgts.append(t)
    heads = rh.ROIHeads(num_classes=cfg.num_classes)

    def lsp(batched):
        rh.BATCHED_IMAGES = batched
        try:
            return heads.label_and_sample_proposals(props, tgts)
        finally:
            rh.BATCHED_IMAGES = True

    res["label_and_sample_proposals"] = {"images": cfg.n_images, "proposals_per_image": 2000, "gt_per_image": 8,
                                         "ours_batched_ms": timeit(lambda: lsp(True), 10),
                                         "reference_shaped_loop_eager_ms": timeit(lambda: lsp(False), 5)}
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
