"""Generates the fixtures in this directory.  Run ONLY in the build container, where the upstream
checkout is mounted at /root/reference:

    python tests/golden/make_golden.py

What it records (all seeded, fp32 / int64):
  * roi_align_*.npz — outputs and input-gradients of the reference's own `ROIAlign` module
    (`detectron2/layers/roi_align.py`, loaded verbatim by file path) running the installed torchvision
    CPU op, on small maps with interior, over-hanging, degenerate and empty RoIs.
  * nms_*.npz       — kept indices of `detectron2/layers/nms.py:19-39` (restated in oracle/torch_ref.py;
    the file itself needs `detectron2.utils.env`) running the installed torchvision CPU `batched_nms`.
  * head_*.npz / align_*.npz — outputs of oracle/torch_ref.py (the reference has no test for these
    pieces: "parity unpinned"); `align_world2.npz` additionally comes from the reference's own
    `GatherLayer` (`detectron2/modeling/backbone/clipcap/gather.py`, loaded verbatim) under a real
    2-process gloo group, which pins the gather/backward semantics of `caption_consistency_world`.

  * box_reg.npz — `Box2BoxTransform.get_deltas` of the reference (`detectron2/modeling/box_regression.py`, loaded
    verbatim with its unused imports stubbed) + fvcore's published smooth-L1 (fvcore is not vendored).

  * match.npz — the reference's own `Matcher` (`detectron2/modeling/matcher.py`) on its own `pairwise_iou`
    (`detectron2/structures/boxes.py`), both loaded verbatim with two import stubs: ROI-head and RPN settings,
    an image without ground truth, duplicate gt boxes (argmax ties), degenerate proposals.

  * head_ref.npz — the reference's own `FastRCNNOutputLayers` (`detectron2/modeling/roi_heads/fast_rcnn.py`, loaded
    verbatim by file path with its package imports stubbed: fvcore.nn, detectron2.config.configurable (identity),
    detectron2.layers (ShapeSpec / cat / cross_entropy / nonzero_tuple from `layers/wrappers.py`), soft_nms,
    box_regression (the reference's own file), structures, utils.events (a recording storage)): `forward`,
    `focal_loss`, CE and weighted CE through `losses`, `_log_classification_stats`, gradients by autograd; zero and
    learned background row, one saturated row.  THIS pins pieces 3 (head) to the reference itself.
  * align_ref.npz — the literal source lines `detectron2/modeling/meta_arch/rcnn.py:455-470` (region level),
    `:305-319` (image level) and `:270-272` (KD L1), read from the file at generation time and executed on seeded
    tensors (single process with an identity gather, and under a real 2-process gloo group with the reference's
    `GatherLayer`).  THIS pins piece 4 to the reference itself.

  * pretrain_ref.npz — RegionCLIP pretraining losses: the literal lines `clip_rcnn.py:590-611`, `:624-640` and the
    literal class `MILCrossEntropy` (`utils/comm.py:332-355`) executed on seeded tensors.

  * inference_ref.npz — test-time post-processing: the reference's own `fast_rcnn_inference_single_image` on its own
    `Boxes` / `Instances` / `batched_nms` (all loaded verbatim, torchvision CPU underneath).

The tests never read /root/reference; they read these files.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from cddmsl_b200 import synth  # noqa: E402
from oracle import torch_ref  # noqa: E402


def load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod  # torch.jit.script (box_regression.py) inspects the defining module
    spec.loader.exec_module(mod)
    return mod


def roi_cases():
    ref = load_by_path("ref_roi_align", "detectron2/layers/roi_align.py")
    g = synth.generator(11)
    cases = {}
    for tag, (n, c, h, w, r, p, sr, aligned, scale) in {
        "p7_s0": (2, 6, 10, 16, 24, 7, 0, True, 1.0 / 16),
        "p14_s0": (2, 5, 12, 20, 24, 14, 0, True, 1.0 / 16),
        "p7_s2": (2, 4, 10, 16, 20, 7, 2, True, 1.0 / 16),
        "p5_legacy": (1, 3, 9, 11, 12, 5, 0, False, 0.25),
        "p14_big": (1, 3, 38, 63, 10, 14, 0, True, 1.0 / 16),
    }.items():
        img_h, img_w = int(h / scale), int(w / scale)
        feat = torch.randn(n, c, h, w, generator=g)
        parts = []
        per = r // n
        for i in range(n):
            b = synth.make_boxes(per, img_h, img_w, g, min_side=4.0, degenerate_frac=0.1)
            # over-hanging RoIs: shift a few outside the image (negative and beyond the far border)
            b[0] += torch.tensor([-0.6 * img_w, -0.3 * img_h, -0.2 * img_w, 0.0])
            b[1] += torch.tensor([0.5 * img_w, 0.4 * img_h, 0.9 * img_w, 0.8 * img_h])
            b[2] = torch.tensor([0.0, 0.0, float(img_w), float(img_h)])          # whole image
            b[3] = torch.tensor([3.0 * 4, 4.0 * 4, 5.0 * 4, 4.0 * 4])            # empty (zero height)
            b[4] = torch.tensor([0.7 * img_w, 0.7 * img_h, 0.4 * img_w, 0.5 * img_h])  # inverted
            parts.append(torch.cat([torch.full((per, 1), float(i)), b], 1))
        rois = torch.cat(parts, 0)
        op = ref.ROIAlign((p, p), scale, sr, aligned=aligned)
        x = feat.clone().requires_grad_(True)
        out = op(x, rois)
        gout = torch.randn(out.shape, generator=g)
        out.backward(gout)
        cases[tag] = dict(feat=feat.numpy(), rois=rois.numpy(), out=out.detach().numpy(), gout=gout.numpy(),
                          gin=x.grad.numpy(), meta=np.array([p, p, sr, int(aligned)], dtype=np.int64),
                          scale=np.array([scale], dtype=np.float64))
    for tag, d in cases.items():
        np.savez_compressed(os.path.join(HERE, f"roi_align_{tag}.npz"), **d)
    print("roi_align:", list(cases))


def nms_cases():
    g = synth.generator(12)
    out = {}
    for tag, (m, k, thrs, hw) in {
        "c1_m300": (300, 1, (0.2, 0.5, 0.7, 0.8), (200, 300)),
        "c7_m600_trick": (600, 7, (0.2, 0.5, 0.8), (200, 300)),     # numel 2400 <= 4000 -> coordinate trick
        "c5_m1500_vanilla": (1500, 5, (0.3, 0.7), (300, 400)),      # numel 6000 > 4000 -> per-class loop
        "rpn_m3000": (3000, 1, (0.7,), (600, 1000)),
    }.items():
        boxes, scores, idxs = synth.make_nms_inputs(m, hw[0], hw[1], g, num_classes=k, tie_frac=0.05)
        d = dict(boxes=boxes.numpy(), scores=scores.numpy(), idxs=idxs.numpy(),
                 thrs=np.array(thrs, dtype=np.float64))
        for t in thrs:
            d[f"keep_{t}"] = torch_ref.batched_nms(boxes, scores, idxs, t).numpy()
        out[tag] = d
    # reference test_nms.py:16-29 shape: random_boxes(2000, 200), 50 classes
    torch.manual_seed(121)
    n = 2000
    b = torch.rand(n, 4) * 100
    b.clamp_(min=1.0)
    b[:, 2:] += b[:, :2]
    s = torch.rand(n)
    ids = torch.randint(0, 50, (n,))
    d = dict(boxes=b.numpy(), scores=s.numpy(), idxs=ids.numpy(), thrs=np.array([0.2, 0.5, 0.8]))
    for t in (0.2, 0.5, 0.8):
        d[f"keep_{t}"] = torch_ref.batched_nms(b, s, ids, t).numpy()
    out["ref_test_n2000_c50"] = d
    for tag, d in out.items():
        np.savez_compressed(os.path.join(HERE, f"nms_{tag}.npz"), **d)
    print("nms:", list(out))


def head_cases():
    g = synth.generator(13)
    cfg = synth.CONFIGS["tiny"]
    r = 96
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=r)
    w_bg2 = torch.randn(1, cfg.emb_dim, generator=g) * 0.1   # a checkpoint may overwrite the zero bg row
    d = dict(x=x.numpy(), w=w.numpy(), gt=gt.numpy(), w_bg2=w_bg2.numpy(),
             params=np.array([cfg.temperature, cfg.focal_gamma, cfg.bg_weight], dtype=np.float64))
    for tag, wb in (("zero_bg", w_bg), ("learned_bg", w_bg2)):
        xx = x.clone().requires_grad_(True)
        scores = torch_ref.clip_head_scores(xx, w, wb, cfg.temperature)
        loss = torch_ref.focal_loss(scores, gt, cfg.num_classes, cfg.focal_gamma, cfg.bg_weight)
        loss.backward()
        d[f"scores_{tag}"] = scores.detach().numpy()
        d[f"loss_{tag}"] = loss.detach().numpy()
        d[f"dx_{tag}"] = xx.grad.numpy()
        d[f"stats_{tag}"] = np.array(torch_ref.classification_stats(scores.detach(), gt), dtype=np.int64)
        for lt, (gam, bw) in {"ce": (None, None), "wce": (None, cfg.bg_weight)}.items():
            xx = x.clone().requires_grad_(True)
            sc = torch_ref.clip_head_scores(xx, w, wb, cfg.temperature)
            ls = torch_ref.cls_loss(sc, gt, cfg.num_classes, gam, bw)
            ls.backward()
            d[f"loss_{lt}_{tag}"] = ls.detach().numpy()
            d[f"dx_{lt}_{tag}"] = xx.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "head_tiny.npz"), **d)
    print("head: tiny")


def _gather_worker(rank, world, a_locals, b_locals, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29591"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gl = load_by_path("ref_gather", "detectron2/modeling/backbone/clipcap/gather.py")
    a = a_locals[rank].clone().requires_grad_(True)
    b = b_locals[rank].clone().requires_grad_(True)
    a_all = torch.cat(gl.GatherLayer.apply(a), dim=0)      # rcnn.py:455-456
    b_all = torch.cat(gl.GatherLayer.apply(b), dim=0)
    loss = torch_ref.caption_consistency_loss(a_all, b_all)  # rcnn.py:458-468
    loss.backward()
    q.put((rank, loss.detach().numpy(), a.grad.numpy(), b.grad.numpy()))
    dist.destroy_process_group()


def align_cases():
    import torch.multiprocessing as mp

    g = synth.generator(14)
    d = {}
    for tag, (n, dim) in {"n16": (16, 256), "n40": (40, 64), "n256": (256, 256)}.items():
        a, b = torch.randn(n, dim, generator=g), torch.randn(n, dim, generator=g)
        aa, bb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        loss = torch_ref.caption_consistency_loss(aa, bb)
        loss.backward()
        d.update({f"a_{tag}": a.numpy(), f"b_{tag}": b.numpy(), f"loss_{tag}": loss.detach().numpy(),
                  f"da_{tag}": aa.grad.numpy(), f"db_{tag}": bb.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "align_single.npz"), **d)

    world, n_l, dim = 2, 12, 64
    a_locals = [torch.randn(n_l, dim, generator=g) for _ in range(world)]
    b_locals = [torch.randn(n_l, dim, generator=g) for _ in range(world)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, a_locals, b_locals, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join() for p in procs]
    emu_loss, emu_ga, emu_gb = torch_ref.caption_consistency_world(a_locals, b_locals)
    for r in range(world):   # the emulation must agree with the real GatherLayer before it is recorded
        assert np.allclose(res[r][1], emu_loss.numpy(), rtol=1e-6, atol=1e-7)
        assert np.allclose(res[r][2], emu_ga[r].numpy(), rtol=1e-5, atol=1e-8)
        assert np.allclose(res[r][3], emu_gb[r].numpy(), rtol=1e-5, atol=1e-8)
    w = {"loss": res[0][1]}
    for r in range(world):
        w.update({f"a{r}": a_locals[r].numpy(), f"b{r}": b_locals[r].numpy(), f"da{r}": res[r][2],
                  f"db{r}": res[r][3]})
    np.savez_compressed(os.path.join(HERE, "align_world2.npz"), **w)
    print("align: single + world2 (real GatherLayer under gloo agrees with the emulation)")


def box_reg_cases():
    """get_deltas comes from the reference's own Box2BoxTransform (detectron2/modeling/box_regression.py, loaded
    verbatim; its imports of fvcore / detectron2.layers / detectron2.structures are stubbed, get_deltas uses none of
    them); the smooth-L1 reduction is fvcore's published definition (oracle/torch_ref.py:smooth_l1_sum)."""
    import types

    stubs = {}
    for name, attrs in {"fvcore": {}, "fvcore.nn": {"giou_loss": None, "smooth_l1_loss": None},
                        "detectron2": {}, "detectron2.layers": {"cat": torch.cat},
                        "detectron2.structures": {"Boxes": object}}.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            stubs[name] = m
    try:
        ref = load_by_path("ref_box_regression", "detectron2/modeling/box_regression.py")
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
    weights = (10.0, 10.0, 5.0, 5.0)
    b2b = ref.Box2BoxTransform(weights=weights)
    g = synth.generator(15)
    r, k = 300, 7
    prop = synth.make_boxes(r, 600, 1000, g, degenerate_frac=0.0)
    gtb = synth.make_boxes(r, 600, 1000, g, degenerate_frac=0.0)
    gt = torch.randint(0, k + 1, (r,), generator=g)
    bg = (gt == k).nonzero()[:6, 0]
    prop[bg] = torch.tensor([5.0, 5.0, 5.0, 5.0])                    # degenerate background proposals: never evaluated
    fg = ((gt >= 0) & (gt < k)).nonzero(as_tuple=True)[0]
    ref_deltas = b2b.get_deltas(prop[fg], gtb[fg])
    out = dict(prop=prop.numpy(), gtb=gtb.numpy(), gt=gt.numpy(), fg=fg.numpy(), ref_deltas=ref_deltas.numpy(),
               weights=np.array(weights, dtype=np.float64), k=np.array([k]))
    for tag, width, beta in (("agnostic", 4, 0.5), ("perclass", 4 * k, 0.5), ("l1", 4, 0.0)):
        pred = torch.randn(r, width, generator=g).requires_grad_(True)
        fg_pred = pred[fg] if width == 4 else pred.view(-1, k, 4)[fg, gt[fg]]
        loss = torch_ref.smooth_l1_sum(fg_pred, ref_deltas, beta) / max(r, 1.0)
        loss.backward()
        out[f"pred_{tag}"] = pred.detach().numpy()
        out[f"beta_{tag}"] = np.array([beta])
        out[f"loss_{tag}"] = np.array([loss.item()], dtype=np.float32)
        out[f"dpred_{tag}"] = pred.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "box_reg.npz"), **out)
    print("box_reg: agnostic perclass l1")


def _load_with_stubs(name, rel, stubs):
    import types

    added = {}
    for mod_name, attrs in stubs.items():
        if mod_name not in sys.modules:
            m = types.ModuleType(mod_name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[mod_name] = m
            added[mod_name] = m
    try:
        return load_by_path(name, rel)
    finally:
        for mod_name in added:
            sys.modules.pop(mod_name, None)


def match_cases():
    """Matcher + pairwise_iou come from the reference's own files (detectron2/modeling/matcher.py and
    detectron2/structures/boxes.py, loaded verbatim; `detectron2.layers.nonzero_tuple` and
    `detectron2.utils.env.TORCH_VERSION` stubbed)."""
    def nonzero_tuple(x):
        return x.nonzero().unbind(1) if x.dim() else x.unsqueeze(0).nonzero().unbind(1)

    stubs = {"detectron2": {}, "detectron2.layers": {"nonzero_tuple": nonzero_tuple, "cat": torch.cat},
             "detectron2.utils": {}, "detectron2.utils.env": {"TORCH_VERSION": (2, 0)}}
    ref_m = _load_with_stubs("ref_matcher", "detectron2/modeling/matcher.py", stubs)
    ref_b = _load_with_stubs("ref_boxes", "detectron2/structures/boxes.py", stubs)
    g = synth.generator(16)
    out = {}
    g_len, m_len = [7, 0, 3, 40], [300, 200, 257, 1000]
    for b, (ng, nm) in enumerate(zip(g_len, m_len)):
        gt = synth.make_boxes(ng, 600, 1000, g, min_side=32.0, degenerate_frac=0.0) if ng else torch.zeros(0, 4)
        if ng >= 3:
            gt[2] = gt[0]                                   # duplicate gt: argmax tie -> first index
        props = synth.make_boxes(nm, 600, 1000, g, degenerate_frac=0.03)
        if ng:
            jit = gt[torch.randint(0, ng, (nm // 3,), generator=g)] + torch.randn(nm // 3, 4, generator=g) * 6.0
            props[: nm // 3] = jit                          # a third of the proposals hug a gt box
            props = torch.cat([props, gt])                  # proposal_append_gt
        out[f"gt_{b}"], out[f"boxes_{b}"] = gt.numpy(), props.numpy()
        mqm = ref_b.pairwise_iou(ref_b.Boxes(gt), ref_b.Boxes(props))
        out[f"iou_max_{b}"] = (mqm.max(dim=0).values if ng else torch.zeros(len(props))).numpy()
        for tag, (thr, lab, low) in {"roi": ([0.5], [0, 1], False), "rpn": ([0.3, 0.7], [0, -1, 1], True)}.items():
            matches, labels = ref_m.Matcher(thr, lab, allow_low_quality_matches=low)(mqm)
            out[f"matches_{tag}_{b}"], out[f"labels_{tag}_{b}"] = matches.numpy(), labels.numpy()
    out["n_images"] = np.array([len(g_len)])
    np.savez_compressed(os.path.join(HERE, "match.npz"), **out)
    print("match:", g_len, m_len)


class _RecordingStorage:
    def __init__(self):
        self.scalars = {}

    def put_scalar(self, name, value, **kw):
        self.scalars[name] = float(value)


def _load_ref_fast_rcnn(storage):
    """detectron2/modeling/roi_heads/fast_rcnn.py, verbatim, with its package imports stubbed."""
    import collections
    import types

    wr = _load_with_stubs("ref_wrappers", "detectron2/layers/wrappers.py", {})
    br = _load_with_stubs("ref_box_regression2", "detectron2/modeling/box_regression.py", {
        "fvcore": {}, "fvcore.nn": {"giou_loss": None, "smooth_l1_loss": None}, "detectron2": {},
        "detectron2.layers": {"cat": torch.cat}, "detectron2.structures": {"Boxes": object}})
    ShapeSpec = collections.namedtuple("ShapeSpec", ["channels", "height", "width", "stride"],
                                       defaults=(None, None, None, None))

    def configurable(f=None, **kw):   # @configurable with explicit keyword arguments == the plain constructor
        return f

    def smooth_l1_loss(inp, tgt, beta, reduction="none"):   # fvcore's published definition
        n = torch.abs(inp - tgt)
        loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
        return loss.sum() if reduction == "sum" else loss.mean() if reduction == "mean" else loss

    stubs = {
        "fvcore": {}, "fvcore.nn": {"giou_loss": None, "smooth_l1_loss": smooth_l1_loss},
        "detectron2": {}, "detectron2.config": {"configurable": configurable},
        "detectron2.layers": {"ShapeSpec": ShapeSpec, "batched_nms": None, "cat": wr.cat,
                              "cross_entropy": wr.cross_entropy, "nonzero_tuple": wr.nonzero_tuple},
        "detectron2.layers.soft_nms": {"batched_soft_nms": None},
        "detectron2.modeling": {}, "detectron2.modeling.box_regression": {"Box2BoxTransform": br.Box2BoxTransform},
        "detectron2.structures": {"Boxes": object, "Instances": object},
        "detectron2.utils": {}, "detectron2.utils.events": {"get_event_storage": lambda: storage},
    }
    return _load_with_stubs("ref_fast_rcnn", "detectron2/modeling/roi_heads/fast_rcnn.py", stubs), br, ShapeSpec


class _Props:   # the two fields `losses` reads from a proposal `Instances`
    def __init__(self, boxes, gt):
        self.proposal_boxes = type("B", (), {"tensor": boxes})()
        self.gt_classes = gt

    def has(self, name):
        return False


def head_ref_cases():
    import tempfile

    storage = _RecordingStorage()
    ref, br, ShapeSpec = _load_ref_fast_rcnn(storage)
    g = synth.generator(23)
    cfg = synth.CONFIGS["tiny"]
    r, k, d_emb = 80, cfg.num_classes, cfg.emb_dim
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=r)
    x[5] = 6.0 * w[int(gt[5]) if int(gt[5]) < k else 0]      # one saturated row (softmax p_t == 1 in fp32)
    gt[5] = gt[5] if int(gt[5]) < k else 0
    w_bg2 = torch.randn(1, d_emb, generator=g) * 0.1
    boxes = synth.make_boxes(r, 600, 1000, g, degenerate_frac=0.0)
    out = dict(x=x.numpy(), w=w.numpy(), gt=gt.numpy(), w_bg2=w_bg2.numpy(),
               params=np.array([cfg.temperature, cfg.focal_gamma, cfg.bg_weight], dtype=np.float64))
    with tempfile.TemporaryDirectory() as td:
        emb = os.path.join(td, "emb.pth")
        torch.save(w.clone(), emb)
        for mode, focal, bgw in (("focal", cfg.focal_gamma, cfg.bg_weight), ("ce", None, None),
                                 ("wce", None, cfg.bg_weight)):
            for tag, wb in (("zero_bg", None), ("learned_bg", w_bg2)):
                head = ref.FastRCNNOutputLayers(
                    ShapeSpec(channels=d_emb), box2box_transform=br.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)),
                    num_classes=k, clip_cls_emb=(True, emb, "CLIPRes5ROIHeads", d_emb), bg_cls_loss_weight=bgw,
                    openset_test=(None, None, cfg.temperature, focal))
                head.train()
                if wb is not None:
                    with torch.no_grad():
                        head.cls_bg_score.weight.copy_(wb)
                xx = x.clone().requires_grad_(True)
                scores, deltas = head(xx)                                       # fast_rcnn.py:529-572
                storage.scalars.clear()
                losses = head.losses((scores, deltas.detach()), [_Props(boxes, gt)])   # :574-622 (+ :624-644, :100-127)
                loss = losses["loss_cls"]
                loss.backward()
                key = f"{mode}_{tag}"
                out[f"scores_{tag}"] = scores.detach().numpy()
                out[f"loss_{key}"] = loss.detach().numpy()
                out[f"dx_{key}"] = xx.grad.numpy()            # NaN in the saturated row for the focal loss (:630-633)
                out[f"stats_{tag}"] = np.array([storage.scalars.get("fast_rcnn/cls_accuracy", -1.0),
                                                storage.scalars.get("fast_rcnn/fg_cls_accuracy", -1.0),
                                                storage.scalars.get("fast_rcnn/false_negative", -1.0)])
    np.savez_compressed(os.path.join(HERE, "head_ref.npz"), **out)
    print("head_ref:", sorted(k for k in out if k.startswith("loss_")))


def inference_ref_cases():
    """inference_ref.npz: the reference's OWN `fast_rcnn_inference_single_image` (fast_rcnn.py:101-209, loaded verbatim)
    running on its own `Boxes` / `Instances` (structures/boxes.py, instances.py, verbatim) and its own `batched_nms`
    (layers/nms.py, verbatim, over torchvision CPU): class-specific and class-agnostic boxes, non-finite rows, a
    threshold nothing passes, `vis` scores, no top-k limit."""
    def nonzero_tuple(x):
        return x.nonzero().unbind(1) if x.dim() else x.unsqueeze(0).nonzero().unbind(1)

    base = {"detectron2": {}, "detectron2.layers": {"nonzero_tuple": nonzero_tuple, "cat": torch.cat},
            "detectron2.utils": {}, "detectron2.utils.env": {"TORCH_VERSION": (2, 0)}}
    ref_b = _load_with_stubs("ref_boxes_inf", "detectron2/structures/boxes.py", base)
    ref_i = _load_with_stubs("ref_instances_inf", "detectron2/structures/instances.py", {})
    import types
    # (nms.py binds the rotated-box op at import time: give it the pre-1.7 branch with a placeholder extension module)
    ref_n = _load_with_stubs("ref_nms_inf", "detectron2/layers/nms.py", {
        "detectron2": {"_C": types.SimpleNamespace(nms_rotated=None)}, "detectron2.utils": {},
        "detectron2.utils.env": {"TORCH_VERSION": (1, 6)}})
    storage = _RecordingStorage()
    ref, _, _ = _load_ref_fast_rcnn(storage)
    ref.Boxes, ref.Instances, ref.batched_nms = ref_b.Boxes, ref_i.Instances, ref_n.batched_nms
    g = synth.generator(41)
    shape = (600, 1000)
    cases = {  # name: (R, K, per-class boxes, score_thresh, nms_thresh, topk, vis, poison rows)
        "perclass": (200, 20, True, 0.05, 0.5, 100, False, 0),
        "agnostic": (300, 8, False, 0.3, 0.5, 10, False, 0),
        "nonfinite": (120, 6, True, 0.05, 0.6, 50, False, 5),
        "nothing": (64, 5, True, 2.0, 0.5, 100, False, 0),
        "vis": (150, 10, False, 0.05, 0.4, 30, True, 0),
        "nolimit": (1000, 3, True, 0.2, 0.7, -1, False, 0),
    }
    out = {"names": np.array(list(cases))}
    for name, (r, k, per_class, st, nt, topk, vis, poison) in cases.items():
        centre = synth.make_boxes(r, shape[0], shape[1], g, degenerate_frac=0.0)
        if per_class:
            boxes = (centre[:, None, :] + torch.randn(r, k, 4, generator=g) * 8.0).reshape(r, k * 4)
        else:
            boxes = centre.clone()
        boxes[: r // 4] = boxes[r // 4: 2 * (r // 4)] + torch.randn_like(boxes[: r // 4]) * 3.0   # overlapping groups
        scores = torch.softmax(torch.randn(r, k + 1, generator=g) * 2.5, dim=1)
        sbf = torch.softmax(torch.randn(r, k + 1, generator=g) * 2.5, dim=1)
        if poison:
            boxes[3, 1] = float("nan")
            boxes[17, 0] = float("inf")
            scores[29, 2] = float("nan")
            scores[44, 0] = float("inf")
            boxes[60, 3] = float("-inf")
        res, kept = ref.fast_rcnn_inference_single_image(boxes.clone(), scores.clone(), shape, st, nt, False, "gaussian",
                                                         0.5, 0.001, topk, sbf.clone(), vis)
        out.update({f"boxes_{name}": boxes.numpy(), f"scores_{name}": scores.numpy(), f"sbf_{name}": sbf.numpy(),
                    f"params_{name}": np.array([st, nt, topk, float(vis)]),
                    f"pred_boxes_{name}": res.pred_boxes.tensor.numpy(), f"pred_scores_{name}": res.scores.numpy(),
                    f"pred_classes_{name}": res.pred_classes.numpy(), f"kept_{name}": kept.numpy()})
        print("inference_ref", name, "detections", len(kept))
    np.savez_compressed(os.path.join(HERE, "inference_ref.npz"), **out)


def _ref_lines(rel, first, last):
    with open(os.path.join(REF, rel)) as fh:
        lines = fh.read().splitlines()[first - 1:last]
    import textwrap
    return textwrap.dedent("\n".join(lines))


def _ref_align_fns():
    """The literal alignment-loss source of detectron2/modeling/meta_arch/rcnn.py wrapped into functions."""
    import textwrap

    from torch import nn
    region = "def region(self, src_features, target_features, GatherLayer):\n" + textwrap.indent(
        _ref_lines("detectron2/modeling/meta_arch/rcnn.py", 455, 470), "    ")
    image = "def image(self, student_features_trgt, student_features_src, kd_loss, GatherLayer):\n" + textwrap.indent(
        _ref_lines("detectron2/modeling/meta_arch/rcnn.py", 305, 319), "    ")
    kd = "def kd(teacher_features, student_features_src):\n" + textwrap.indent(
        _ref_lines("detectron2/modeling/meta_arch/rcnn.py", 270, 272), "    ") + "\n    return kd_loss\n"
    ns = {"torch": torch, "nn": nn}
    for src in (region, image, kd):
        exec(src, ns)
    return ns["region"], ns["image"], ns["kd"]


class _Self:
    device = torch.device("cpu")


class _IdentityGather:
    @staticmethod
    def apply(x):
        return (x,)


def _align_ref_worker(rank, world, a_locals, b_locals, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "29593"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gl = load_by_path("ref_gather", "detectron2/modeling/backbone/clipcap/gather.py")
    region, image, _ = _ref_align_fns()
    res = []
    for fn, extra in ((region, ()), (image, (None,))):
        a = a_locals[rank].clone().requires_grad_(True)
        b = b_locals[rank].clone().requires_grad_(True)
        loss = fn(_Self(), a, b, *extra, gl.GatherLayer)
        loss = loss[0] if isinstance(loss, tuple) else loss
        loss.backward()
        res.append((loss.detach().numpy(), a.grad.numpy(), b.grad.numpy()))
    q.put((rank, res))
    dist.destroy_process_group()


def align_ref_cases():
    import torch.multiprocessing as mp

    region, image, kd = _ref_align_fns()
    g = synth.generator(24)
    d = {}
    for tag, (n, dim) in {"n16": (16, 256), "n48": (48, 96), "n256": (256, 256)}.items():
        a, b = torch.randn(n, dim, generator=g), torch.randn(n, dim, generator=g)
        for kind, fn, extra in (("region", region, ()), ("image", image, (None,))):
            aa, bb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
            loss = fn(_Self(), aa, bb, *extra, _IdentityGather)
            loss = loss[0] if isinstance(loss, tuple) else loss
            loss.backward()
            d.update({f"a_{tag}": a.numpy(), f"b_{tag}": b.numpy(), f"{kind}_loss_{tag}": loss.detach().numpy(),
                      f"{kind}_da_{tag}": aa.grad.numpy(), f"{kind}_db_{tag}": bb.grad.numpy()})
    # KD regulariser (rcnn.py:265-272): L1 between the frozen teacher's and the student's V2L features [B, 768]
    t, s_ = torch.randn(8, 768, generator=g), torch.randn(8, 768, generator=g)
    ss = s_.clone().requires_grad_(True)
    kl = kd(t, ss)
    kl.backward()
    d.update(kd_teacher=t.numpy(), kd_student=s_.numpy(), kd_loss=kl.detach().numpy(), kd_dstudent=ss.grad.numpy())
    # two ranks, the reference's own GatherLayer
    world, n_l, dim = 2, 10, 64
    a_locals = [torch.randn(n_l, dim, generator=g) for _ in range(world)]
    b_locals = [torch.randn(n_l, dim, generator=g) for _ in range(world)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_align_ref_worker, args=(r, world, a_locals, b_locals, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=180) for _ in range(world))
    [p.join() for p in procs]
    for r in range(world):
        d.update({f"w2_a{r}": a_locals[r].numpy(), f"w2_b{r}": b_locals[r].numpy()})
        for ki, kind in enumerate(("region", "image")):
            d.update({f"w2_{kind}_loss": res[r][ki][0], f"w2_{kind}_da{r}": res[r][ki][1],
                      f"w2_{kind}_db{r}": res[r][ki][2]})
    np.savez_compressed(os.path.join(HERE, "align_ref.npz"), **d)
    print("align_ref: region/image single + world2, kd")


def pretrain_ref_cases():
    """pretrain_ref.npz: the literal lines detectron2/modeling/meta_arch/clip_rcnn.py:590-611 (region-concept KL
    distillation + MIL contrastive loss) and :624-640 (image-text matching, single process) with the literal class
    `MILCrossEntropy` of detectron2/utils/comm.py:332-355, executed on seeded tensors; gradients by autograd."""
    import textwrap

    import torch.nn.functional as F
    from torch import nn

    ns = {"torch": torch, "nn": nn, "F": F}
    exec(_ref_lines("detectron2/utils/comm.py", 332, 355), ns)
    body = _ref_lines("detectron2/modeling/meta_arch/clip_rcnn.py", 590, 611)
    exec("def region_concept(self, keep_region_feats, concept_scores, target_embs, label_mtx, losses, "
         "use_distill=True, use_contrastive=True):\n" + textwrap.indent(body, "    "), ns)
    body = _ref_lines("detectron2/modeling/meta_arch/clip_rcnn.py", 624, 640)
    exec("def image_text(self, global_feats, text_embs, losses):\n" + textwrap.indent(body, "    "), ns)
    g = synth.generator(31)
    out = {}
    for tag, (r, d, k, temp) in {"small": (48, 64, 30, 0.01), "lvis": (40, 128, 1203, 0.01)}.items():
        feats = torch.randn(r, d, generator=g)
        concept_emb = torch.randn(k, d, generator=g)
        teacher = torch.softmax(torch.randn(r, k, generator=g) * 3.0, dim=1)
        teacher[0, :5] = 0.0                                     # exact zeros in the target (0 log 0 = 0)
        tgt_idx = torch.randint(0, 12, (r,), generator=g)        # pseudo concept of every kept region
        target_embs = concept_emb[tgt_idx]
        label_mtx = (tgt_idx[:, None] == tgt_idx[None, :]).float()

        class S:
            pass

        self_ = S()
        self_.concept_emb, self_.matching_temp, self_.gather_gpus, self_.device = concept_emb, temp, False, torch.device("cpu")
        x = feats.clone().requires_grad_(True)
        losses = {}
        ns["region_concept"](self_, x, teacher, target_embs, label_mtx, losses)
        gd, = torch.autograd.grad(losses["loss_region_distill"], x, retain_graph=True)
        gc, = torch.autograd.grad(losses["loss_concept_contrastive"], x)
        out.update({f"feats_{tag}": feats.numpy(), f"concept_emb_{tag}": concept_emb.numpy(),
                    f"teacher_{tag}": teacher.numpy(), f"target_embs_{tag}": target_embs.numpy(),
                    f"label_mtx_{tag}": label_mtx.numpy(), f"temp_{tag}": np.array([temp]),
                    f"distill_{tag}": losses["loss_region_distill"].detach().numpy(), f"distill_dx_{tag}": gd.numpy(),
                    f"contrastive_{tag}": losses["loss_concept_contrastive"].detach().numpy(),
                    f"contrastive_dx_{tag}": gc.numpy()})
    n, d, temp = 24, 96, 0.07
    a, b = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)

    class S2:
        matching_temp, gather_gpus, device = temp, False, torch.device("cpu")

    aa, bb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    losses = {}
    ns["image_text"](S2(), aa, bb, losses)
    losses["loss_img_txt_level"].backward()
    out.update(it_feats=a.numpy(), it_text=b.numpy(), it_temp=np.array([temp]),
               it_loss=losses["loss_img_txt_level"].detach().numpy(), it_dfeats=aa.grad.numpy(), it_dtext=bb.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "pretrain_ref.npz"), **out)
    print("pretrain_ref: region-concept small/lvis, image-text")


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    if len(sys.argv) > 1 and sys.argv[1] in ("box_reg", "match", "head_ref", "align_ref", "pretrain_ref",
                                             "inference_ref"):
        {"box_reg": box_reg_cases, "match": match_cases, "head_ref": head_ref_cases,
         "align_ref": align_ref_cases, "pretrain_ref": pretrain_ref_cases,
         "inference_ref": inference_ref_cases}[sys.argv[1]]()
        sys.exit(0)
    roi_cases()
    nms_cases()
    head_cases()
    align_cases()
    box_reg_cases()
    match_cases()
    head_ref_cases()
    align_ref_cases()
    pretrain_ref_cases()
    inference_ref_cases()
    sizes = {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")}
    print(sizes, sum(sizes.values()))
