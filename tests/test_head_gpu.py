"""GPU parity: CLIP box-predictor head (cosine logits, focal / CE / weighted-CE loss, dx, statistics).
The reference holds no test for this piece (parity unpinned upstream); the oracle is the restatement of
fast_rcnn.py:543-565, :624-644 in oracle/torch_ref.py.  Tolerance: 1e-5 relative forward, 1e-4 backward."""
import os

import numpy as np
import pytest
import torch

from cddmsl_b200 import synth
from oracle import torch_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, want, rtol, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(float(np.abs(want).max()), 1e-12)
    err = np.abs(got - want)
    ok = err <= rtol * np.abs(want) + rtol * scale
    assert ok.all(), f"{what}: max abs err {err.max():.3e} (scale {scale:.3e}), {(~ok).sum()}/{ok.size} outside rtol={rtol}"


def _predictor(k, d, w, w_bg=None, gamma=0.5, bgw=0.2, temperature=0.01, strict=False):
    from cddmsl_b200.modeling import Box2BoxTransform, FastRCNNOutputLayers

    m = FastRCNNOutputLayers(d, box2box_transform=Box2BoxTransform((10.0, 10.0, 5.0, 5.0)), num_classes=k,
                             clip_cls_emb=(True, w, "CLIPRes5ROIHeads", d), bg_cls_loss_weight=bgw,
                             openset_test=(None, None, temperature, gamma), strict_focal_nan=strict)
    if w_bg is not None:
        with torch.no_grad():
            m.cls_bg_score.weight.copy_(w_bg)
    return m.to(DEV).train()


def _proposals(gt, boxes=None):
    from cddmsl_b200.structures import Boxes, Instances

    r = len(gt)
    if boxes is None:
        boxes = torch.tensor([[10.0, 10.0, 60.0, 80.0]]).repeat(r, 1)
    inst = Instances((600, 1000))
    inst.proposal_boxes = Boxes(boxes.to(DEV))
    inst.gt_boxes = Boxes((boxes + 3.0).to(DEV))
    inst.gt_classes = gt.to(DEV)
    return [inst]


def test_fixture_tiny_all_modes(golden_dir):
    d = np.load(os.path.join(golden_dir, "head_tiny.npz"))
    T, gamma, bgw = (float(v) for v in d["params"])
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    k, dim = w.shape
    for tag, wb in (("zero_bg", None), ("learned_bg", torch.from_numpy(d["w_bg2"]))):
        for mode, (gam, bw) in {"": (gamma, bgw), "ce_": (None, None), "wce_": (None, bgw)}.items():
            m = _predictor(k, dim, w, wb, gamma=gam, bgw=bw, temperature=T, strict=True)
            xx = x.to(DEV).requires_grad_(True)
            preds = m(xx)
            losses = m.losses(preds, _proposals(gt))
            losses["loss_cls"].backward()
            if mode == "":
                _close(preds[0].detach().cpu().numpy(), d[f"scores_{tag}"], 1e-5, "scores")
            _close(losses["loss_cls"].item(), d[f"loss_{mode}{tag}"], 1e-5, f"loss {mode}{tag}")
            want = d[f"dx_{mode}{tag}"]
            got = xx.grad.cpu().numpy()
            nan_rows = np.isnan(want).any(axis=1)
            assert np.array_equal(np.isnan(got).any(axis=1), nan_rows), "strict mode must reproduce autograd's NaN rows"
            _close(got[~nan_rows], want[~nan_rows], 1e-4, f"dx {mode}{tag}")


def test_reference_fast_rcnn_fixture_all_modes(golden_dir):
    """head_ref.npz: outputs of the reference's OWN FastRCNNOutputLayers.forward / losses / focal_loss /
    _log_classification_stats (fast_rcnn.py loaded verbatim, tests/golden/make_golden.py:head_ref_cases)."""
    from cddmsl_b200.modeling import fast_rcnn as fr

    d = np.load(os.path.join(golden_dir, "head_ref.npz"))
    T, gamma, bgw = (float(v) for v in d["params"])
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    k, dim = w.shape
    for tag, wb in (("zero_bg", None), ("learned_bg", torch.from_numpy(d["w_bg2"]))):
        for mode, (gam, bw) in {"focal": (gamma, bgw), "ce": (None, None), "wce": (None, bgw)}.items():
            m = _predictor(k, dim, w, wb, gamma=gam, bgw=bw, temperature=T, strict=True)
            xx = x.to(DEV).requires_grad_(True)
            seen = {}
            fr.set_scalar_sink(lambda kk, v: seen.__setitem__(kk, v))
            try:
                preds = m(xx)
                losses = m.losses(preds, _proposals(gt))
            finally:
                fr.set_scalar_sink(None)
            losses["loss_cls"].backward()
            _close(preds[0].detach().cpu().numpy(), d[f"scores_{tag}"], 1e-5, "scores")
            _close(losses["loss_cls"].item(), d[f"loss_{mode}_{tag}"], 1e-5, f"loss {mode} {tag}")
            want = d[f"dx_{mode}_{tag}"]
            got = xx.grad.cpu().numpy()
            nan_rows = np.isnan(want).any(axis=1)
            assert np.array_equal(np.isnan(got).any(axis=1), nan_rows), "strict mode reproduces the reference's NaN rows"
            _close(got[~nan_rows], want[~nan_rows], 1e-4, f"dx {mode} {tag}")
            # per-element relative check where the value is not tiny (tolerance note)
            big = np.abs(want) > 1e-2 * np.abs(want[~nan_rows]).max()
            big &= ~nan_rows[:, None]
            assert (np.abs(got[big] - want[big]) <= 1e-4 * np.abs(want[big])).all()
            acc, fgacc, fn = d[f"stats_{tag}"]
            assert abs(seen["fast_rcnn/cls_accuracy"] - acc) < 1e-12
            assert abs(seen["fast_rcnn/fg_cls_accuracy"] - fgacc) < 1e-12
            assert abs(seen["fast_rcnn/false_negative"] - fn) < 1e-12


def test_saturated_row_default_is_analytic_limit(golden_dir):
    d = np.load(os.path.join(golden_dir, "head_tiny.npz"))
    T, gamma, bgw = (float(v) for v in d["params"])
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    want = d["dx_learned_bg"]
    nan_rows = np.isnan(want).any(axis=1)
    assert nan_rows.sum() >= 1   # the fixture holds an adversarial saturated row
    m = _predictor(w.shape[0], w.shape[1], w, torch.from_numpy(d["w_bg2"]), gamma=gamma, bgw=bgw, temperature=T)
    xx = x.to(DEV).requires_grad_(True)
    m.losses(m(xx), _proposals(gt))["loss_cls"].backward()
    got = xx.grad.cpu().numpy()
    assert np.isfinite(got).all() and np.abs(got[nan_rows]).max() == 0.0
    _close(got[~nan_rows], want[~nan_rows], 1e-4)


def test_stats_and_scalar_names(golden_dir):
    from cddmsl_b200.modeling import fast_rcnn as fr

    d = np.load(os.path.join(golden_dir, "head_tiny.npz"))
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    seen = {}
    fr.set_scalar_sink(lambda k, v: seen.__setitem__(k, v))
    try:
        m = _predictor(w.shape[0], w.shape[1], w, temperature=float(d["params"][0]))
        m.losses(m(x.to(DEV)), _proposals(gt))
    finally:
        fr.set_scalar_sink(None)
    acc, nfg, fgacc, fn = (int(v) for v in d["stats_zero_bg"])
    assert seen["fast_rcnn/cls_accuracy"] == acc / len(gt)
    assert seen["fast_rcnn/fg_cls_accuracy"] == fgacc / nfg
    assert seen["fast_rcnn/false_negative"] == fn / nfg


@pytest.mark.parametrize("cfg_name", ["cpu_ref", "voc", "city"])
def test_config_shapes_vs_oracle(cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    g = synth.generator(cfg.seed)
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g)
    xr = x.clone().requires_grad_(True)
    s_ref = torch_ref.clip_head_scores(xr, w, w_bg, cfg.temperature)
    l_ref = torch_ref.focal_loss(s_ref, gt, cfg.num_classes, cfg.focal_gamma, cfg.bg_weight)
    l_ref.backward()
    m = _predictor(cfg.num_classes, cfg.emb_dim, w, gamma=cfg.focal_gamma, bgw=cfg.bg_weight, temperature=cfg.temperature)
    xx = x.to(DEV).requires_grad_(True)
    preds = m(xx)
    loss = m.losses(preds, _proposals(gt))["loss_cls"]
    loss.backward()
    _close(preds[0].detach().cpu().numpy(), s_ref.detach().numpy(), 1e-5, "scores")
    _close(loss.item(), l_ref.item(), 1e-5, "loss")
    _close(xx.grad.cpu().numpy(), xr.grad.numpy(), 1e-4, "dx")


def test_unfused_path_scores_autograd():
    """scores that did not come straight out of forward() (here: a re-created tensor) take the two-op path:
    our cosine-logit op (with its own backward kernel) + PyTorch loss arithmetic."""
    cfg = synth.CONFIGS["tiny"]
    g = synth.generator(3)
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=200)
    xr = x.clone().requires_grad_(True)
    l_ref = torch_ref.focal_loss(torch_ref.clip_head_scores(xr, w, w_bg, 0.05), gt, cfg.num_classes, 0.5, 0.2)
    l_ref.backward()
    m = _predictor(cfg.num_classes, cfg.emb_dim, w, temperature=0.05)
    xx = x.to(DEV).requires_grad_(True)
    scores, deltas = m(xx)
    loss = m.losses((scores * 1.0, deltas), _proposals(gt))["loss_cls"]
    loss.backward()
    _close(loss.item(), l_ref.item(), 1e-5)
    _close(xx.grad.cpu().numpy(), xr.grad.numpy(), 1e-4)


@pytest.mark.parametrize("path,r,dim,k", [("cuda_cores", 2048, 1024, 1203), ("tcgen05", 2048, 1024, 1203),
                                          ("tcgen05", 130, 64, 259), ("tcgen05", 700, 96, 1203)])
def test_large_vocabulary(path, r, dim, k):
    """LVIS-scale vocabulary (BASELINE.json configs[3]: 1203 concepts) at a row count the CPU oracle finishes
    quickly, through both implementations: the generic warp-per-row kernel and the tcgen05 3xTF32 GEMM path
    (ragged M, N and a K that is not a multiple of the tile)."""
    from cddmsl_b200 import _lib

    _lib.tune("head_tc", 0 if path == "cuda_cores" else 1)
    try:
        _large_vocab_case(r, dim, k)
    finally:
        _lib.tune("head_tc", 2)


def _large_vocab_case(r, dim, k):
    g = synth.generator(3)
    x = torch.randn(r, dim, generator=g)
    w = torch.randn(k, dim, generator=g)
    gt = torch.randint(0, k + 1, (r,), generator=g)
    xr = x.clone().requires_grad_(True)
    s_ref = torch_ref.clip_head_scores(xr, w, torch.zeros(1, dim), 0.01)
    l_ref = torch_ref.focal_loss(s_ref, gt, k, 0.5, 0.2)
    l_ref.backward()
    m = _predictor(k, dim, w)
    xx = x.to(DEV).requires_grad_(True)
    preds = m(xx)
    loss = m.losses(preds, _proposals(gt))["loss_cls"]
    loss.backward()
    _close(preds[0].detach().cpu().numpy(), s_ref.detach().numpy(), 1e-5, "scores")
    _close(loss.item(), l_ref.item(), 1e-5, "loss")
    _close(xx.grad.cpu().numpy(), xr.grad.numpy(), 1e-4, "dx")


def test_empty_batch_and_box_reg_and_inference():
    from cddmsl_b200.structures import Boxes, Instances

    cfg = synth.CONFIGS["tiny"]
    g = synth.generator(4)
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=64)
    m = _predictor(cfg.num_classes, cfg.emb_dim, w)
    # R = 0: gradient-connected zero (the reference crashes at fast_rcnn.py:626-627; documented divergence)
    x0 = torch.zeros(0, cfg.emb_dim, device=DEV, requires_grad=True)
    out = m.losses(m(x0), [])
    assert out["loss_cls"].item() == 0.0 and out["loss_box_reg"].item() == 0.0
    # box regression loss is plain PyTorch and must be finite / differentiable
    boxes = synth.make_boxes(64, 600, 1000, g, degenerate_frac=0.0)
    xx = x.to(DEV).requires_grad_(True)
    losses = m.losses(m(xx), _proposals(gt, boxes))
    (losses["loss_cls"] + losses["loss_box_reg"]).backward()
    assert torch.isfinite(xx.grad).all() and m.bbox_pred.weight.grad is not None
    # inference: softmax probabilities + class-aware NMS + top-k, against the oracle's pieces
    m.eval()
    m.test_score_thresh, m.test_nms_thresh, m.test_topk_per_image = 0.05, 0.5, 20
    inst = Instances((600, 1000))
    inst.proposal_boxes = Boxes(boxes.to(DEV))
    with torch.no_grad():
        preds = m(x.to(DEV))
        res, kept = m.inference(preds, [inst])
    probs = torch.softmax(preds[0], -1).cpu()
    pb = m.predict_boxes(preds, [inst])[0].cpu()
    bx = pb.view(-1, cfg.num_classes, 4).clone()
    bx[..., 0::2] = bx[..., 0::2].clamp(0, 1000)
    bx[..., 1::2] = bx[..., 1::2].clamp(0, 600)
    mask = probs[:, :-1] > 0.05
    inds = mask.nonzero()
    keep = torch_ref.batched_nms(bx[mask], probs[:, :-1][mask], inds[:, 1], 0.5)[:20]
    assert torch.equal(res[0].pred_classes.cpu(), inds[keep][:, 1])
    assert torch.allclose(res[0].scores.cpu(), probs[:, :-1][mask][keep], rtol=1e-6)
    assert torch.equal(kept[0].cpu(), inds[keep][:, 0])


@pytest.mark.parametrize("tag", ["agnostic", "perclass", "l1"])
def test_box_reg_loss_fixture(golden_dir, tag):
    """Fused box-regression loss through the C ABI vs the fixture built on the reference's get_deltas
    (fast_rcnn.py:646-689, box_regression.py:42-75).  fp32: 1e-5 forward, 1e-4 backward."""
    from cddmsl_b200 import ops

    f = np.load(os.path.join(golden_dir, "box_reg.npz"))
    prop, gtb, gt = (torch.from_numpy(f[k]).to(DEV) for k in ("prop", "gtb", "gt"))
    pred = torch.from_numpy(f[f"pred_{tag}"]).to(DEV).requires_grad_(True)
    w = [float(v) for v in f["weights"]]
    loss, _ = ops.box_reg_loss(prop, gtb, pred, gt, int(f["k"][0]), w, float(f[f"beta_{tag}"][0]), True)
    (loss * 3.0).backward()
    _close(loss.item(), float(f[f"loss_{tag}"][0]), 1e-5, "loss")
    _close(pred.grad.cpu().numpy(), 3.0 * f[f"dpred_{tag}"], 1e-4, "dpred")
    assert torch.isfinite(pred.grad).all()      # degenerate background proposals are never evaluated


def test_box_reg_loss_through_predictor_matches_oracle_and_is_sync_free():
    import cddmsl_b200.modeling.fast_rcnn as fr

    cfg = synth.CONFIGS["tiny"]
    g = synth.generator(21)
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=128)
    boxes = synth.make_boxes(128, 600, 1000, g, degenerate_frac=0.0)
    m = _predictor(cfg.num_classes, cfg.emb_dim, w)
    xx = x.to(DEV).requires_grad_(True)
    scores, deltas = m(xx)
    props = _proposals(gt, boxes)
    out = m.losses((scores, deltas), props)
    ref = torch_ref.box_reg_loss(boxes, boxes + 3.0, deltas.detach().cpu(), gt, cfg.num_classes,
                                 m.box2box_transform.weights, m.smooth_l1_beta)
    _close(out["loss_box_reg"].item(), ref.item(), 1e-5, "loss_box_reg")
    out["loss_box_reg"].backward()
    assert m.bbox_pred.weight.grad is not None and torch.isfinite(m.bbox_pred.weight.grad).all()
    # the strict (reference-shaped, syncing) evaluation gives the same number
    fr.STRICT_BOX_REG_SYNC = True
    try:
        strict = m.losses((scores.detach(), deltas.detach()), props)["loss_box_reg"].item()
    finally:
        fr.STRICT_BOX_REG_SYNC = False
    _close(out["loss_box_reg"].item(), strict, 1e-5, "strict")
    # empty batch
    z, _ = __import__("cddmsl_b200").ops.box_reg_loss(torch.zeros(0, 4, device=DEV), torch.zeros(0, 4, device=DEV),
                                                      torch.zeros(0, 4, device=DEV),
                                                      torch.zeros(0, dtype=torch.int64, device=DEV), 3,
                                                      [10.0, 10.0, 5.0, 5.0], 0.5, True)
    assert z.item() == 0.0


def test_inference_equals_reference_fast_rcnn_inference(golden_dir):
    """a6: `fast_rcnn_inference_single_image` (fast_rcnn.py:130-209) against outputs of the reference's own function on
    its own Boxes / Instances / batched_nms (tests/golden/make_golden.py:inference_ref_cases): kept rows and classes
    identical, boxes and scores bit-equal."""
    from cddmsl_b200.modeling.fast_rcnn import fast_rcnn_inference, fast_rcnn_inference_single_image

    f = np.load(os.path.join(golden_dir, "inference_ref.npz"))
    names = [str(n) for n in f["names"]]
    assert len(names) == 6
    per_image = {}
    for name in names:
        st, nt, topk, vis = (float(v) for v in f[f"params_{name}"])
        boxes, scores, sbf = (torch.from_numpy(f[f"{k}_{name}"]).to(DEV) for k in ("boxes", "scores", "sbf"))
        res, kept = fast_rcnn_inference_single_image(boxes, scores, (600, 1000), st, nt, False, "gaussian", 0.5, 0.001,
                                                     int(topk), sbf, bool(vis))
        assert np.array_equal(kept.cpu().numpy(), f[f"kept_{name}"]), name
        assert np.array_equal(res.pred_classes.cpu().numpy(), f[f"pred_classes_{name}"]), name
        assert np.array_equal(res.pred_boxes.tensor.cpu().numpy(), f[f"pred_boxes_{name}"]), name
        assert np.array_equal(res.scores.cpu().numpy(), f[f"pred_scores_{name}"]), name
        per_image[name] = (boxes, scores, sbf)
    # the list form (fast_rcnn.py:42-98) over two images with the same settings
    a, b = per_image["perclass"], per_image["perclass"]
    st, nt, topk, _ = (float(v) for v in f["params_perclass"])
    res, kept = fast_rcnn_inference([a[0], b[0]], [a[1], b[1]], [(600, 1000)] * 2, st, nt, topk_per_image=int(topk))
    for i in range(2):
        assert np.array_equal(kept[i].cpu().numpy(), f["kept_perclass"])
