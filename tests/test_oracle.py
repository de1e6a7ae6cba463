"""CPU: pins the oracle (oracle/c/*.c and oracle/torch_ref.py) against the reference's golden vectors and
the fixtures generated from the reference wrapper + installed torchvision CPU ops (tests/golden)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_ref, torch_ref


# ---- the reference's own golden vectors: tests/layers/test_roi_align.py:14-47 -----------------------
OLD = [[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]]
NEW = [[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]]


@pytest.mark.parametrize("impl", ["c", "torch"])
def test_reference_golden_forward(impl):
    inp = np.arange(25, dtype=np.float32).reshape(1, 1, 5, 5)
    roi = np.array([[0, 1, 1, 3, 3]], dtype=np.float32)
    for aligned, want in ((False, OLD), (True, NEW)):
        if impl == "c":
            out = c_ref.roi_align_fwd(inp, roi, (4, 4), 1.0, 0, aligned)[0, 0]
        else:
            out = torch_ref.roi_align(torch.from_numpy(inp), torch.from_numpy(roi), (4, 4), 1.0, 0, aligned)[0, 0].numpy()
        assert np.allclose(out.flatten(), np.asarray(want, dtype=np.float32).flatten())


def test_reference_empty_box_and_batch():
    # tests/layers/test_roi_align.py:111-128: empty box -> zero output AND zero input gradient; empty batch
    rng = np.random.RandomState(0)
    img = rng.rand(1, 1, 5, 5).astype(np.float32)
    roi = np.array([[0, 3, 4, 5, 4]], dtype=np.float32)
    out = c_ref.roi_align_fwd(img, roi, (7, 7), 1.0, 0, True)
    assert out.shape == (1, 1, 7, 7) and (out == 0).all()
    gin = c_ref.roi_align_bwd(np.ones_like(out), roi, img.shape, 1.0, 0, True)
    assert (gin == 0).all()
    out = c_ref.roi_align_fwd(np.zeros((0, 3, 10, 10), np.float32), np.zeros((0, 5), np.float32), (7, 7), 1.0, 0, True)
    assert out.shape == (0, 3, 7, 7)


def test_reference_resize_property():
    # tests/layers/test_roi_align.py:52-62 with an exact 2x2 box-filter downscale in place of cv2.resize
    rng = np.random.RandomState(1)
    H = W = 30
    inp = (rng.rand(H, W) * 100).astype(np.float32)
    out = c_ref.roi_align_fwd(inp[None, None], [[0, 10, 10, 20, 20]], (5, 5), 1.0, 0, True)
    inp2 = inp.reshape(H // 2, 2, W // 2, 2).mean(axis=(1, 3)).astype(np.float32)
    out2 = c_ref.roi_align_fwd(inp2[None, None], [[0, 5, 5, 10, 10]], (5, 5), 1.0, 0, True)
    assert np.abs(out2 - out).max() < 1e-4


# ---- fixtures from the reference wrapper + torchvision CPU ----------------------------------------
def _roi_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "roi_align_*.npz")))


def test_c_oracle_roi_align_bit_exact(golden_dir):
    files = _roi_files(golden_dir)
    assert len(files) >= 5
    for f in files:
        d = np.load(f)
        p, _, sr, al = d["meta"]
        sc = float(d["scale"][0])
        out = c_ref.roi_align_fwd(d["feat"], d["rois"], (int(p), int(p)), sc, int(sr), bool(al))
        gin = c_ref.roi_align_bwd(d["gout"], d["rois"], d["feat"].shape, sc, int(sr), bool(al))
        assert np.array_equal(out, d["out"]), f
        assert np.array_equal(gin, d["gin"]), f


def test_torch_oracle_roi_align_matches_fixture(golden_dir):
    for f in _roi_files(golden_dir):
        d = np.load(f)
        p, _, sr, al = d["meta"]
        x = torch.from_numpy(d["feat"]).requires_grad_(True)
        out = torch_ref.roi_align(x, torch.from_numpy(d["rois"]), (int(p), int(p)), float(d["scale"][0]), int(sr), bool(al))
        out.backward(torch.from_numpy(d["gout"]))
        assert np.array_equal(out.detach().numpy(), d["out"])
        assert np.array_equal(x.grad.numpy(), d["gin"])


def canon_ties(keep, scores):
    """kept indices with exactly equal scores put in ascending index order (the reference's per-class branch
    re-sorts with an unstable sort, so that order is unspecified there)."""
    keep = np.asarray(keep)
    s = scores[keep]
    order = np.lexsort((keep, -s))
    return keep[order]


def test_c_oracle_nms_matches_torchvision(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "nms_*.npz")))
    assert len(files) >= 5
    for f in files:
        d = np.load(f)
        for t in d["thrs"]:
            ref = d[f"keep_{t}"]
            got = c_ref.batched_nms(d["boxes"], d["scores"], d["idxs"], float(t))
            vanilla = d["boxes"].size > 4000
            if vanilla:
                assert np.array_equal(got, canon_ties(ref, d["scores"])), (f, t)
                assert np.array_equal(d["scores"][got], d["scores"][ref])
            else:
                assert np.array_equal(got, ref), (f, t)


def test_torch_oracle_nms_matches_fixture(golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "nms_*.npz"))):
        d = np.load(f)
        for t in d["thrs"]:
            got = torch_ref.batched_nms(torch.from_numpy(d["boxes"]), torch.from_numpy(d["scores"]),
                                        torch.from_numpy(d["idxs"]), float(t)).numpy()
            assert np.array_equal(got, d[f"keep_{t}"])


def test_nms_threshold_is_compared_in_double():
    # inter 1, union 5 -> fp32 IoU is exactly float32(0.2) = 0.2000000030 which is > 0.2 as a double:
    # the CPU kernel (double threshold) suppresses, a float-threshold compare would not.
    boxes = np.array([[0, 0, 1, 3], [0, 2, 1, 5]], dtype=np.float32)
    scores = np.array([0.9, 0.8], dtype=np.float32)
    assert np.float32(1) / np.float32(5) == np.float32(0.2)
    got = c_ref.batched_nms(boxes, scores, None, 0.2)
    ref = torch_ref.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.zeros(2, dtype=torch.int64),
                                0.2).numpy()
    assert np.array_equal(got, ref) and len(got) == 1
    # and 0.7: float32(0.7) < 0.7, IoU == float32(0.7) must NOT suppress (inter 7, union 10)
    boxes = np.array([[0, 0, 1, 8.5], [0, 1.5, 1, 10]], dtype=np.float32)
    assert np.float32(7) / np.float32(10) == np.float32(0.7)
    got = c_ref.batched_nms(boxes, scores, None, 0.7)
    ref = torch_ref.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.zeros(2, dtype=torch.int64),
                                0.7).numpy()
    assert np.array_equal(got, ref) and len(got) == 2


def test_head_and_align_fixtures_reproduce(golden_dir):
    d = np.load(os.path.join(golden_dir, "head_tiny.npz"))
    T, gamma, bgw = d["params"]
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    k = w.shape[0]
    for tag, wb in (("zero_bg", torch.zeros(1, x.shape[1])), ("learned_bg", torch.from_numpy(d["w_bg2"]))):
        xx = x.clone().requires_grad_(True)
        s = torch_ref.clip_head_scores(xx, w, wb, float(T))
        loss = torch_ref.focal_loss(s, gt, k, float(gamma), float(bgw))
        loss.backward()
        assert np.allclose(s.detach().numpy(), d[f"scores_{tag}"], rtol=1e-6, atol=1e-5)
        assert np.allclose(loss.item(), d[f"loss_{tag}"], rtol=1e-6)
        assert np.allclose(xx.grad.numpy(), d[f"dx_{tag}"], rtol=1e-5, atol=1e-7, equal_nan=True)
    a = np.load(os.path.join(golden_dir, "align_single.npz"))
    for tag in ("n16", "n40", "n256"):
        loss = torch_ref.caption_consistency_loss(torch.from_numpy(a[f"a_{tag}"]), torch.from_numpy(a[f"b_{tag}"]))
        assert np.allclose(loss.item(), a[f"loss_{tag}"], rtol=1e-6)


def test_oracle_head_equals_reference_fast_rcnn(golden_dir):
    """head_ref.npz was produced by the reference's OWN FastRCNNOutputLayers (fast_rcnn.py loaded verbatim by
    tests/golden/make_golden.py:head_ref_cases): forward :529-572, losses :574-622, focal_loss :624-644,
    _log_classification_stats :100-127.  The restatement in oracle/torch_ref.py must reproduce it."""
    d = np.load(os.path.join(golden_dir, "head_ref.npz"))
    T, gamma, bgw = (float(v) for v in d["params"])
    x, w, gt = torch.from_numpy(d["x"]), torch.from_numpy(d["w"]), torch.from_numpy(d["gt"])
    k = w.shape[0]
    for tag, wb in (("zero_bg", torch.zeros(1, x.shape[1])), ("learned_bg", torch.from_numpy(d["w_bg2"]))):
        for mode, (gam, bw) in {"focal": (gamma, bgw), "ce": (None, None), "wce": (None, bgw)}.items():
            xx = x.clone().requires_grad_(True)
            s = torch_ref.clip_head_scores(xx, w, wb, T)
            loss = torch_ref.cls_loss(s, gt, k, gam, bw)
            loss.backward()
            assert np.array_equal(s.detach().numpy(), d[f"scores_{tag}"]), "logits are bit-equal to the reference"
            assert np.allclose(loss.item(), d[f"loss_{mode}_{tag}"], rtol=1e-6, atol=0)
            assert np.allclose(xx.grad.numpy(), d[f"dx_{mode}_{tag}"], rtol=1e-5, atol=1e-7, equal_nan=True)
        acc, nfg, fgacc, fn = torch_ref.classification_stats(s.detach(), gt)
        want = d[f"stats_{tag}"]
        assert np.allclose([acc / len(gt), fgacc / nfg, fn / nfg], want, rtol=0, atol=1e-12)
    assert np.isnan(d["dx_focal_zero_bg"]).any(), "the fixture keeps the reference's NaN row at softmax saturation"


def test_oracle_alignment_equals_reference_rcnn_lines(golden_dir):
    """align_ref.npz was produced by executing the literal lines rcnn.py:455-470 / :305-319 / :270-272 (read from
    the reference file at generation time), single-process and under a 2-rank gloo group with the reference's own
    GatherLayer."""
    a = np.load(os.path.join(golden_dir, "align_ref.npz"))
    for tag in ("n16", "n48", "n256"):
        for kind in ("region", "image"):
            x, y = torch.from_numpy(a[f"a_{tag}"]), torch.from_numpy(a[f"b_{tag}"])
            xx, yy = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            # region level: S = src @ tgt^T with (src, tgt) = (a, b); image level: S = trgt @ src^T with (trgt, src) = (a, b)
            loss = torch_ref.caption_consistency_loss(xx, yy)
            loss.backward()
            assert np.allclose(loss.item(), a[f"{kind}_loss_{tag}"], rtol=1e-6, atol=0)
            assert np.allclose(xx.grad.numpy(), a[f"{kind}_da_{tag}"], rtol=1e-5, atol=1e-8)
            assert np.allclose(yy.grad.numpy(), a[f"{kind}_db_{tag}"], rtol=1e-5, atol=1e-8)
    t, s_ = torch.from_numpy(a["kd_teacher"]), torch.from_numpy(a["kd_student"]).requires_grad_(True)
    kl = torch_ref.kd_l1_loss(t, s_)
    kl.backward()
    assert np.allclose(kl.item(), a["kd_loss"], rtol=1e-6) and np.array_equal(s_.grad.numpy(), a["kd_dstudent"])
    al = [torch.from_numpy(a["w2_a0"]), torch.from_numpy(a["w2_a1"])]
    bl = [torch.from_numpy(a["w2_b0"]), torch.from_numpy(a["w2_b1"])]
    loss, ga, gb = torch_ref.caption_consistency_world(al, bl)
    for kind in ("region", "image"):
        assert np.allclose(loss.item(), a[f"w2_{kind}_loss"], rtol=1e-6)
        for r in range(2):
            assert np.allclose(ga[r].numpy(), a[f"w2_{kind}_da{r}"], rtol=1e-5, atol=1e-8)
            assert np.allclose(gb[r].numpy(), a[f"w2_{kind}_db{r}"], rtol=1e-5, atol=1e-8)


def test_world_emulation_matches_real_gatherlayer_fixture(golden_dir):
    w = np.load(os.path.join(golden_dir, "align_world2.npz"))
    a = [torch.from_numpy(w["a0"]), torch.from_numpy(w["a1"])]
    b = [torch.from_numpy(w["b0"]), torch.from_numpy(w["b1"])]
    loss, ga, gb = torch_ref.caption_consistency_world(a, b)
    assert np.allclose(loss.item(), w["loss"], rtol=1e-6)
    for r in range(2):
        assert np.allclose(ga[r].numpy(), w[f"da{r}"], rtol=1e-5, atol=1e-8)
        assert np.allclose(gb[r].numpy(), w[f"db{r}"], rtol=1e-5, atol=1e-8)


def test_focal_saturation_nan_is_a_reference_property():
    # SURVEY.md §7 hard part 5: autograd gives NaN when softmax saturates to p_t == 1
    s = torch.tensor([[100.0, 0.0, -50.0]], requires_grad=True)
    loss = torch_ref.focal_loss(s, torch.tensor([0]), 2, 0.5, 0.2)
    loss.backward()
    assert loss.item() == 0.0 and torch.isnan(s.grad).all()


def test_box_reg_oracle_matches_reference_get_deltas_fixture(golden_dir):
    """box_reg.npz: get_deltas from the reference's own Box2BoxTransform (box_regression.py:42-75)."""
    f = np.load(os.path.join(golden_dir, "box_reg.npz"))
    prop, gtb, gt = torch.from_numpy(f["prop"]), torch.from_numpy(f["gtb"]), torch.from_numpy(f["gt"])
    fg = torch.from_numpy(f["fg"])
    w = tuple(float(v) for v in f["weights"])
    k = int(f["k"][0])
    d = torch_ref.get_deltas(prop[fg], gtb[fg], w)
    assert np.array_equal(d.numpy(), f["ref_deltas"])          # same fp32 arithmetic, same order: bit-equal
    for tag in ("agnostic", "perclass", "l1"):
        pred = torch.from_numpy(f[f"pred_{tag}"]).requires_grad_(True)
        loss = torch_ref.box_reg_loss(prop, gtb, pred, gt, k, w, float(f[f"beta_{tag}"][0]))
        loss.backward()
        assert abs(loss.item() - float(f[f"loss_{tag}"][0])) <= 1e-6 * abs(loss.item())
        assert np.allclose(pred.grad.numpy(), f[f"dpred_{tag}"], rtol=1e-6, atol=0)
    # the reference asserts on degenerate FOREGROUND proposals (box_regression.py:74)
    bad = prop.clone()
    bad[fg[0]] = torch.tensor([3.0, 3.0, 3.0, 9.0])
    with pytest.raises(AssertionError):
        torch_ref.box_reg_loss(bad, gtb, torch.zeros(len(gt), 4), gt, k, w, 0.5)


MATCH_CFG = {"roi": ([0.5], [0, 1], False), "rpn": ([0.3, 0.7], [0, -1, 1], True)}


def test_matcher_oracle_and_host_mirror_match_reference_fixture(golden_dir):
    """match.npz comes from the reference's own Matcher + pairwise_iou (matcher.py, boxes.py loaded verbatim)."""
    from cddmsl_b200.modeling import matcher as host

    f = np.load(os.path.join(golden_dir, "match.npz"))
    for b in range(int(f["n_images"][0])):
        gt, boxes = torch.from_numpy(f[f"gt_{b}"]), torch.from_numpy(f[f"boxes_{b}"])
        mqm = torch_ref.pairwise_iou(gt, boxes)
        assert torch.equal(mqm, host.pairwise_iou(gt, boxes))
        if len(gt):
            assert np.array_equal(mqm.max(dim=0).values.numpy(), f[f"iou_max_{b}"])
        for tag, (thr, lab, low) in MATCH_CFG.items():
            for fn in (lambda q: torch_ref.matcher(q, thr, lab, low), host.Matcher(thr, lab, low)):
                m, l = fn(mqm)
                assert np.array_equal(m.numpy(), f[f"matches_{tag}_{b}"]), (tag, b)
                assert np.array_equal(l.numpy(), f[f"labels_{tag}_{b}"]), (tag, b)


def test_oracle_pretraining_losses_equal_reference_clip_rcnn_lines(golden_dir):
    """pretrain_ref.npz was produced by executing the literal lines clip_rcnn.py:590-611 / :624-640 with the literal
    class MILCrossEntropy (utils/comm.py:332-355), read from the reference at generation time."""
    p = np.load(os.path.join(golden_dir, "pretrain_ref.npz"))
    for tag in ("small", "lvis"):
        t = lambda k: torch.from_numpy(p[f"{k}_{tag}"])
        for which, key in ((0, "distill"), (1, "contrastive")):
            x = t("feats").clone().requires_grad_(True)
            loss = torch_ref.region_concept_losses(x, t("concept_emb"), t("teacher"), t("target_embs"), t("label_mtx"),
                                                   float(p[f"temp_{tag}"][0]))[which]
            loss.backward()
            assert np.allclose(loss.item(), p[f"{key}_{tag}"], rtol=1e-6), key
            assert np.allclose(x.grad.numpy(), p[f"{key}_dx_{tag}"], rtol=1e-5, atol=1e-8), key
    a = torch.from_numpy(p["it_feats"]).requires_grad_(True)
    b = torch.from_numpy(p["it_text"]).requires_grad_(True)
    loss = torch_ref.image_text_matching_loss(a, b, float(p["it_temp"][0]))
    loss.backward()
    assert np.allclose(loss.item(), p["it_loss"], rtol=1e-6)
    assert np.allclose(a.grad.numpy(), p["it_dfeats"], rtol=1e-5, atol=1e-8)
    assert np.allclose(b.grad.numpy(), p["it_dtext"], rtol=1e-5, atol=1e-8)
