"""Proposal matching / sampling through the C ABI vs the reference's own Matcher (fixture
match.npz) and vs the per-image reference-shaped loop.  Matching is integer/index work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from cddmsl_b200 import synth
from oracle import torch_ref

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
MATCH_CFG = {"roi": ([0.5], [0, 1], False), "rpn": ([0.3, 0.7], [0, -1, 1], True)}


def _padded(f):
    nb = int(f["n_images"][0])
    gts = [torch.from_numpy(f[f"gt_{b}"]) for b in range(nb)]
    bxs = [torch.from_numpy(f[f"boxes_{b}"]) for b in range(nb)]
    gmax, mmax = max(max(len(g) for g in gts), 1), max(len(x) for x in bxs)
    gt = torch.zeros(nb, gmax, 4)
    bx = torch.zeros(nb, mmax, 4)
    for b in range(nb):
        gt[b, : len(gts[b])] = gts[b]
        bx[b, : len(bxs[b])] = bxs[b]
    return nb, gts, bxs, gt, bx


@pytest.mark.parametrize("tag", ["roi", "rpn"])
def test_fused_matching_equals_reference_matcher_fixture(golden_dir, tag):
    from cddmsl_b200.modeling import Matcher

    f = np.load(os.path.join(golden_dir, "match.npz"))
    nb, gts, bxs, gt, bx = _padded(f)
    thr, lab, low = MATCH_CFG[tag]
    gc = torch.tensor([len(g) for g in gts], dtype=torch.int32, device=DEV)
    mc = torch.tensor([len(x) for x in bxs], dtype=torch.int32, device=DEV)
    matches, labels, vals = Matcher(thr, lab, low).match_boxes(gt.to(DEV), gc, bx.to(DEV), mc)
    for b in range(nb):
        n = len(bxs[b])
        assert np.array_equal(matches[b, :n].cpu().numpy(), f[f"matches_{tag}_{b}"]), b
        assert np.array_equal(labels[b, :n].cpu().numpy(), f[f"labels_{tag}_{b}"]), b
        assert np.array_equal(vals[b, :n].cpu().numpy(), f[f"iou_max_{b}"]), b          # fp32 IoU, bit-exact
        assert not matches[b, n:].any() and not labels[b, n:].any()                     # padding untouched


def test_many_ground_truth_boxes_and_anchor_scale():
    """More gt boxes than one shared-memory chunk (1024) and an RPN-sized candidate set, vs the oracle."""
    from cddmsl_b200.modeling import Matcher

    g = synth.generator(41)
    gt = synth.make_boxes(1500, 600, 1000, g, min_side=8.0, degenerate_frac=0.0)
    boxes = synth.make_boxes(35910, 600, 1000, g, degenerate_frac=0.01)
    for thr, lab, low in MATCH_CFG.values():
        want_m, want_l = torch_ref.matcher(torch_ref.pairwise_iou(gt, boxes), thr, lab, low)
        m, l, _ = Matcher(thr, lab, low).match_boxes(gt[None].to(DEV), torch.tensor([1500], dtype=torch.int32, device=DEV),
                                                     boxes[None].to(DEV), None)
        assert torch.equal(m[0].cpu(), want_m) and torch.equal(l[0].cpu(), want_l)


def test_label_and_sample_proposals_batched():
    """roi_heads.py:236-319: deterministic parts equal the reference flow, the random subset obeys its law."""
    import cddmsl_b200.modeling.roi_heads as rh
    from cddmsl_b200.structures import Boxes, Instances

    g = synth.generator(42)
    k = 20
    props, tgts = [], []
    for ng, nm in [(6, 2000), (0, 1500), (2, 2000)]:
        t = Instances((600, 1000))
        t.gt_boxes = Boxes(synth.make_boxes(ng, 600, 1000, g, min_side=48.0, degenerate_frac=0.0).to(DEV) if ng
                           else torch.zeros(0, 4, device=DEV))
        t.gt_classes = torch.randint(0, k, (ng,), generator=g).to(DEV)
        p = Instances((600, 1000))
        b = synth.make_boxes(nm, 600, 1000, g, degenerate_frac=0.0)
        if ng:
            b[:400] = t.gt_boxes.tensor.cpu()[torch.randint(0, ng, (400,), generator=g)] + torch.randn(400, 4, generator=g) * 5
        p.proposal_boxes = Boxes(b.to(DEV))
        p.objectness_logits = torch.randn(nm, generator=g).to(DEV)
        props.append(p)
        tgts.append(t)
    heads = rh.ROIHeads(num_classes=k)
    seen = {}
    rh._fr.set_scalar_sink(lambda name, v: seen.__setitem__(name, v))
    try:
        out = heads.label_and_sample_proposals(props, tgts)
    finally:
        rh._fr.set_scalar_sink(None)
    assert set(seen) == {"roi_head/num_fg_samples", "roi_head/num_bg_samples"}
    for p, t, o in zip(props, tgts, out):
        ng = len(t)
        all_boxes = torch.cat([p.proposal_boxes.tensor, t.gt_boxes.tensor]).cpu()
        mqm = torch_ref.pairwise_iou(t.gt_boxes.tensor.cpu(), all_boxes)
        m, l = torch_ref.matcher(mqm, [0.5], [0, 1], False)
        cls = torch_ref.assign_classes(m, l, t.gt_classes.cpu(), k)
        n_pos_all = int(((cls != -1) & (cls != k)).sum())
        n_neg_all = int((cls == k).sum())
        got_cls = o.gt_classes.cpu()
        n_fg = int((got_cls != k).sum())
        assert n_fg == min(n_pos_all, 128) and len(o) == n_fg + min(n_neg_all, 512 - n_fg)
        # every sampled row is a row of the candidate set carrying exactly the class / gt box the reference assigns
        ob = o.proposal_boxes.tensor.cpu()
        eq = (ob[:, None, :] == all_boxes[None, :, :]).all(-1)
        assert eq.any(1).all()
        idx = eq.float().argmax(1)
        assert len(torch.unique(idx)) == len(idx) or ng > 0      # (gt duplicates may repeat a box value)
        assert torch.equal(got_cls, cls[idx])
        if ng:
            assert torch.equal(o.gt_boxes.tensor.cpu(), t.gt_boxes.tensor.cpu()[m[idx]])
        else:
            assert not o.has("gt_boxes") and (got_cls == k).all()
    # the upstream-shaped loop gives samples of the same sizes
    rh.BATCHED_IMAGES = False
    try:
        loop = heads.label_and_sample_proposals(props, tgts)
    finally:
        rh.BATCHED_IMAGES = True
    assert [len(a) for a in loop] == [len(a) for a in out]
    assert [int((a.gt_classes != k).sum()) for a in loop] == [int((a.gt_classes != k).sum()) for a in out]


class _UpstreamMatcher:
    """The attributes detectron2/modeling/matcher.py:36-61 leaves on a Matcher (thresholds WITH the +-inf sentinels)."""

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        self.thresholds = [-float("inf")] + list(thresholds) + [float("inf")]
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches


class _UpstreamROIHeads:
    """Only what detectron2/modeling/roi_heads/roi_heads.py:140-164 sets on the reference class."""

    def __init__(self, k, only_fg=False):
        self.num_classes = k
        self.batch_size_per_image = 512
        self.positive_fraction = 0.25
        self.proposal_matcher = _UpstreamMatcher([0.5], [0, 1])
        self.proposal_append_gt = True
        self.only_sample_fg_proposals = only_fg


def _toy_batch(g, k):
    from cddmsl_b200.structures import Boxes, Instances

    props, tgts = [], []
    for ng, nm in [(5, 900), (0, 700), (3, 1100)]:
        t = Instances((600, 1000))
        t.gt_boxes = Boxes(synth.make_boxes(ng, 600, 1000, g, min_side=48.0, degenerate_frac=0.0).to(DEV) if ng
                           else torch.zeros(0, 4, device=DEV))
        t.gt_classes = torch.randint(0, k, (ng,), generator=g).to(DEV)
        p = Instances((600, 1000))
        b = synth.make_boxes(nm, 600, 1000, g, degenerate_frac=0.0)
        if ng:
            b[:300] = t.gt_boxes.tensor.cpu()[torch.randint(0, ng, (300,), generator=g)] + torch.randn(300, 4, generator=g) * 5
        p.proposal_boxes = Boxes(b.to(DEV))
        p.objectness_logits = torch.randn(nm, generator=g).to(DEV)
        props.append(p)
        tgts.append(t)
    return props, tgts


def test_patched_into_an_upstream_shaped_class():
    """INTEGRATION.md assigns `label_and_sample_proposals` to the REFERENCE's ROIHeads: the function may only use
    what that class has (no helper methods, the upstream Matcher's attributes)."""
    from cddmsl_b200.modeling import label_and_sample_proposals

    k = 20
    props, tgts = _toy_batch(synth.generator(43), k)
    _UpstreamROIHeads.label_and_sample_proposals = label_and_sample_proposals
    out = _UpstreamROIHeads(k).label_and_sample_proposals(props, tgts)
    assert len(out) == 3
    for o, t in zip(out, tgts):
        n_fg = int((o.gt_classes != k).sum())
        assert len(o) <= 512 and n_fg <= 128 and o.has("gt_classes")
        if len(t):
            assert o.has("gt_boxes") and len(o.gt_boxes) == len(o)


def test_only_sample_fg_proposals():
    """MODEL.CLIP.ONLY_SAMPLE_FG_PROPOSALS (roi_heads.py:216-228): positives only; an image without ground truth keeps
    one background proposal."""
    import cddmsl_b200.modeling.roi_heads as rh

    k = 20
    props, tgts = _toy_batch(synth.generator(44), k)
    heads = rh.ROIHeads(num_classes=k, only_sample_fg_proposals=True)
    for batched in (True, False):
        rh.BATCHED_IMAGES = batched
        try:
            out = heads.label_and_sample_proposals(props, tgts)
        finally:
            rh.BATCHED_IMAGES = True
        assert (out[0].gt_classes != k).all() and 0 < len(out[0]) <= 128
        assert len(out[1]) == 1 and int(out[1].gt_classes[0]) == k
        assert (out[2].gt_classes != k).all() and 0 < len(out[2]) <= 128
