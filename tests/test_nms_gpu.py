"""GPU parity: batched NMS must return bit-identical kept indices to the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from cddmsl_b200 import synth
from oracle import c_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gpu(boxes, scores, idxs, thr):
    from cddmsl_b200.layers import batched_nms

    return batched_nms(torch.as_tensor(boxes).to(DEV), torch.as_tensor(scores).to(DEV), torch.as_tensor(idxs).to(DEV),
                       thr).cpu().numpy()


def _canon(keep, scores):
    keep = np.asarray(keep)
    return keep[np.lexsort((keep, -scores[keep]))]


def test_fixtures_from_torchvision_cpu(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "nms_*.npz")))
    assert len(files) >= 5
    for f in files:
        d = np.load(f)
        for t in d["thrs"]:
            got = _gpu(d["boxes"], d["scores"], d["idxs"], float(t))
            ref = d[f"keep_{t}"]
            if d["boxes"].size > 4000:   # per-class branch upstream: order among exactly tied scores unspecified
                assert np.array_equal(got, _canon(ref, d["scores"])), (f, t)
            else:
                assert np.array_equal(got, ref), (f, t)


@pytest.mark.parametrize("m,k", [(1, 1), (2, 1), (63, 1), (64, 1), (65, 3), (129, 1), (1000, 1), (1000, 20),
                                 (4097, 7), (12000, 1), (12000, 20)])
@pytest.mark.parametrize("thr", [0.5, 0.7])
def test_seeded_random_bit_exact(m, k, thr):
    g = synth.generator(1000 + m + k)
    boxes, scores, idxs = synth.make_nms_inputs(m, 600, 1000, g, num_classes=k, tie_frac=0.02)
    got = _gpu(boxes, scores, idxs, thr)
    want = c_ref.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr)
    assert np.array_equal(got, want)


def test_reference_test_shape_and_inputs_not_mutated():
    # tests/layers/test_nms.py:16-29 upstream: N=2000, 50 classes, IoU in {.2,.5,.8}; boxes must not be modified
    torch.manual_seed(7)
    n = 2000
    b = torch.rand(n, 4) * 100
    b.clamp_(min=1.0)
    b[:, 2:] += b[:, :2]
    s, ids = torch.rand(n), torch.randint(0, 50, (n,))
    bd = b.to(DEV)
    backup = bd.clone()
    from cddmsl_b200.layers import batched_nms

    for iou in (0.2, 0.5, 0.8):
        got = batched_nms(bd, s.to(DEV), ids.to(DEV), iou).cpu().numpy()
        assert torch.equal(bd, backup)
        assert np.array_equal(got, c_ref.batched_nms(b.numpy(), s.numpy(), ids.numpy(), iou))


def test_threshold_compared_in_double_and_nan_iou():
    boxes = np.array([[0, 0, 1, 3], [0, 2, 1, 5]], dtype=np.float32)   # IoU == float32(0.2) > 0.2
    scores = np.array([0.9, 0.8], dtype=np.float32)
    assert len(_gpu(boxes, scores, np.zeros(2, np.int64), 0.2)) == 1
    boxes = np.array([[0, 0, 1, 8.5], [0, 1.5, 1, 10]], dtype=np.float32)  # IoU == float32(0.7) < 0.7
    assert len(_gpu(boxes, scores, np.zeros(2, np.int64), 0.7)) == 2
    # zero-area duplicates: 0/0 = NaN IoU never suppresses
    boxes = np.array([[5, 5, 5, 5], [5, 5, 5, 5], [5, 5, 5, 5]], dtype=np.float32)
    scores = np.array([0.3, 0.9, 0.3], dtype=np.float32)
    got = _gpu(boxes, scores, np.zeros(3, np.int64), 0.5)
    assert np.array_equal(got, c_ref.batched_nms(boxes, scores, np.zeros(3, np.int64), 0.5))
    assert np.array_equal(got, [1, 0, 2])


def test_both_class_modes_and_plain_nms():
    from cddmsl_b200 import ops
    from cddmsl_b200.layers import nms

    g = synth.generator(77)
    boxes, scores, idxs = synth.make_nms_inputs(3000, 600, 1000, g, num_classes=9, tie_frac=0.02)
    for trick in (True, False):
        got = ops.batched_nms(boxes.to(DEV), scores.to(DEV), idxs.to(DEV), 0.6, trick).cpu().numpy()
        assert np.array_equal(got, c_ref.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.6, coord_trick=trick))
    got = nms(boxes.to(DEV), scores.to(DEV), 0.6).cpu().numpy()
    assert np.array_equal(got, c_ref.batched_nms(boxes.numpy(), scores.numpy(), None, 0.6))
    assert nms(torch.zeros(0, 4, device=DEV), torch.zeros(0, device=DEV), 0.5).numel() == 0


def test_above_40000_boxes_branch():
    # nms.py:28: len(boxes) >= 40000 takes detectron2's own per-class loop
    g = synth.generator(5)
    boxes, scores, idxs = synth.make_nms_inputs(41000, 1024, 2048, g, num_classes=4, tie_frac=0.0)
    got = _gpu(boxes, scores, idxs, 0.7)
    assert np.array_equal(got, c_ref.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.7, coord_trick=False))


@pytest.mark.parametrize("m,k,thr", [(24576, 1, 0.7), (30001, 6, 0.5), (65536, 3, 0.7)])
def test_large_single_image_bit_exact(m, k, thr):
    """24 576 - 65 536 boxes in one image: same keep list as the oracle, incl. score ties, a ragged last 64-box block
    and per-class suppression (the scan's removed-bitmap spans more than one pass of 256 column words here)."""
    g = synth.generator(77 + m)
    boxes, scores, idxs = synth.make_nms_inputs(m, 1024, 2048, g, num_classes=k, tie_frac=0.01)
    got = _gpu(boxes, scores, idxs, thr)
    want = c_ref.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr, coord_trick=False)
    assert np.array_equal(got, _canon(want, scores.numpy()))


@pytest.mark.parametrize("max_keep", [1, 64, 700, 2000, 5000])
def test_topk_limited_batched_nms_is_a_prefix_of_the_full_result(max_keep):
    """cddmsl_nms_batched_topk (`keep[:post_nms_topk]`, proposal_utils.py:116-118): bit-identical prefix of the full keep
    list for every image -- images that finish inside the first pass (few overlaps), images that need the full pass
    (heavily overlapping boxes: fewer than max_keep survive among the top 2*max_keep), short and empty images."""
    from cddmsl_b200.layers import batched_nms_images

    g = synth.generator(123)
    m = 6000
    bx, sc, counts = [], [], []
    for i in range(5):
        b, s, _ = synth.make_nms_inputs(m, 600, 1000, g, num_classes=1, tie_frac=0.01)
        if i == 1:      # near-duplicates of 40 boxes: almost everything is suppressed
            base = b[:40]
            b = base[torch.randint(0, 40, (m,), generator=g)] + torch.randn(m, 4, generator=g) * 1.5
        if i == 2:      # the best-scoring third is one tight cluster, the rest is spread out
            o = torch.argsort(s, descending=True)[: m // 3]
            b[o] = b[o[0]] + torch.randn(len(o), 4, generator=g) * 2.0
        bx.append(b)
        sc.append(s)
        counts.append([m, m, m, 300, 0][i])
    boxes, scores = torch.stack(bx).to(DEV), torch.stack(sc).to(DEV)
    cnt = torch.tensor(counts, dtype=torch.int32, device=DEV)
    full, nfull = batched_nms_images(boxes, scores, None, cnt, 0.7)
    part, npart = batched_nms_images(boxes, scores, None, cnt, 0.7, max_keep=max_keep)
    nfull, npart = nfull.tolist(), npart.tolist()
    for i in range(5):
        want = min(nfull[i], max_keep)
        assert npart[i] == want, (i, npart[i], want)
        assert torch.equal(part[i, :want], full[i, :want]), i
    # and against the oracle for one image
    ref = c_ref.batched_nms(bx[2].numpy(), sc[2].numpy(), np.zeros(m, np.int64), 0.7)
    assert np.array_equal(part[2, : npart[2]].cpu().numpy(), _canon(ref, sc[2].numpy())[: npart[2]])


@pytest.mark.parametrize("m,k,max_keep", [(20000, 20, 100), (5000, 8, 100), (3000, 3, 1000), (900, 1, 5)])
def test_topk_limited_single_image_nms_is_a_prefix(m, k, max_keep):
    """cddmsl_nms_topk (test-time `keep[:topk_per_image]`, fast_rcnn.py:186-187), class-aware and coordinate-trick modes."""
    from cddmsl_b200.layers import batched_nms

    g = synth.generator(321 + m)
    boxes, scores, idxs = synth.make_nms_inputs(m, 600, 1000, g, num_classes=k, tie_frac=0.01)
    if m == 5000:   # one class is a single tight cluster: the limited pass cannot fill max_keep from the top boxes alone
        boxes[:] = boxes[0] + torch.randn(m, 4, generator=g) * 1.0
    b, s, i = boxes.to(DEV), scores.to(DEV), idxs.to(DEV)
    full = batched_nms(b, s, i, 0.5)
    part = batched_nms(b, s, i, 0.5, max_keep=max_keep)
    assert len(part) == min(len(full), max_keep)
    assert torch.equal(part, full[: len(part)])


def test_presorted_scores_skip_the_sort():
    """`presorted` (the RPN path hands over its sorted top-k): identical to the call that sorts, incl. tied scores."""
    from cddmsl_b200.layers import batched_nms_images

    g = synth.generator(55)
    m = 5000
    bx, sc = [], []
    for _ in range(3):
        b, s, _ = synth.make_nms_inputs(m, 600, 1000, g, num_classes=1, tie_frac=0.05)
        o = torch.sort(s, descending=True, stable=True).indices
        bx.append(b[o])
        sc.append(s[o])
    boxes, scores = torch.stack(bx).to(DEV), torch.stack(sc).to(DEV)
    cnt = torch.tensor([m, 4097, 1], dtype=torch.int32, device=DEV)
    for max_keep in (0, 1000):
        a, na = batched_nms_images(boxes, scores, None, cnt, 0.7, max_keep=max_keep)
        b, nb = batched_nms_images(boxes, scores, None, cnt, 0.7, max_keep=max_keep, presorted=True)
        assert torch.equal(na, nb)
        for i, n in enumerate(na.tolist()):
            assert torch.equal(a[i, :n], b[i, :n])


def test_large_properties_256k():
    """configs[4] upper end (256k boxes): sorted by score, idempotent, and no kept pair overlaps above thr
    (checked on the top-scoring 4096 kept boxes with the exact fp32 formula)."""
    from cddmsl_b200.layers import nms

    g = synth.generator(9)
    m = 262144
    boxes, scores, _ = synth.make_nms_inputs(m, 1024, 2048, g, tie_frac=0.0)
    bd, sd = boxes.to(DEV), scores.to(DEV)
    keep = nms(bd, sd, 0.7)
    ks = sd[keep]
    assert (ks[:-1] >= ks[1:]).all()
    again = nms(bd[keep], ks, 0.7)
    assert torch.equal(again, torch.arange(len(keep), device=DEV))
    top = bd[keep[:4096]]
    area = (top[:, 2] - top[:, 0]) * (top[:, 3] - top[:, 1])
    lt = torch.max(top[:, None, :2], top[None, :, :2])
    rb = torch.min(top[:, None, 2:], top[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    iou = inter / (area[:, None] + area[None, :] - inter)
    iou.fill_diagonal_(0)
    assert not (iou.double() > 0.7).any()


def test_find_top_rpn_proposals_matches_oracle():
    from cddmsl_b200.modeling import find_top_rpn_proposals
    from oracle import torch_ref

    g = synth.generator(21)
    n_img, a = 2, 6000
    props = torch.stack([synth.make_boxes(a, 700, 1100, g, degenerate_frac=0.02) - 40.0 for _ in range(n_img)])
    logits = torch.randn(n_img, a, generator=g)
    res = find_top_rpn_proposals([props.to(DEV)], [logits.to(DEV)], [(600, 1000)] * n_img, 0.7, 4000, 1000, 0.0, True)
    for i in range(n_img):
        lg, idx = logits[i].sort(descending=True)
        wb, ws = torch_ref.find_top_rpn_proposals_single_image(props[i][idx[:4000]], lg[:4000], (600, 1000), 0.7, 1000)
        assert torch.equal(res[i].proposal_boxes.tensor.cpu(), wb)
        assert torch.equal(res[i].objectness_logits.cpu(), ws)
    bad = props.clone()
    bad[0, 3, 2] = float("inf")
    with pytest.raises(FloatingPointError):
        find_top_rpn_proposals([bad.to(DEV)], [logits.to(DEV)], [(600, 1000)] * n_img, 0.7, 4000, 1000, 0.0, True)


@pytest.mark.parametrize("trick_limit", [4000, 100000])
def test_batched_images_equals_per_image_calls(trick_limit):
    """cddmsl_nms_batched (padded [B,M], counts on the device) is bit-identical, image by image, to cddmsl_nms on
    that image's first counts[b] boxes -- ragged counts, an empty image, several classes, both class-handling modes."""
    import importlib

    lnms = importlib.import_module("cddmsl_b200.layers.nms")   # (`cddmsl_b200.layers.nms` the attribute is the function)
    from cddmsl_b200.layers import batched_nms, batched_nms_images

    g = synth.generator(33)
    nb, m = 5, 700
    counts = [700, 1, 0, 333, 64]
    boxes = torch.stack([synth.make_boxes(m, 600, 1000, g) for _ in range(nb)])
    boxes[1] += 3000.0            # a different max coordinate per image (coordinate-trick offset is per image)
    scores = torch.randn(nb, m, generator=g).round(decimals=1)      # plenty of exact ties
    idxs = torch.randint(0, 3, (nb, m), generator=g)
    old = lnms.COORD_TRICK_NUMEL_LIMIT
    lnms.COORD_TRICK_NUMEL_LIMIT = trick_limit
    try:
        keep, nk = batched_nms_images(boxes.to(DEV), scores.to(DEV), idxs.to(DEV),
                                      torch.tensor(counts, dtype=torch.int32, device=DEV), 0.5)
        nk = nk.tolist()
        for b in range(nb):
            c = counts[b]
            # same mode as the batched call (rule evaluated on the padded size)
            lnms.COORD_TRICK_NUMEL_LIMIT = 10 ** 9 if m * 4 <= trick_limit else 0
            want = batched_nms(boxes[b, :c].to(DEV), scores[b, :c].to(DEV), idxs[b, :c].to(DEV), 0.5)
            lnms.COORD_TRICK_NUMEL_LIMIT = trick_limit
            assert nk[b] == want.numel()
            assert torch.equal(keep[b, : nk[b]], want)
    finally:
        lnms.COORD_TRICK_NUMEL_LIMIT = old
    # B = 1 with a device-side count, no class ids
    keep, nk = batched_nms_images(boxes[:1].to(DEV), scores[:1].to(DEV), None,
                                  torch.tensor([500], dtype=torch.int32, device=DEV), 0.5)
    want = batched_nms(boxes[0, :500].to(DEV), scores[0, :500].to(DEV), torch.zeros(500, dtype=torch.int64, device=DEV), 0.5)
    assert torch.equal(keep[0, : int(nk[0])], want)


def test_find_top_rpn_proposals_batched_equals_loop():
    """The all-images-at-once evaluation (one NMS call, one sync) returns exactly what the upstream-shaped loop does,
    including non-finite boxes dropped at inference and the min_box_size filter."""
    import cddmsl_b200.modeling.proposal_utils as pu

    g = synth.generator(34)
    n_img, a = 3, 3000
    props = torch.stack([synth.make_boxes(a, 700, 1100, g, degenerate_frac=0.05) - 40.0 for _ in range(n_img)])
    logits = torch.randn(n_img, a, generator=g)
    props[1, 5, 1] = float("nan")
    props[2, 9, 3] = float("inf")
    logits[0, 17] = float("-inf")
    sizes = [(600, 1000), (580, 990), (600, 800)]
    args = ([props.to(DEV)], [logits.to(DEV)], sizes, 0.7, 2000, 500, 4.0, False)
    batched = pu.find_top_rpn_proposals(*args)
    pu.BATCHED_IMAGES = False
    try:
        loop = pu.find_top_rpn_proposals(*args)
    finally:
        pu.BATCHED_IMAGES = True
    for rb, rl in zip(batched, loop):
        assert torch.equal(rb.proposal_boxes.tensor, rl.proposal_boxes.tensor)
        assert torch.equal(rb.objectness_logits, rl.objectness_logits)
        assert len(rb.objectness_logits) > 0


@pytest.mark.parametrize("hw_feat,img", [((38, 63), (600, 1000)), ((64, 128), (1024, 2048))])
def test_predict_proposals_fused_decode(hw_feat, img):
    """rpn.py:482-533 + box_regression.py:77-117 + proposal_utils.py:95-130 on 35 910 / 122 880 anchors per image:
    the fused decode + clip + non-empty + compaction kernel in front of the batched NMS against (a) the
    reference-shaped sequence of torch ops on the same device -- bit-exact -- and (b) the CPU oracle."""
    import cddmsl_b200.modeling.proposal_utils as pu
    from cddmsl_b200.modeling import Box2BoxTransform, predict_proposals
    from oracle import torch_ref

    g = synth.generator(91)
    n_img, na = 3, hw_feat[0] * hw_feat[1] * 15
    # anchors: 15 per location (3 ratios x 5 sizes) around the cell centres, some hanging over the border
    ys, xs = torch.meshgrid(torch.arange(hw_feat[0]) * 16.0, torch.arange(hw_feat[1]) * 16.0, indexing="ij")
    ctr = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4)
    sizes = torch.tensor([32.0, 64.0, 128.0, 256.0, 512.0])
    ratios = torch.tensor([0.5, 1.0, 2.0])
    w_ = (sizes[None, :] / ratios[:, None].sqrt()).reshape(-1)
    h_ = (sizes[None, :] * ratios[:, None].sqrt()).reshape(-1)
    cell = torch.stack([-w_ / 2, -h_ / 2, w_ / 2, h_ / 2], -1)[None]
    anchors = (ctr + cell).reshape(-1, 4).contiguous()
    assert anchors.shape[0] == na
    deltas = torch.randn(n_img, na, 4, generator=g) * torch.tensor([1.0, 1.0, 2.0, 2.0])
    deltas[0, :50, 2] = 40.0                       # beyond the scale clamp
    logits = torch.randn(n_img, na, generator=g)
    b2b = Box2BoxTransform((1.0, 1.0, 1.0, 1.0))
    args = dict(box2box_transform=b2b, nms_thresh=0.7, pre_nms_topk=12000, post_nms_topk=2000, min_box_size=0.0,
                training=True)
    fused = predict_proposals([anchors.to(DEV)], [logits.to(DEV)], [deltas.to(DEV)], [img] * n_img, **args)
    props = pu.decode_proposals([anchors.to(DEV)], [deltas.to(DEV)], b2b)
    eager = pu.find_top_rpn_proposals(props, [logits.to(DEV)], [img] * n_img, 0.7, 12000, 2000, 0.0, True)
    for f, e in zip(fused, eager):
        assert torch.equal(f.proposal_boxes.tensor, e.proposal_boxes.tensor)
        assert torch.equal(f.objectness_logits, e.objectness_logits)
    # CPU oracle (torch CPU exp may differ from the device's by an ulp: boxes to 1e-5, same kept count +- borderline)
    cpu_props = pu.decode_proposals([anchors], [deltas], b2b)[0]
    for i in range(n_img):
        lg, idx = logits[i].sort(descending=True)
        wb, ws = torch_ref.find_top_rpn_proposals_single_image(cpu_props[i][idx[:12000]], lg[:12000], img, 0.7, 2000)
        gb = fused[i].proposal_boxes.tensor.cpu()
        assert abs(len(gb) - len(wb)) <= 2
        m = min(len(gb), len(wb), 200)            # the head of the list is far from any tie
        assert torch.allclose(gb[:m], wb[:m], rtol=1e-5, atol=1e-3)
        assert torch.equal(fused[i].objectness_logits.cpu()[:m], ws[:m])
    bad = deltas.clone()
    bad[1, int(logits[1].argmax()), 0] = float("nan")
    with pytest.raises(FloatingPointError):
        predict_proposals([anchors.to(DEV)], [logits.to(DEV)], [bad.to(DEV)], [img] * n_img, **args)
    args["training"] = False                      # inference: the non-finite candidate is dropped silently
    out = predict_proposals([anchors.to(DEV)], [logits.to(DEV)], [bad.to(DEV)], [img] * n_img, **args)
    assert torch.isfinite(out[1].proposal_boxes.tensor).all()
