"""GPU parity: ROIAlign fwd/bwd through the reference-shaped API -> custom op -> C ABI -> sm_100a kernels,
against the oracle.  Tolerances (BASELINE.json north_star): forward 1e-5 relative, backward 1e-4 (the
accumulation order differs from the serial CPU kernel); `atol` scales with the magnitude of the data."""
import glob
import os

import numpy as np
import pytest
import torch

from cddmsl_b200 import synth
from oracle import c_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_RTOL, BWD_RTOL = 1e-5, 1e-4


def _close(got, want, rtol):
    scale = max(float(np.abs(want).max()), 1e-6)
    err = np.abs(got - want)
    ok = err <= rtol * np.abs(want) + rtol * scale
    assert ok.all(), f"max abs err {err.max():.3e} (scale {scale:.3e}), {(~ok).sum()} of {ok.size} outside rtol={rtol}"


def _run(feat, rois, p, scale, sr, aligned, gout=None):
    from cddmsl_b200.layers import ROIAlign

    x = torch.from_numpy(feat).to(DEV).requires_grad_(True)
    r = torch.from_numpy(rois).to(DEV)
    out = ROIAlign((p[0], p[1]), scale, sr, aligned=aligned)(x, r)
    gin = None
    if gout is not None:
        out.backward(torch.from_numpy(gout).to(DEV))
        gin = x.grad.cpu().numpy()
    return out.detach().cpu().numpy(), gin


def test_reference_golden_vectors():
    # tests/layers/test_roi_align.py:14-47 upstream
    inp = np.arange(25, dtype=np.float32).reshape(1, 1, 5, 5)
    roi = np.array([[0, 1, 1, 3, 3]], dtype=np.float32)
    old = [[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]]
    new = [[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]]
    o, _ = _run(inp, roi, (4, 4), 1.0, 0, False)
    assert np.allclose(o[0, 0], np.asarray(old))
    o, _ = _run(inp, roi, (4, 4), 1.0, 0, True)
    assert np.allclose(o[0, 0], np.asarray(new))


def test_reference_empty_box_and_empty_batch():
    # tests/layers/test_roi_align.py:111-128 upstream
    from cddmsl_b200.layers import ROIAlign

    img = np.random.RandomState(0).rand(1, 1, 5, 5).astype(np.float32)
    roi = np.array([[0, 3, 4, 5, 4]], dtype=np.float32)
    o, g = _run(img, roi, (7, 7), 1.0, 0, True, gout=np.ones((1, 1, 7, 7), np.float32))
    assert o.shape == (1, 1, 7, 7) and (o == 0).all() and (g == 0).all()
    out = ROIAlign((7, 7), 1.0, 0, aligned=True)(torch.zeros(0, 3, 10, 10, device=DEV), torch.zeros(0, 5, device=DEV))
    assert out.shape == (0, 3, 7, 7)


def test_fixtures_from_reference_wrapper(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "roi_align_*.npz")))
    assert len(files) >= 5
    for f in files:
        d = np.load(f)
        p, _, sr, al = (int(v) for v in d["meta"])
        o, g = _run(d["feat"], d["rois"], (p, p), float(d["scale"][0]), sr, bool(al), gout=d["gout"])
        _close(o, d["out"], FWD_RTOL)
        _close(g, d["gin"], BWD_RTOL)


@pytest.mark.parametrize("p,sr,aligned,c,hw", [
    (14, 0, True, 64, (38, 63)),    # the CDDMSL pooler configuration on a VOC-shaped map
    (7, 0, True, 40, (38, 63)),     # microbench sweep: 7x7 bins
    (14, 2, True, 33, (64, 128)),   # fixed sampling ratio, Cityscapes-shaped map, ragged channel chunk
    (7, 2, False, 8, (20, 30)),     # legacy (aligned=False) with the 1x1 minimum size
    (14, 0, True, 3, (150, 300)),   # footprints beyond the shared-memory budget -> global-memory path
    (28, 0, True, 5, (38, 63)),     # > 16 bins per side -> global-memory path
])
def test_seeded_random_vs_oracle(p, sr, aligned, c, hw):
    g = synth.generator(100 + p + sr)
    n, r = 3, 96
    h, w = hw
    feat = torch.randn(n, c, h, w, generator=g).numpy()
    parts = []
    for i in range(n):
        b = synth.make_boxes(r // n, h * 16, w * 16, g, degenerate_frac=0.05)
        b[0] = torch.tensor([0.0, 0.0, w * 16.0, h * 16.0])                 # whole image
        b[1] = torch.tensor([-40.0, -60.0, 90.0, 70.0])                      # over-hanging top-left
        b[2] = torch.tensor([w * 16 - 50.0, h * 16 - 30.0, w * 16 + 80.0, h * 16 + 90.0])  # over-hanging bottom-right
        parts.append(torch.cat([torch.full((r // n, 1), float(i)), b], 1))
    rois = torch.cat(parts).numpy()
    gout = torch.randn(rois.shape[0], c, p, p, generator=g).numpy()
    o, gin = _run(feat, rois, (p, p), 1.0 / 16, sr, aligned, gout=gout)
    _close(o, c_ref.roi_align_fwd(feat, rois, (p, p), 1.0 / 16, sr, aligned), FWD_RTOL)
    _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, sr, aligned), BWD_RTOL)


@pytest.mark.parametrize("knobs", [
    {"roi_pr": 0},                      # round-1 channels-last kernels (forward fallback / default backward)
    {"roi_pr": 0, "roi_use_cl": 0},     # generic NCHW kernels (what runs when no workspace is passed)
    {"roi_pr": 3},                      # plane-resident forward AND the transposed backward (experimental knob)
    {"roi_pr": 1, "roi_pr_chunk": 32},  # many small units per image
], ids=["channels_last", "generic_nchw", "pr_fwd_bwd", "pr_small_chunks"])
def test_every_kernel_family_on_the_14x14_shape(knobs):
    """The 14x14 pooler normally takes the plane-resident forward and the channels-last backward; force every other
    kernel family through the tuning knobs and hold it to the same oracle."""
    from cddmsl_b200 import _lib

    g = synth.generator(55)
    feat = torch.randn(2, 48, 38, 63, generator=g).numpy()
    rois = synth.make_rois(synth.PathConfig("t", 2, 600, 1000, 40, 5), g).numpy()
    gout = torch.randn(rois.shape[0], 48, 14, 14, generator=g).numpy()
    defaults = {"roi_pr": 1, "roi_use_cl": 1, "roi_pr_chunk": 0}
    for k, v in knobs.items():
        assert _lib.tune(k, v)
    try:
        o, gin = _run(feat, rois, (14, 14), 1.0 / 16, 0, True, gout=gout)
    finally:
        for k in knobs:
            _lib.tune(k, defaults[k])
    _close(o, c_ref.roi_align_fwd(feat, rois, (14, 14), 1.0 / 16, 0, True), FWD_RTOL)
    _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, 0, True), BWD_RTOL)


@pytest.mark.parametrize("channels,sr,aligned", [(64, 0, True), (96, 0, False), (32, 2, True), (160, 1, True),
                                                 (64, 2, False), (64, 3, True)])
def test_row_walk_backward(channels, sr, aligned):
    """roi_align_rw.cu (default backward for 14x14, C % 32 == 0): random boxes of every size, boxes hanging over all
    four borders, degenerate and whole-map boxes, a box on an out-of-range image index, RoIs with > 10 samples per bin
    (reference-shaped slow path) and fixed sampling grids that are not monotone (slow path); against the oracle and
    against the round-1 channels-last kernel."""
    from cddmsl_b200 import _lib

    g = synth.generator(58 + channels)
    shape = (2, channels, 38, 63)
    rois = synth.make_rois(synth.PathConfig("t", 2, 600, 1000, 48, 5), g).numpy()
    extra = np.array([[0, -200, -100, 1300, 700], [1, 0, 0, 1008, 608], [0, 990, 590, 1100, 700], [1, -50, 300, 30, 330],
                      [0, 500, -40, 530, 20], [1, 100, 100, 100, 100], [0, 3, 3, 5, 600], [1, 2, 2, 1000, 9],
                      [0, 0, 0, 16, 16], [1, 1007, 607, 1008, 608], [5, 10, 10, 200, 200],
                      [0, -3000, -3000, 4000, 4000]], dtype=np.float32)
    rois = np.concatenate([rois, extra]).astype(np.float32)
    gout = torch.randn(rois.shape[0], channels, 14, 14, generator=g).numpy()
    # an image index outside the batch is undefined behaviour in the reference (and in the oracle): the kernels skip
    # such a RoI, so the oracle sees the list without it
    valid = rois[:, 0] < shape[0]
    want = c_ref.roi_align_bwd(gout[valid], rois[valid], shape, 1.0 / 16, sr, aligned)
    feat = np.zeros(shape, np.float32)
    assert _lib.tune("roi_rw_min_units", 0)   # small lists normally stay on the channels-last kernel
    try:
        _, gin = _run(feat, rois, (14, 14), 1.0 / 16, sr, aligned, gout=gout)
        _close(gin, want, BWD_RTOL)
        if channels % 64 == 0:
            assert _lib.tune("roi_rw_cpl", 2)     # two channels per lane (knob)
            _, gin2 = _run(feat, rois, (14, 14), 1.0 / 16, sr, aligned, gout=gout)
            _lib.tune("roi_rw_cpl", 1)
            _close(gin2, want, BWD_RTOL)
        assert _lib.tune("roi_rw", 0)
        _, gin_cl = _run(feat, rois, (14, 14), 1.0 / 16, sr, aligned, gout=gout)
    finally:
        _lib.tune("roi_rw", 1)
        _lib.tune("roi_rw_cpl", 1)
        _lib.tune("roi_rw_min_units", 65536)
    _close(gin_cl, want, BWD_RTOL)
    assert not np.array_equal(gin, gin_cl)   # two different kernels ran


def test_row_walk_backward_fuzz_against_channels_last():
    """40 random shapes (maps down to 1 x 1, C = 32 ... 128, any sampling ratio, legacy and aligned coordinates, boxes
    far outside the map, zero-area boxes): the row-walk backward and the round-1 channels-last backward are two
    independent implementations of the same sum and have to agree to 1e-4 of the gradient scale."""
    from cddmsl_b200 import _lib, ops

    rng = np.random.RandomState(7)
    g = synth.generator(99)
    for trial in range(40):
        n = int(rng.randint(1, 4))
        c = int(rng.choice([32, 64, 96, 128]))
        h, w = int(rng.choice([1, 2, 3, 7, 20, 38, 50])), int(rng.choice([1, 2, 5, 16, 63, 90]))
        r = int(rng.randint(1, 60))
        sr = int(rng.choice([0, 0, 0, 1, 2, 3]))
        aligned = bool(rng.randint(0, 2))
        cx, cy = rng.uniform(-0.3, 1.3, r) * w * 16, rng.uniform(-0.3, 1.3, r) * h * 16
        bw = np.exp(rng.uniform(np.log(1.0), np.log(max(w * 16 * 1.5, 2.0)), r))
        bh = np.exp(rng.uniform(np.log(1.0), np.log(max(h * 16 * 1.5, 2.0)), r))
        rois = np.stack([rng.randint(0, n, r).astype(np.float64), cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        zero = rng.rand(r) < 0.06
        rois[zero, 3] = rois[zero, 1]                      # zero width
        rois = torch.from_numpy(rois.astype(np.float32)).to(DEV)
        gout = torch.randn(r, c, 14, 14, generator=g).to(DEV)
        _lib.tune("roi_rw_min_units", 0)
        try:
            a = ops.roi_align_backward(gout, rois, 1.0 / 16, 14, 14, n, c, h, w, sr, aligned)
            _lib.tune("roi_rw", 0)
            b = ops.roi_align_backward(gout, rois, 1.0 / 16, 14, 14, n, c, h, w, sr, aligned)
        finally:
            _lib.tune("roi_rw", 1)
            _lib.tune("roi_rw_min_units", 65536)
        scale = max(float(b.abs().max()), 1e-6)
        err = float((a - b).abs().max())
        assert err <= 1e-4 * scale, (trial, n, c, h, w, r, sr, aligned, err, scale)


def test_wide_bands_and_sparse_sampling_grids():
    """Plane-resident classes beyond the common ones: RoIs wider than 6 cells per bin (B halves of the records), a
    fixed sampling grid on huge bins (per-sample path), boxes hanging over every border, on a map with H*W odd."""
    g = synth.generator(56)
    feat = torch.randn(1, 6, 33, 131, generator=g).numpy()
    rois = np.array([[0, 0, 0, 2090, 520], [0, -300, -200, 2400, 700], [0, 5, 5, 1900, 40], [0, 10, 10, 30, 500],
                     [0, 100, 100, 101, 101], [0, 2000, 400, 2300, 600], [0, 700, 100, 1500, 420]], dtype=np.float32)
    gout = torch.randn(rois.shape[0], 6, 14, 14, generator=g).numpy()
    for sr in (0, 1, 2):
        o, gin = _run(feat, rois, (14, 14), 1.0 / 16, sr, True, gout=gout)
        _close(o, c_ref.roi_align_fwd(feat, rois, (14, 14), 1.0 / 16, sr, True), FWD_RTOL)
        _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, sr, True), BWD_RTOL)


def test_dual_map_equals_two_single_calls():
    """cddmsl_roi_align_fwd2 / bwd2 (clip_roi_heads.py:117-132: source and target maps, identical boxes): bit-equal
    to two single calls, gradients flow to both maps."""
    from cddmsl_b200 import ops
    from cddmsl_b200.layers import ROIAlign

    g = synth.generator(57)
    fa = torch.randn(3, 40, 38, 63, generator=g).to(DEV).requires_grad_(True)
    fb = torch.randn(3, 40, 38, 63, generator=g).to(DEV).requires_grad_(True)
    rois = synth.make_rois(synth.PathConfig("t", 3, 600, 1000, 16, 5), g).to(DEV)
    op = ROIAlign((14, 14), 1.0 / 16, 0, aligned=True)
    oa, ob = op.forward_pair(fa, fb, rois)
    assert torch.equal(oa, op(fa, rois)) and torch.equal(ob, op(fb, rois))
    ga, gb = torch.randn_like(oa), torch.randn_like(ob)
    torch.autograd.backward([oa, ob], [ga, gb])
    ra = ops.roi_align_backward(ga, rois, 1.0 / 16, 14, 14, 3, 40, 38, 63, 0, True)
    rb = ops.roi_align_backward(gb, rois, 1.0 / 16, 14, 14, 3, 40, 38, 63, 0, True)
    _close(fa.grad.cpu().numpy(), ra.cpu().numpy(), BWD_RTOL)   # (atomic order differs between two launches)
    _close(fb.grad.cpu().numpy(), rb.cpu().numpy(), BWD_RTOL)
    # a 7x7 pooler on few channels goes through the same entry point
    op7 = ROIAlign((7, 7), 1.0 / 16, 2, aligned=True)
    pa, pb = op7.forward_pair(fa.detach()[:, :3].contiguous(), fb.detach()[:, :3].contiguous(), rois)
    assert torch.equal(pa, op7(fa.detach()[:, :3].contiguous(), rois))
    assert torch.equal(pb, op7(fb.detach()[:, :3].contiguous(), rois))


def test_clip_res5_roi_heads_forward_get_features():
    """CLIPRes5ROIHeads.forward_get_features (clip_roi_heads.py:117-132): pooler -> res5 -> attnpool on the source
    and the target map with the same proposal boxes, against the oracle's ROIAlign + the same torch modules."""
    from cddmsl_b200.modeling import Box2BoxTransform, CLIPRes5ROIHeads, FastRCNNOutputLayers, ROIPooler
    from cddmsl_b200.structures import Boxes, Instances

    g = synth.generator(58)
    c, k = 24, 5
    torch.manual_seed(3)
    res5 = torch.nn.Sequential(torch.nn.Conv2d(c, 16, 3, stride=2, padding=1), torch.nn.ReLU()).to(DEV)
    attnpool = lambda x: x.mean(dim=(2, 3))
    feats_s = torch.randn(2, c, 38, 63, generator=g)
    feats_t = torch.randn(2, c, 38, 63, generator=g)
    boxes = [synth.make_boxes(16, 600, 1000, g, degenerate_frac=0.0) for _ in range(2)]
    props = []
    for b in boxes:
        inst = Instances((600, 1000))
        inst.proposal_boxes = Boxes(b.to(DEV))
        props.append(inst)
    w = torch.randn(k, 16, generator=g)
    head = CLIPRes5ROIHeads(
        in_features=["res4"], pooler=ROIPooler(14, (1.0 / 16,), 0, "ROIAlignV2"), num_classes=k,
        box_predictor=FastRCNNOutputLayers(16, box2box_transform=Box2BoxTransform((10.0, 10.0, 5.0, 5.0)),
                                           num_classes=k, clip_cls_emb=(True, w, "CLIPRes5ROIHeads", 16),
                                           bg_cls_loss_weight=0.2, openset_test=(None, None, 0.01, 0.5)).to(DEV))
    head.eval()
    a_s, a_t = head.forward_get_features({"res4": feats_s.to(DEV)}, {"res4": feats_t.to(DEV)}, props, res5=res5,
                                         attnpool=attnpool)
    rois = torch.cat([torch.cat([torch.full((16, 1), float(i)), b], 1) for i, b in enumerate(boxes)]).numpy()
    for got, f in ((a_s, feats_s), (a_t, feats_t)):
        pooled = torch.from_numpy(c_ref.roi_align_fwd(f.numpy(), rois, (14, 14), 1.0 / 16, 0, True)).to(DEV)
        with torch.no_grad():
            want = attnpool(res5(pooled))
        _close(got.detach().cpu().numpy(), want.cpu().numpy(), 1e-4)
    # the training / inference `forward` of the same head (box branch) runs end to end
    tg = []
    for b in boxes:
        t = Instances((600, 1000))
        t.gt_boxes = Boxes(b[:3].to(DEV))
        t.gt_classes = torch.tensor([0, 1, 2], device=DEV)
        tg.append(t)
    for p in props:
        p.objectness_logits = torch.zeros(len(p), device=DEV)
    head.train()
    _, losses = head.forward(None, {"res4": feats_s.to(DEV)}, props, tg, res5=res5, attnpool=attnpool)
    assert set(losses) == {"loss_cls", "loss_box_reg"} and all(torch.isfinite(v).all() for v in losses.values())
    head.eval()
    inst, _ = head.forward(None, {"res4": feats_s.to(DEV)}, props, None, res5=res5, attnpool=attnpool)
    assert len(inst) == 2 and all(i.has("pred_boxes") and i.has("scores") for i in inst)


def test_rois_in_arbitrary_image_order():
    g = synth.generator(5)
    feat = torch.randn(4, 16, 20, 30, generator=g).numpy()
    rois = synth.make_rois(synth.PathConfig("t", 4, 320, 480, 10, 5), g).numpy()
    perm = torch.randperm(rois.shape[0], generator=g).numpy()
    rois = rois[perm]
    gout = torch.randn(rois.shape[0], 16, 14, 14, generator=g).numpy()
    o, gin = _run(feat, rois, (14, 14), 1.0 / 16, 0, True, gout=gout)
    _close(o, c_ref.roi_align_fwd(feat, rois, (14, 14), 1.0 / 16, 0, True), FWD_RTOL)
    _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, 0, True), BWD_RTOL)


def test_config1_slice_vs_oracle():
    """BASELINE.json configs[0] shape (2 images, 38x63 map, 512 RoIs/img) on a 128-channel slice so the CPU
    oracle finishes in seconds."""
    cfg = synth.CONFIGS["cpu_ref"]
    g = synth.generator(cfg.seed)
    feat = synth.make_features(cfg, g)[:, :128].contiguous().numpy()
    rois = synth.make_rois(cfg, g).numpy()
    gout = torch.randn(rois.shape[0], 128, 14, 14, generator=g).numpy()
    o, gin = _run(feat, rois, (14, 14), 1.0 / 16, 0, True, gout=gout)
    _close(o, c_ref.roi_align_fwd(feat, rois, (14, 14), 1.0 / 16, 0, True), FWD_RTOL)
    _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, 0, True), BWD_RTOL)


def test_full_size_properties_voc():
    """configs[1] at full size (16 x 1024 x 38 x 63, 8192 RoIs -> 6.6 GB out): size-independent properties.
    (a) adjointness <A x, g> == <x, A^T g>; (b) linearity; (c) a constant map pools to the constant for
    interior RoIs; (d) a 64-channel slice equals the oracle."""
    from cddmsl_b200.layers import ROIAlign

    cfg = synth.CONFIGS["voc"]
    g = synth.generator(cfg.seed)
    feat = synth.make_features(cfg, g).to(DEV)
    rois = synth.make_rois(cfg, g).to(DEV)
    op = ROIAlign((14, 14), 1.0 / 16, 0, aligned=True)
    x = feat.clone().requires_grad_(True)
    out = op(x, rois)
    assert out.shape == (cfg.n_rois, 1024, 14, 14)
    gout = torch.empty_like(out).normal_(generator=None)
    out.backward(gout)
    lhs = torch.dot(out.detach().flatten().double(), gout.flatten().double()).item()
    rhs = torch.dot(feat.flatten().double(), x.grad.flatten().double()).item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs)) + 1e-2, (lhs, rhs)
    # (d) oracle on a channel slice
    sl = slice(500, 564)
    want = c_ref.roi_align_fwd(feat[:, sl].cpu().numpy(), rois.cpu().numpy(), (14, 14), 1.0 / 16, 0, True)
    _close(out.detach()[:, sl].cpu().numpy(), want, FWD_RTOL)
    gw = c_ref.roi_align_bwd(gout[:, sl].cpu().numpy(), rois.cpu().numpy(), (16, 64, 38, 63), 1.0 / 16, 0, True)
    _close(x.grad[:, sl].cpu().numpy(), gw, BWD_RTOL)
    del gout, x
    # (b) linearity on a 256-channel slice
    a = feat[:, :256].contiguous()
    b = torch.randn_like(a)
    with torch.no_grad():
        l1 = op(2.5 * a + b, rois)
        l2 = 2.5 * op(a, rois) + op(b, rois)
    assert torch.allclose(l1, l2, rtol=1e-4, atol=1e-4)
    # (c) constant map
    with torch.no_grad():
        ones = op(torch.full((16, 8, 38, 63), 3.0, device=DEV), rois)
    r = rois.cpu()
    interior = (r[:, 1] > 32) & (r[:, 2] > 32) & (r[:, 3] < 1000 - 32) & (r[:, 4] < 600 - 32) & \
        (r[:, 3] - r[:, 1] > 1) & (r[:, 4] - r[:, 2] > 1)
    assert interior.sum() > 1000
    assert torch.allclose(ones[interior.to(DEV)], torch.tensor(3.0, device=DEV), rtol=1e-5, atol=1e-5)


def test_roi_pooler_single_level_and_empty():
    # tests/modeling/test_roi_pooler.py:110-118 upstream + the C4 fast path (poolers.py:228-229)
    from cddmsl_b200.modeling import ROIPooler
    from cddmsl_b200.structures import Boxes

    g = synth.generator(3)
    feat = torch.randn(2, 8, 20, 30, generator=g)
    boxes = [synth.make_boxes(7, 320, 480, g), synth.make_boxes(5, 320, 480, g)]
    pooler = ROIPooler((14, 14), (1.0 / 16,), 0, "ROIAlignV2")
    out = pooler([feat.to(DEV)], [Boxes(b.to(DEV)) for b in boxes])
    rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b], 1) for i, b in enumerate(boxes)])
    _close(out.cpu().numpy(), c_ref.roi_align_fwd(feat.numpy(), rois.numpy(), (14, 14), 1.0 / 16, 0, True), FWD_RTOL)
    empty = pooler([torch.zeros(0, 8, 20, 30, device=DEV)], [])
    assert empty.shape == (0, 8, 14, 14)
