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


def test_generic_nchw_kernels_on_the_14x14_shape():
    """The 14x14 pooler normally takes the channels-last fast path; force the generic NCHW kernels (what runs
    when no workspace is passed through the C ABI) and hold them to the same oracle."""
    from cddmsl_b200 import _lib

    g = synth.generator(55)
    feat = torch.randn(2, 48, 38, 63, generator=g).numpy()
    rois = synth.make_rois(synth.PathConfig("t", 2, 600, 1000, 40, 5), g).numpy()
    gout = torch.randn(rois.shape[0], 48, 14, 14, generator=g).numpy()
    assert _lib.tune("roi_use_cl", 0)
    try:
        o, gin = _run(feat, rois, (14, 14), 1.0 / 16, 0, True, gout=gout)
    finally:
        _lib.tune("roi_use_cl", 1)
    _close(o, c_ref.roi_align_fwd(feat, rois, (14, 14), 1.0 / 16, 0, True), FWD_RTOL)
    _close(gin, c_ref.roi_align_bwd(gout, rois, feat.shape, 1.0 / 16, 0, True), BWD_RTOL)


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
