"""GPU parity: caption-consistency alignment loss (rcnn.py:305-317, :455-468) incl. the GatherLayer gradient
semantics.  No reference test exists upstream (parity unpinned); the multi-rank fixture was produced by the
reference's own GatherLayer under a real 2-process group (tests/golden/make_golden.py)."""
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
    assert ok.all(), f"{what}: max abs err {err.max():.3e} (scale {scale:.3e}), {(~ok).sum()}/{ok.size} outside"


def test_single_rank_fixtures(golden_dir):
    from cddmsl_b200.modeling import caption_consistency_loss

    a = np.load(os.path.join(golden_dir, "align_single.npz"))
    for tag in ("n16", "n40", "n256"):
        s = torch.from_numpy(a[f"a_{tag}"]).to(DEV).requires_grad_(True)
        t = torch.from_numpy(a[f"b_{tag}"]).to(DEV).requires_grad_(True)
        loss = caption_consistency_loss(s, t)
        (loss * 1.0).backward()
        _close(loss.item(), a[f"loss_{tag}"], 1e-5, "loss")
        _close(s.grad.cpu().numpy(), a[f"da_{tag}"], 1e-4, "da")
        _close(t.grad.cpu().numpy(), a[f"db_{tag}"], 1e-4, "db")


def test_reference_rcnn_lines_fixture(golden_dir):
    """align_ref.npz: the literal lines rcnn.py:455-470 (region), :305-319 (image) and :270-272 (KD) executed on
    seeded tensors, single process and with the reference's GatherLayer under a 2-rank gloo group."""
    from cddmsl_b200.modeling import caption_consistency_loss, image_caption_consistency_loss, kd_l1_loss

    a = np.load(os.path.join(golden_dir, "align_ref.npz"))
    for tag in ("n16", "n48", "n256"):
        for kind, fn in (("region", caption_consistency_loss), ("image", image_caption_consistency_loss)):
            s = torch.from_numpy(a[f"a_{tag}"]).to(DEV).requires_grad_(True)
            t = torch.from_numpy(a[f"b_{tag}"]).to(DEV).requires_grad_(True)
            loss = fn(s, t)
            loss.backward()
            _close(loss.item(), a[f"{kind}_loss_{tag}"], 1e-5, f"{kind} loss")
            _close(s.grad.cpu().numpy(), a[f"{kind}_da_{tag}"], 1e-4, f"{kind} da")
            _close(t.grad.cpu().numpy(), a[f"{kind}_db_{tag}"], 1e-4, f"{kind} db")
    al = [torch.from_numpy(a["w2_a0"]), torch.from_numpy(a["w2_a1"])]
    bl = [torch.from_numpy(a["w2_b0"]), torch.from_numpy(a["w2_b1"])]
    for r in range(2):
        loss, da, db = _emulated_rank(al, bl, r)
        for kind in ("region", "image"):
            _close(loss.item(), a[f"w2_{kind}_loss"], 1e-5, "w2 loss")
            _close(da.cpu().numpy(), a[f"w2_{kind}_da{r}"], 1e-4, "w2 da")
            _close(db.cpu().numpy(), a[f"w2_{kind}_db{r}"], 1e-4, "w2 db")
    # KD regulariser: loss within 1e-5, gradient bit-exact (sign / numel)
    t = torch.from_numpy(a["kd_teacher"]).to(DEV)
    s = torch.from_numpy(a["kd_student"]).to(DEV).requires_grad_(True)
    kl = kd_l1_loss(t, s)
    (kl * 1.0).backward()
    _close(kl.item(), a["kd_loss"], 1e-5, "kd loss")
    assert np.array_equal(s.grad.cpu().numpy(), a["kd_dstudent"])
    s2 = torch.from_numpy(a["kd_student"]).to(DEV).requires_grad_(True)
    (kd_l1_loss(t, s2) * 0.25).backward()
    assert np.array_equal(s2.grad.cpu().numpy(), a["kd_dstudent"] * np.float32(0.25))
    with torch.no_grad():
        _close(kd_l1_loss(t, s).item(), a["kd_loss"], 1e-5, "kd loss (no grad)")
    big = torch.randn(70000, 7, device=DEV)   # more elements than one grid pass
    ref = (big - big.roll(1, 0)).abs().double().mean().item()
    _close(kd_l1_loss(big.roll(1, 0), big).item(), ref, 1e-5, "kd loss, grid-stride")


def test_pair_api_equals_two_single_calls():
    """`caption_consistency_losses` (one all-gather for the image- and the region-level branch) == the two calls."""
    from cddmsl_b200.modeling import (caption_consistency_loss, caption_consistency_losses,
                                      image_caption_consistency_loss)

    g = synth.generator(7)
    ti, si = torch.randn(16, 256, generator=g), torch.randn(16, 256, generator=g)
    sr, tr = torch.randn(256, 256, generator=g), torch.randn(256, 256, generator=g)
    a = [t.to(DEV).requires_grad_(True) for t in (ti, si, sr, tr)]
    b = [t.to(DEV).requires_grad_(True) for t in (ti, si, sr, tr)]
    l1, l2 = caption_consistency_losses(*a)
    (l1 * 0.5 + l2 * 2.0).backward()
    m1, m2 = image_caption_consistency_loss(b[0], b[1]), caption_consistency_loss(b[2], b[3])
    (m1 * 0.5 + m2 * 2.0).backward()
    assert torch.equal(l1, m1) and torch.equal(l2, m2)
    for x, y in zip(a, b):
        assert torch.equal(x.grad, y.grad)


def test_upstream_gradient_scale_is_applied():
    from cddmsl_b200.modeling import caption_consistency_loss, image_caption_consistency_loss

    g = synth.generator(1)
    a, b = torch.randn(24, 256, generator=g), torch.randn(24, 256, generator=g)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    (torch_ref.caption_consistency_loss(ar, br) * 0.37).backward()
    s, t = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    (caption_consistency_loss(s, t) * 0.37).backward()
    _close(s.grad.cpu().numpy(), ar.grad.numpy(), 1e-4)
    _close(t.grad.cpu().numpy(), br.grad.numpy(), 1e-4)
    # image-level operand order (rcnn.py:313: joint = trgt @ src.T)
    l1 = image_caption_consistency_loss(b.to(DEV), a.to(DEV))
    _close(l1.item(), torch_ref.caption_consistency_loss(b, a).item(), 1e-5)


def _emulated_rank(a_locals, b_locals, rank):
    """What rank `rank` computes, with the all-gather replaced by concatenating every rank's packed buffer on
    one device (1 GPU cannot host 2 NCCL ranks)."""
    from cddmsl_b200 import ops

    packs, norms = [], None
    for r, (a, b) in enumerate(zip(a_locals, b_locals)):
        p, n = ops.align_pack(a.to(DEV), b.to(DEV))
        packs.append(p)
        if r == rank:
            norms = n
    packed_all = torch.stack(packs, 0)
    loss, da, db = ops.align_loss(packed_all, norms, rank, None, True)
    return loss, da, db


def test_two_rank_fixture_from_real_gatherlayer(golden_dir):
    w = np.load(os.path.join(golden_dir, "align_world2.npz"))
    a = [torch.from_numpy(w["a0"]), torch.from_numpy(w["a1"])]
    b = [torch.from_numpy(w["b0"]), torch.from_numpy(w["b1"])]
    for r in range(2):
        loss, da, db = _emulated_rank(a, b, r)
        _close(loss.item(), w["loss"], 1e-5, "loss")
        _close(da.cpu().numpy(), w[f"da{r}"], 1e-4, f"da{r}")
        _close(db.cpu().numpy(), w[f"db{r}"], 1e-4, f"db{r}")


@pytest.mark.parametrize("world,n_local,dim", [(8, 256, 256), (4, 32, 256), (3, 5, 64), (2, 16, 1024),
                                               (2, 9, 1300), (1, 3, 30), (2, 600, 64)])
def test_world_emulation_vs_oracle(world, n_local, dim):
    """configs[2]: 8 ranks x 16 regions x 16 images = 2048 gathered rows."""
    g = synth.generator(world * 100 + n_local)
    a = [torch.randn(n_local, dim, generator=g) for _ in range(world)]
    b = [torch.randn(n_local, dim, generator=g) for _ in range(world)]
    ranks = [0, world - 1] if world > 3 else list(range(world))
    a_all, b_all = torch.cat(a), torch.cat(b)
    for r in ranks:
        ar = a_all.clone().requires_grad_(True)
        br = b_all.clone().requires_grad_(True)
        l_ref = torch_ref.caption_consistency_loss(ar, br)
        l_ref.backward()
        sl = slice(r * n_local, (r + 1) * n_local)
        loss, da, db = _emulated_rank(a, b, r)
        _close(loss.item(), l_ref.item(), 1e-5, "loss")
        _close(da.cpu().numpy(), ar.grad[sl].numpy(), 1e-4, "da")
        _close(db.cpu().numpy(), br.grad[sl].numpy(), 1e-4, "db")


def _nccl_worker(rank, world, port, golden_dir, q):
    import torch.distributed as dist

    from cddmsl_b200.modeling import caption_consistency_loss, image_caption_consistency_loss

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out = {}
        for name, fn, kind in (("align_world2.npz", caption_consistency_loss, ""),
                               ("align_ref.npz", caption_consistency_loss, "w2_region_"),
                               ("align_ref.npz", image_caption_consistency_loss, "w2_image_")):
            w = np.load(os.path.join(golden_dir, name))
            pre = "w2_" if kind else ""
            a = torch.from_numpy(w[f"{pre}a{rank}"]).to(dev).requires_grad_(True)
            b = torch.from_numpy(w[f"{pre}b{rank}"]).to(dev).requires_grad_(True)
            loss = fn(a, b)
            loss.backward()
            out[kind or "world2_"] = (loss.item(), a.grad.cpu().numpy(), b.grad.cpu().numpy())
        # both branches through ONE all-gather (image level: operands (trgt, src); region level: (src, tgt))
        from cddmsl_b200.modeling import caption_consistency_losses
        w = np.load(os.path.join(golden_dir, "align_ref.npz"))
        t = [torch.from_numpy(w[f"w2_{n}{rank}"]).to(dev).requires_grad_(True) for n in ("a", "b", "a", "b")]
        l_img, l_reg = caption_consistency_losses(*t)
        (l_img + l_reg).backward()
        out["pair_"] = (l_img.item(), l_reg.item(), [x.grad.cpu().numpy() for x in t])
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_nccl(golden_dir):
    """The real thing: two processes, two GPUs, NCCL all-gather inside `caption_consistency_loss`, against the fixtures
    the reference's own GatherLayer + rcnn.py lines produced under gloo."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`; the gloo host-logic test covers the N > 1 path on CPU)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, 29611, golden_dir, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    w1 = np.load(os.path.join(golden_dir, "align_world2.npz"))
    w2 = np.load(os.path.join(golden_dir, "align_ref.npz"))
    for r in range(2):
        loss, da, db = res[r]["world2_"]
        _close(loss, w1["loss"], 1e-5, "loss")
        _close(da, w1[f"da{r}"], 1e-4, f"da{r}")
        _close(db, w1[f"db{r}"], 1e-4, f"db{r}")
        for kind in ("w2_region_", "w2_image_"):
            loss, da, db = res[r][kind]
            _close(loss, w2[f"{kind}loss"], 1e-5, kind + "loss")
            _close(da, w2[f"{kind}da{r}"], 1e-4, kind + "da")
            _close(db, w2[f"{kind}db{r}"], 1e-4, kind + "db")
        l_img, l_reg, grads = res[r]["pair_"]
        _close(l_img, w2["w2_image_loss"], 1e-5, "pair image loss")
        _close(l_reg, w2["w2_region_loss"], 1e-5, "pair region loss")
        _close(grads[0], w2[f"w2_image_da{r}"], 1e-4, "pair image da")
        _close(grads[3], w2[f"w2_region_db{r}"], 1e-4, "pair region db")
