"""GPU parity: RegionCLIP pretraining losses (clip_rcnn.py:590-611 region-concept distillation + MIL contrastive,
:624-640 image-text matching with comm.py:268-322 gather).  The fixture holds the reference's own lines executed on
seeded tensors (tests/golden/make_golden.py pretrain_ref_cases); larger shapes compare against the oracle.
Tolerance: loss 1e-5 relative, gradients 1e-4 of the gradient scale (fp32, 3xTF32 logits for K >= 255)."""
import os

import numpy as np
import pytest
import torch

from cddmsl_b200 import ops, synth
from cddmsl_b200.modeling import (concept_contrastive_loss, image_text_matching_loss,
                                  region_concept_distill_loss)
from oracle import torch_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, want, rtol, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(float(np.abs(want).max()), 1e-12)
    err = np.abs(got - want)
    ok = err <= rtol * np.abs(want) + rtol * scale
    assert ok.all(), f"{what}: max abs err {err.max():.3e} (scale {scale:.3e}), {(~ok).sum()}/{ok.size} outside"


def test_reference_clip_rcnn_lines_fixture(golden_dir):
    p = np.load(os.path.join(golden_dir, "pretrain_ref.npz"))
    for tag in ("small", "lvis"):
        t = lambda k: torch.from_numpy(p[f"{k}_{tag}"]).to(DEV)
        temp = float(p[f"temp_{tag}"][0])
        x = t("feats").requires_grad_(True)
        loss = region_concept_distill_loss(x, t("concept_emb"), t("teacher"), temp)
        loss.backward()
        _close(loss.item(), p[f"distill_{tag}"], 1e-5, f"distill {tag}")
        _close(x.grad.cpu().numpy(), p[f"distill_dx_{tag}"], 1e-4, f"distill dx {tag}")
        x = t("feats").requires_grad_(True)
        loss = concept_contrastive_loss(x, t("target_embs"), t("label_mtx"), temp)
        loss.backward()
        _close(loss.item(), p[f"contrastive_{tag}"], 1e-5, f"contrastive {tag}")
        _close(x.grad.cpu().numpy(), p[f"contrastive_dx_{tag}"], 1e-4, f"contrastive dx {tag}")
    a = torch.from_numpy(p["it_feats"]).to(DEV).requires_grad_(True)
    b = torch.from_numpy(p["it_text"]).to(DEV).requires_grad_(True)
    loss = image_text_matching_loss(a, b, float(p["it_temp"][0]))
    loss.backward()
    _close(loss.item(), p["it_loss"], 1e-5, "image-text")
    _close(a.grad.cpu().numpy(), p["it_dfeats"], 1e-4, "image-text dfeats")
    _close(b.grad.cpu().numpy(), p["it_dtext"], 1e-4, "image-text dtext")


@pytest.mark.parametrize("r,d,k,temp", [(300, 1024, 4764, 0.01), (7, 64, 3, 0.05), (513, 512, 255, 0.01),
                                        (1, 32, 2, 1.0)])
def test_region_concept_losses_vs_oracle(r, d, k, temp):
    """RegionCLIP pretraining shapes: 4764 concepts x 1024-d (configs of pretrain/RegionCLIP_RN50.yaml), plus ragged."""
    g = synth.generator(r + k)
    feats = torch.randn(r, d, generator=g)
    concept = torch.randn(k, d, generator=g)
    teacher = torch.softmax(torch.randn(r, k, generator=g) * 4.0, dim=1)
    idx = torch.randint(0, max(2, r // 3), (r,), generator=g)
    tgt = torch.randn(max(2, r // 3), d, generator=g)[idx]
    lab = (idx[:, None] == idx[None, :]).float()
    xr = feats.clone().requires_grad_(True)
    want_d, _ = torch_ref.region_concept_losses(xr, concept, teacher, tgt, lab, temp)
    gd, = torch.autograd.grad(want_d, xr)
    xr = feats.clone().requires_grad_(True)
    _, want_c = torch_ref.region_concept_losses(xr, concept, teacher, tgt, lab, temp)
    gc, = torch.autograd.grad(want_c, xr)

    x = feats.to(DEV).requires_grad_(True)
    loss = region_concept_distill_loss(x, concept.to(DEV), teacher.to(DEV), temp)
    loss.backward()
    _close(loss.item(), want_d.item(), 1e-5, "distill")
    _close(x.grad.cpu().numpy(), gd.numpy(), 1e-4, "distill dx")
    x = feats.to(DEV).requires_grad_(True)
    loss = concept_contrastive_loss(x, tgt.to(DEV), lab.to(DEV), temp)
    loss.backward()
    _close(loss.item(), want_c.item(), 1e-5, "contrastive")
    _close(x.grad.cpu().numpy(), gc.numpy(), 1e-4, "contrastive dx")


def test_losses_scale_with_upstream_gradient_and_skip_grad_under_no_grad():
    g = synth.generator(5)
    feats, concept = torch.randn(20, 64, generator=g).to(DEV), torch.randn(9, 64, generator=g).to(DEV)
    teacher = torch.softmax(torch.randn(20, 9, generator=g), 1).to(DEV)
    x1 = feats.clone().requires_grad_(True)
    region_concept_distill_loss(x1, concept, teacher, 0.01).backward()
    x2 = feats.clone().requires_grad_(True)
    (region_concept_distill_loss(x2, concept, teacher, 0.01) * 0.25).backward()
    _close(x2.grad.cpu().numpy(), 0.25 * x1.grad.cpu().numpy(), 1e-6, "scaled")
    with torch.no_grad():
        l0 = region_concept_distill_loss(feats, concept, teacher, 0.01)
    assert not l0.requires_grad


@pytest.mark.parametrize("world,n_local,dim,temp", [(8, 12, 1024, 0.01), (2, 5, 64, 0.07), (1, 24, 96, 0.07)])
def test_image_text_world_emulation_vs_oracle(world, n_local, dim, temp):
    """`gather_tensors` is diffdist's differentiable all-gather: every rank back-propagates the same full loss and the
    gradient of a rank's rows is the sum over ranks = world x its slice of d loss / d gathered."""
    g = synth.generator(world * 7 + n_local)
    a = [torch.randn(n_local, dim, generator=g) for _ in range(world)]
    b = [torch.randn(n_local, dim, generator=g) for _ in range(world)]
    ar, br = torch.cat(a).requires_grad_(True), torch.cat(b).requires_grad_(True)
    want = torch_ref.image_text_matching_loss(ar, br, temp)
    want.backward()
    packs = [ops.align_pack(x.to(DEV), y.to(DEV)) for x, y in zip(a, b)]
    packed_all = torch.stack([p for p, _ in packs])
    for r in sorted({0, world - 1}):
        loss, da, db = ops.contrastive_loss(packed_all, packs[r][1], r, 1.0 / temp, float(world), True)
        sl = slice(r * n_local, (r + 1) * n_local)
        _close(loss.item(), want.item(), 1e-5, "loss")
        _close(da.cpu().numpy(), world * ar.grad[sl].numpy(), 1e-4, "da")
        _close(db.cpu().numpy(), world * br.grad[sl].numpy(), 1e-4, "db")
