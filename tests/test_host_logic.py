"""CPU: host-side logic of the mirror (no kernels): pooler box format, sharding, GatherLayer under a real
world_size-2 gloo group, box transform, containers, the PyTorch-arithmetic loss helpers."""
import math
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from cddmsl_b200 import synth
from oracle import torch_ref


def test_pooler_format_matches_reference_restatement():
    from cddmsl_b200.modeling import convert_boxes_to_pooler_format
    from cddmsl_b200.structures import Boxes

    g = synth.generator(0)
    lists = [synth.make_boxes(n, 600, 1000, g) for n in (5, 0, 7)]
    got = convert_boxes_to_pooler_format([Boxes(b) for b in lists])
    want = torch_ref.convert_boxes_to_pooler_format(lists)
    assert torch.equal(got, want) and got.shape == (12, 5)
    assert torch.equal(convert_boxes_to_pooler_format(lists), want)   # raw tensors accepted too


def test_shard_bounds_partition_images_evenly():
    from cddmsl_b200.modeling.caption_consistency import shard_bounds

    for world in (1, 2, 4, 8):
        spans = [shard_bounds(16, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 16
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert len({b - a for a, b in spans}) == 1
    with pytest.raises(AssertionError):
        shard_bounds(10, 4, 0)


def _gather_worker(rank, world, port, golden, q):
    import torch.distributed as dist

    from cddmsl_b200.modeling import GatherLayer
    from cddmsl_b200.modeling.caption_consistency import shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = np.load(os.path.join(golden, "align_world2.npz"))
        a = torch.from_numpy(w[f"a{rank}"]).requires_grad_(True)
        b = torch.from_numpy(w[f"b{rank}"]).requires_grad_(True)
        a_all = torch.cat(GatherLayer.apply(a), dim=0)
        b_all = torch.cat(GatherLayer.apply(b), dim=0)
        lo, hi = shard_bounds(a_all.shape[0], world, rank)
        assert torch.equal(a_all[lo:hi], a.detach())            # rank-major layout the fused kernel assumes
        loss = torch_ref.caption_consistency_loss(a_all, b_all)  # the cited reference arithmetic
        loss.backward()
        ok = (np.allclose(loss.item(), w["loss"], rtol=1e-6)
              and np.allclose(a.grad.numpy(), w[f"da{rank}"], rtol=1e-5, atol=1e-8)
              and np.allclose(b.grad.numpy(), w[f"db{rank}"], rtol=1e-5, atol=1e-8))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gatherlayer_world2_gloo(golden_dir):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, golden_dir, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_box2box_round_trip_and_containers():
    from cddmsl_b200.modeling import Box2BoxTransform
    from cddmsl_b200.structures import Boxes, Instances

    g = synth.generator(2)
    src = synth.make_boxes(50, 600, 1000, g, degenerate_frac=0.0)
    dst = synth.make_boxes(50, 600, 1000, g, degenerate_frac=0.0)
    t = Box2BoxTransform((10.0, 10.0, 5.0, 5.0))
    assert torch.allclose(t.apply_deltas(t.get_deltas(src, dst), src), dst, rtol=1e-4, atol=1e-2)
    b = Boxes(torch.tensor([[-5.0, 2.0, 30.0, 700.0], [4.0, 4.0, 4.0, 9.0]]))
    b.clip((600, 1000))
    assert torch.equal(b.tensor, torch.tensor([[0.0, 2.0, 30.0, 600.0], [4.0, 4.0, 4.0, 9.0]]))
    assert b.nonempty().tolist() == [True, False]
    inst = Instances((600, 1000), proposal_boxes=b, scores=torch.tensor([0.1, 0.9]))
    assert len(inst) == 2 and len(inst[torch.tensor([1])]) == 1 and inst.has("scores")
    with pytest.raises(AssertionError):
        inst.set("bad", torch.zeros(3))


def test_predictor_pytorch_arithmetic_helpers_match_oracle():
    """The non-fused branch of `losses` (scores not produced by forward) is plain PyTorch and runs on CPU."""
    from cddmsl_b200.modeling import Box2BoxTransform, FastRCNNOutputLayers

    cfg = synth.CONFIGS["tiny"]
    g = synth.generator(5)
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=64)
    m = FastRCNNOutputLayers(cfg.emb_dim, box2box_transform=Box2BoxTransform((10.0, 10.0, 5.0, 5.0)),
                             num_classes=cfg.num_classes, clip_cls_emb=(True, w, "CLIPRes5ROIHeads", cfg.emb_dim),
                             bg_cls_loss_weight=0.2, openset_test=(None, None, 0.01, 0.5))
    assert not m.cls_score.weight.requires_grad and not m.cls_bg_score.weight.requires_grad
    assert float(m.cls_bg_score.weight.abs().sum()) == 0.0 and m.temperature == 0.01
    assert set(dict(m.named_parameters())) == {"cls_score.weight", "cls_bg_score.weight", "bbox_pred.weight",
                                               "bbox_pred.bias"}
    scores = torch_ref.clip_head_scores(x, w, w_bg, 0.01)
    got = m.focal_loss(scores, gt, gamma=0.5)
    want = torch_ref.focal_loss(scores, gt, cfg.num_classes, 0.5, 0.2)
    assert torch.allclose(got, want, rtol=1e-6)
    assert m.focal_loss(scores[:0], gt[:0]).item() == 0.0


def test_synthetic_generators_are_deterministic():
    cfg = synth.CONFIGS["voc"]
    r1 = synth.make_rois(cfg, synth.generator(cfg.seed), n_images=2)
    r2 = synth.make_rois(cfg, synth.generator(cfg.seed), n_images=2)
    assert torch.equal(r1, r2) and r1.shape == (1024, 5)
    assert (r1[:, 3] >= r1[:, 1]).all() and (r1[:, 4] >= r1[:, 2]).all()
    assert cfg.feat_hw == (38, 63) and synth.CONFIGS["city"].feat_hw == (64, 128)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU, oracle port) must print ONE JSON line with the agreed keys; the
    non-zero ranks of a torchrun launch print nothing."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "RoIs/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_box_reg_loss_masked_equals_reference_selection():
    """fast_rcnn.py:646-689: the sync-free masked evaluation gives the loss and gradient of the nonzero()-selected one,
    also with degenerate background proposals (which the reference never feeds to get_deltas)."""
    import cddmsl_b200.modeling.fast_rcnn as fr
    from cddmsl_b200.modeling import Box2BoxTransform

    g = torch.Generator().manual_seed(5)
    R, K = 200, 7
    boxes = synth.make_boxes(R, 600, 1000, g, degenerate_frac=0.0)
    gt_boxes = synth.make_boxes(R, 600, 1000, g, degenerate_frac=0.0)
    gt = torch.randint(0, K + 1, (R,), generator=g)
    bg_rows = (gt == K).nonzero()[:5, 0]
    boxes[bg_rows] = torch.tensor([10.0, 10.0, 10.0, 10.0])       # zero-area background rows
    for agnostic in (True, False):
        m = object.__new__(fr.FastRCNNOutputLayers)
        m.num_classes, m.box_reg_loss_type, m.smooth_l1_beta = K, "smooth_l1", 0.5
        m.box2box_transform = Box2BoxTransform((10.0, 10.0, 5.0, 5.0))
        deltas = torch.randn(R, 4 if agnostic else 4 * K, generator=g)
        res = []
        for strict in (True, False):
            fr.STRICT_BOX_REG_SYNC = strict
            try:
                d = deltas.clone().requires_grad_(True)
                l = m.box_reg_loss(boxes, gt_boxes, d, gt)
                l.backward()
                res.append((l.item(), d.grad.clone()))
            finally:
                fr.STRICT_BOX_REG_SYNC = False
        assert math.isfinite(res[1][0])
        assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[0][0])
        assert torch.allclose(res[0][1], res[1][1], rtol=1e-6, atol=1e-8)
