"""Seeded synthetic inputs for the five BASELINE.json configurations.

Everything is generated on the CPU with an explicit `torch.Generator` so that the oracle, the CUDA path,
the golden-fixture script and `bench.py` see bit-identical inputs.  Shapes follow the reference's configs:
res4 stride 16 / 1024 channels / 14x14 pooler (`config/defaults.py:369,423-426`), 512 RoIs per image with
25 % foreground (`:379-381`), RPN pre-NMS 12000 boxes at IoU 0.7 (`:344-355`), temperature 0.01,
focal gamma 0.5, background weight 0.2 (`configs/VOC-Experiments/faster_rcnn_CLIP_R_50_C4.yaml`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch


@dataclass(frozen=True)
class PathConfig:
    name: str
    n_images: int          # images per GPU
    img_h: int
    img_w: int
    rois_per_image: int
    num_classes: int       # K (foreground concepts); the head has K+1 columns
    channels: int = 1024
    pooled: int = 14
    sampling_ratio: int = 0
    stride: int = 16
    emb_dim: int = 1024
    align_dim: int = 256
    regions_per_image: int = 16   # rcnn.py:437
    temperature: float = 0.01
    focal_gamma: float = 0.5
    bg_weight: float = 0.2
    seed: int = 0

    @property
    def feat_hw(self) -> Tuple[int, int]:
        return (math.ceil(self.img_h / self.stride), math.ceil(self.img_w / self.stride))

    @property
    def n_rois(self) -> int:
        return self.n_images * self.rois_per_image


CONFIGS = {
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    "cpu_ref": PathConfig("cpu_ref", 2, 600, 1000, 512, 20, seed=0),
    # configs[1]: VOC->Clipart shape on 1 B200 (the configuration the metric is quoted on)
    "voc": PathConfig("voc", 16, 600, 1000, 512, 20, seed=1),
    # configs[2]: Cityscapes->Foggy/BDD shape, per-GPU share decided by the caller
    "city": PathConfig("city", 16, 1024, 2048, 512, 8, seed=2),
    # small shapes for parity tests
    "tiny": PathConfig("tiny", 2, 160, 256, 24, 5, channels=64, emb_dim=128, align_dim=64, seed=7),
}


def make_boxes(n: int, img_h: int, img_w: int, gen: torch.Generator, min_side: float = 16.0,
               degenerate_frac: float = 0.01) -> torch.Tensor:
    """`n` xyxy boxes: centre ~U(image), sides log-uniform in [min_side, image side], clipped to the image
    (mirrors `Boxes.clip`, structures/boxes.py:192-206); `degenerate_frac` of them are made zero-area or
    border-touching for edge coverage."""
    cx = torch.rand(n, generator=gen) * img_w
    cy = torch.rand(n, generator=gen) * img_h
    w = torch.exp(torch.rand(n, generator=gen) * (math.log(img_w) - math.log(min_side)) + math.log(min_side))
    h = torch.exp(torch.rand(n, generator=gen) * (math.log(img_h) - math.log(min_side)) + math.log(min_side))
    boxes = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], dim=1)
    boxes[:, 0::2].clamp_(0, img_w)
    boxes[:, 1::2].clamp_(0, img_h)
    n_deg = int(round(n * degenerate_frac))
    if n_deg > 0:
        idx = torch.randperm(n, generator=gen)[:n_deg]
        kind = torch.arange(n_deg) % 3
        for k, i in zip(kind.tolist(), idx.tolist()):
            if k == 0:      # zero height
                boxes[i, 3] = boxes[i, 1]
            elif k == 1:    # touches the bottom-right border
                boxes[i, 2], boxes[i, 3] = float(img_w), float(img_h)
            else:           # zero width at the left border
                boxes[i, 0] = boxes[i, 2] = 0.0
    return boxes.float()


def make_rois(cfg: PathConfig, gen: torch.Generator, n_images: Optional[int] = None,
              rois_per_image: Optional[int] = None) -> torch.Tensor:
    """[R,5] pooler-format RoIs (batch idx, x0,y0,x1,y1), grouped by image like
    `convert_boxes_to_pooler_format` produces (poolers.py:61-95)."""
    n_images = cfg.n_images if n_images is None else n_images
    rpi = cfg.rois_per_image if rois_per_image is None else rois_per_image
    parts = []
    for i in range(n_images):
        b = make_boxes(rpi, cfg.img_h, cfg.img_w, gen)
        parts.append(torch.cat([torch.full((rpi, 1), float(i)), b], dim=1))
    return torch.cat(parts, 0) if parts else torch.zeros(0, 5)


def make_features(cfg: PathConfig, gen: torch.Generator, n_images: Optional[int] = None) -> torch.Tensor:
    n_images = cfg.n_images if n_images is None else n_images
    hf, wf = cfg.feat_hw
    return torch.randn(n_images, cfg.channels, hf, wf, generator=gen)


def make_head_inputs(cfg: PathConfig, gen: torch.Generator, n_rois: Optional[int] = None):
    """x [R,D] ~N(0,1); concept embeddings W [K,D] ~N(0,1); frozen zero background embedding
    (fast_rcnn.py:458-461); gt ~ 75 % background (= K), rest uniform foreground."""
    r = cfg.n_rois if n_rois is None else n_rois
    x = torch.randn(r, cfg.emb_dim, generator=gen)
    w = torch.randn(cfg.num_classes, cfg.emb_dim, generator=gen)
    w_bg = torch.zeros(1, cfg.emb_dim)
    is_bg = torch.rand(r, generator=gen) < 0.75
    fg = torch.randint(0, cfg.num_classes, (r,), generator=gen)
    gt = torch.where(is_bg, torch.full_like(fg, cfg.num_classes), fg)
    return x, w, w_bg, gt


def make_align_inputs(cfg: PathConfig, gen: torch.Generator, n_images: Optional[int] = None):
    """(image-level src, tgt [B,256]; region-level src, tgt [16*B,256]) ~N(0,1) — the projector outputs
    feeding rcnn.py:305-317 and :455-468."""
    b = cfg.n_images if n_images is None else n_images
    d = cfg.align_dim
    img_src, img_tgt = torch.randn(b, d, generator=gen), torch.randn(b, d, generator=gen)
    reg_src = torch.randn(b * cfg.regions_per_image, d, generator=gen)
    reg_tgt = torch.randn(b * cfg.regions_per_image, d, generator=gen)
    return img_src, img_tgt, reg_src, reg_tgt


def make_nms_inputs(m: int, img_h: int, img_w: int, gen: torch.Generator, num_classes: int = 1,
                    tie_frac: float = 0.01):
    """NMS candidates: boxes from `make_boxes` (no degenerate-area filtering: zero-area pairs give
    NaN IoU and must be kept), scores ~U(0,1) with `tie_frac` exact ties, class ids uniform."""
    boxes = make_boxes(m, img_h, img_w, gen)
    scores = torch.rand(m, generator=gen)
    n_tie = int(round(m * tie_frac))
    if n_tie > 1:
        idx = torch.randperm(m, generator=gen)[:n_tie]
        scores[idx[1::2]] = scores[idx[: len(idx[1::2]) * 2: 2]]
    idxs = torch.randint(0, num_classes, (m,), generator=gen) if num_classes > 1 else torch.zeros(m, dtype=torch.int64)
    return boxes, scores, idxs


def generator(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(seed)
    return g
