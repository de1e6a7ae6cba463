#!/usr/bin/env python
"""bench.py — RoIs/s of CDDMSL's region-level hot path (ROIAlign fwd+bwd, CLIP region-text head fwd+bwd,
caption-consistency alignment loss fwd+bwd) on N B200s; BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload voc|city]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic input (per GPU: 16 images, 512 RoIs per
image).  `value` is timed with the inputs resident in HBM through the C-ABI-backed ops; `e2e` runs the same
work through the reference-shaped public API (ROIAlign / FastRCNNOutputLayers / caption_consistency_loss +
autograd) with pinned-host inputs copied in and results copied out every step.  `--impl reference` times the
reference's CPU path (oracle/torch_ref.py: torchvision CPU ops + ATen, all host threads) on a bounded sample.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cddmsl_b200 import synth  # noqa: E402

METRIC = "RoIs/sec (ROIAlign+CLIP region-text head fwd+bwd)"
UNIT = "RoIs/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="voc", choices=["voc", "city", "tiny"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--e2e-d2h", choices=["losses", "grads"], default="losses",
                   help="what the e2e arm reads back every step: the step's result (the three losses, 16 bytes -- in "
                        "training the gradients stay on the device and feed the backbone's backward), or additionally "
                        "every gradient the path produces (191 MB at configs[1]; doubles the PCIe traffic)")
    p.add_argument("--tune", action="append", default=[], help="key=value for cddmsl_tune (kernel sweeps)")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled every 100 ms DURING the timed region, through NVML in-process
    (nvidia_ml_py).  Spawning `nvidia-smi -lms` instead costs an NVML initialisation inside the timed region, which
    showed up as single 20+ ms stalls of whatever kernel was running; `nvidia-smi` is only the fallback."""
    NAMES = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
             "hw_power_brake_slowdown": 0x80}

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread, self.h = index, [], False, None, None
        self.recording = False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML enumerates physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                if self.recording:
                    self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.05)

    def warm(self):
        """Start polling BEFORE the warm-up steps: the first NVML queries of a process take the driver lock for tens
        of milliseconds (seen as one 20-100 ms stall of the kernel launches); only samples taken between start()
        and stop() are kept."""
        self.recording = False
        if self.nvml is not None and self.thread is None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def start(self):
        self.warm()
        self.recording = True

    def stop(self):
        if self.nvml is None:
            return self._nvidia_smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return self._nvidia_smi_once()
        sm = [s[0] for s in self.samples]
        reasons = set()
        for _, _, r in self.samples:
            for nm, bit in self.NAMES.items():
                if r & bit:
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.max_sm,
                "power_w_max": max(s[1] for s in self.samples), "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvml"}

    def _nvidia_smi_once(self):
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "power_w_max": float(f[2]), "samples": 1,
                    "reasons": [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")],
                    "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}


# -------------------------------------------------------------------------------------- reference arm
def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers are allocated, so
    the pages the copy engines read and write live on the GPU's own NUMA node (with N ranks streaming ~190 MB each
    way per step, remote-node pinned memory caps the aggregate PCIe rate).  Best effort: returns the number of CPUs
    bound, 0 when NVML / affinity is unavailable or the container's cpuset has no CPU on that node."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except Exception:
                phys = index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 64, 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        use = cpus & os.sched_getaffinity(0)
        if use:
            os.sched_setaffinity(0, use)
        return len(use)
    except Exception:
        return 0


def cpu_reference_step(cfg, sample_images: int, sample_rpi: int, seed: int):
    """One pass of the reference's CPU path on a bounded sample; returns (seconds, n_rois)."""
    from oracle import torch_ref

    g = synth.generator(seed)
    feat = synth.make_features(cfg, g, n_images=sample_images).requires_grad_(True)
    rois = synth.make_rois(cfg, g, n_images=sample_images, rois_per_image=sample_rpi)
    r = rois.shape[0]
    x, w, w_bg, gt = synth.make_head_inputs(cfg, g, n_rois=r)
    x.requires_grad_(True)
    i_s, i_t, r_s, r_t = [t.requires_grad_(True) for t in synth.make_align_inputs(cfg, g, n_images=sample_images)]
    t0 = time.perf_counter()
    out = torch_ref.roi_align(feat, rois, (cfg.pooled, cfg.pooled), 1.0 / cfg.stride, cfg.sampling_ratio, True)
    scores = torch_ref.clip_head_scores(x, w, w_bg, cfg.temperature)
    loss_cls = torch_ref.focal_loss(scores, gt, cfg.num_classes, cfg.focal_gamma, cfg.bg_weight)
    l_img = torch_ref.caption_consistency_loss(i_t, i_s)
    l_reg = torch_ref.caption_consistency_loss(r_s, r_t)
    torch.autograd.backward([out, loss_cls, l_img, l_reg],
                            [out.detach(), torch.ones(()), torch.ones(()), torch.ones(())])
    dt = time.perf_counter() - t0
    return dt, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = synth.CONFIGS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img, rpi = 2, 64
    for i in range(args.warmup):
        cpu_reference_step(cfg, n_img, rpi, 10 + i)
    t, rois = 0.0, 0
    for i in range(args.steps):
        dt, r = cpu_reference_step(cfg, n_img, rpi, 100 + i)
        t += dt
        rois += r
    v = rois / t
    sample = (f"{n_img} images x {rpi} RoIs of the '{cfg.name}' workload per step ({cfg.feat_hw[0]}x{cfg.feat_hw[1]} map, "
              f"{cfg.channels} ch, 14x14, K={cfg.num_classes}) + alignment losses; torchvision CPU ops + ATen "
              f"(oracle/torch_ref.py), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, n_gpus):
    return {"workload": f"{cfg.name}: per GPU {cfg.n_images} images {cfg.img_h}x{cfg.img_w} -> res4 "
                        f"{cfg.channels}x{cfg.feat_hw[0]}x{cfg.feat_hw[1]}, {cfg.rois_per_image} RoIs/img, "
                        f"ROIAlign 14x14 s=0 aligned, {cfg.num_classes}-concept CLIP head + bg (T=0.01, focal 0.5, "
                        f"bg 0.2), caption-consistency loss image-level n={cfg.n_images}/GPU and region-level "
                        f"n={cfg.n_images * cfg.regions_per_image}/GPU (256-d), fwd+bwd",
            "rois_per_gpu": cfg.n_rois, "global_rois": cfg.n_rois * n_gpus, "parallelism": f"dp{n_gpus} (by image)",
            "l2_policy": "inputs larger than L2 (6.6 GB pooled tensor, 157 MB feature map per step)"}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    from cddmsl_b200 import _lib, ops
    from cddmsl_b200.layers import ROIAlign
    from cddmsl_b200.modeling import (Box2BoxTransform, FastRCNNOutputLayers, caption_consistency_loss,
                                      image_caption_consistency_loss)
    from cddmsl_b200.structures import Boxes, Instances

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    numa_cpus = bind_to_gpu_numa_node(local) if (world > 1 and not os.environ.get("CDDMSL_NO_NUMA_BIND")) else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    for kv in args.tune:
        k, v = kv.split("=")
        assert _lib.tune(k, int(v)), f"unknown tuning key {k}"

    cfg = synth.CONFIGS[args.workload]
    g = synth.generator(cfg.seed + 1000 * rank)
    P, scale = cfg.pooled, 1.0 / cfg.stride
    h_feat = synth.make_features(cfg, g).pin_memory()
    h_rois = synth.make_rois(cfg, g).pin_memory()
    hx, hw, hwbg, hgt = synth.make_head_inputs(cfg, g)
    hx, hgt = hx.pin_memory(), hgt.pin_memory()
    h_align = [t.pin_memory() for t in synth.make_align_inputs(cfg, g)]
    R, N, C = cfg.n_rois, cfg.n_images, cfg.channels
    Hf, Wf = cfg.feat_hw

    feat, rois, x, gt = h_feat.to(dev), h_rois.to(dev), hx.to(dev), hgt.to(dev)
    w, w_bg = hw.to(dev), hwbg.to(dev)
    a_is, a_it, a_rs, a_rt = [t.to(dev) for t in h_align]
    one = torch.ones(1, device=dev)

    def align_fused(a, b):
        packed, norms = ops.align_pack(a, b)
        if world > 1:
            allp = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=dev)
            dist.all_gather_into_tensor(allp, packed)
        else:
            allp = packed.unsqueeze(0)
        return ops.align_loss(allp, norms, rank, one, True)

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]

    def device_step(i=None):
        e = ev[i] if i is not None else None
        if e: e[0].record()
        out = ops.roi_align(feat, rois, scale, P, P, cfg.sampling_ratio, True)
        if e: e[1].record()
        loss, _, dx, stats = ops.clip_head_loss(x, w, w_bg, gt, cfg.temperature, ops.LOSS_FOCAL, cfg.focal_gamma,
                                                cfg.bg_weight, one, False, False, True)
        if e: e[2].record()
        l_img = align_fused(a_it, a_is)
        l_reg = align_fused(a_rs, a_rt)
        if e: e[3].record()
        gin = ops.roi_align_backward(out, rois, scale, P, P, N, C, Hf, Wf, cfg.sampling_ratio, True)
        if e: e[4].record()
        return loss, l_img[0], l_reg[0], gin, dx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.warm()
    res = None
    for _ in range(max(args.warmup, 3)):
        # keep the previous step's results alive exactly like the timed loop does: otherwise the caching allocator
        # meets its steady-state footprint (two live gradient maps) only at timed step 1 and stalls the host in
        # cudaMalloc for 10-100 ms (measured)
        res = device_step()
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(args.steps):
        res = device_step(i)
    t1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = t0.elapsed_time(t1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    seg_all = [[ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(args.steps)] for j in range(4)]
    seg = [statistics.mean(v) for v in seg_all]
    seg_minmax = [(min(v), statistics.median(v), max(v)) for v in seg_all]
    mstats = torch.cuda.memory_stats(dev)
    losses = [float(v) for v in res[:3]]

    # ---------------- end to end through the reference-shaped API, host buffers in and out every step
    e2e = None
    if not args.no_e2e:
        pooler = ROIAlign((P, P), scale, cfg.sampling_ratio, aligned=True)
        head = FastRCNNOutputLayers(cfg.emb_dim, box2box_transform=Box2BoxTransform((10.0, 10.0, 5.0, 5.0)),
                                    num_classes=cfg.num_classes, clip_cls_emb=(True, hw, "CLIPRes5ROIHeads", cfg.emb_dim),
                                    bg_cls_loss_weight=cfg.bg_weight,
                                    openset_test=(None, None, cfg.temperature, cfg.focal_gamma)).to(dev).train()
        head.bbox_pred.requires_grad_(False)
        o_gin = torch.empty_like(h_feat).pin_memory()
        o_dx = torch.empty_like(hx).pin_memory()
        o_ga = [torch.empty_like(t).pin_memory() for t in h_align]
        o_sc = torch.empty(4).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in [h_feat, h_rois, hx, hgt] + h_align)
        outs_host = [o_gin, o_dx, o_sc] + o_ga if args.e2e_d2h == "grads" else [o_sc]
        d2h = sum(t.numel() * t.element_size() for t in outs_host)

        # Three streams, double-buffered device inputs: H2D of step i+1 and D2H of step i-1 overlap the compute of
        # step i (what a real input pipeline does).  Every byte still crosses PCIe inside the timed region.
        s_h2d, s_d2h, s_cmp = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
        host_in = [h_feat, h_rois, hx, hgt] + h_align
        dev_in = [[torch.empty_like(t, device=dev) for t in host_in] for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        box_pad = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)  # box-reg branch wants w,h > 0

        def issue_h2d(i):
            b = i & 1
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_cmp[b])          # the compute that last read this buffer set is done
                for d, h in zip(dev_in[b], host_in):
                    d.copy_(h, non_blocking=True)
                ev_in[b].record(s_h2d)

        ring = [None, None]   # device results whose D2H may still be in flight (kept alive instead of record_stream:
                              # a run-ahead host would otherwise make the caching allocator cudaMalloc a fresh
                              # 157 MB block every step -- measured 10 ms each)

        def e2e_step(i):
            b = i & 1
            if ring[b] is not None:
                s_cmp.wait_event(ev_out[b])          # D2H of step i-2 has read them: safe to hand back to the allocator
                ring[b] = None
            s_cmp.wait_event(ev_in[b])
            f = dev_in[b][0].detach().requires_grad_(True)
            r, gg = dev_in[b][1], dev_in[b][3]
            xx = dev_in[b][2].detach().requires_grad_(True)
            al = [t.detach().requires_grad_(True) for t in dev_in[b][4:]]
            out = pooler(f, r)
            inst = Instances((cfg.img_h, cfg.img_w))
            inst.proposal_boxes = Boxes(r[:, 1:] + box_pad)
            inst.gt_classes = gg
            scores, deltas = head(xx)
            lc = head.losses((scores, deltas.detach()), [inst])["loss_cls"]
            li = image_caption_consistency_loss(al[1], al[0])
            lr = caption_consistency_loss(al[2], al[3])
            torch.autograd.backward([out, lc, li, lr], [out.detach(), one[0], one[0], one[0]])
            sc = torch.stack([lc.detach(), li.detach(), lr.detach(), lc.detach() * 0])
            ev_cmp[b].record(s_cmp)
            grads = [f.grad, xx.grad, sc] + [t.grad for t in al] if args.e2e_d2h == "grads" else [sc]
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_cmp[b])
                for o, gsrc in zip(outs_host, grads):
                    o.copy_(gsrc, non_blocking=True)
                ev_out[b].record(s_d2h)
            ring[b] = grads

        def run_e2e(n):
            issue_h2d(0)
            for i in range(n):
                if i + 1 < n:
                    issue_h2d(i + 1)
                e2e_step(i)
            s_cmp.wait_event(ev_out[(n - 1) & 1])
            if n > 1:
                s_cmp.wait_event(ev_out[(n - 2) & 1])
            ring[0] = ring[1] = None

        # warm-up long enough for the caching allocator to meet the steady-state footprint of the run-ahead host
        # (a cudaMalloc inside the timed region costs tens of ms)
        run_e2e(6)
        barrier()
        if os.environ.get("CDDMSL_E2E_PROFILE"):   # diagnostic only: what runs on the device during an e2e step
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                run_e2e(4)
                barrier()
            with open(os.environ["CDDMSL_E2E_PROFILE"], "w") as fh:
                fh.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
            prof.export_chrome_trace(os.environ["CDDMSL_E2E_PROFILE"] + ".trace.json")
        k2 = max(4, args.steps)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        run_e2e(k2)
        s1.record()
        barrier()
        ms2 = s0.elapsed_time(s1)
        if world > 1:
            tt = torch.tensor([ms2], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms2 = float(tt.item())
        e2e = {"value": world * R * k2 / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms2 / k2, "steps": k2,
               "api": "ROIAlign.forward + FastRCNNOutputLayers.forward/losses + caption_consistency_loss + autograd",
               "d2h": "the three losses (gradients stay in HBM, as in training)" if args.e2e_d2h == "losses"
                      else "losses + every gradient of the path",
               "pipelining": "H2D(i+1) and D2H(i-1) on side streams overlap compute(i); all copies inside the timed region",
               "numa_bound_cpus": numa_cpus}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (ROIAlign fwd or bwd), algorithmic bytes / CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bytes_roi = R * (4 * C * P * P + 20) + N * C * Hf * Wf * 4
    names = ["roi_align_fwd", "clip_head_fwd_bwd", "align_loss_x2_fwd_bwd", "roi_align_bwd"]
    kern = {n: {"ms": round(seg[j], 4), "ms_min_med_max": [round(x, 4) for x in seg_minmax[j]]}
            for j, n in enumerate(names)}
    kern["allocator"] = {"num_device_alloc": int(mstats.get("num_device_alloc", 0)),
                         "num_alloc_retries": int(mstats.get("num_alloc_retries", 0)),
                         "reserved_GB": round(mstats.get("reserved_bytes.all.peak", 0) / 1e9, 2)}
    kern["roi_align_fwd"].update(gbs=bytes_roi / (seg[0] * 1e-3) / 1e9, bytes=bytes_roi)
    kern["roi_align_bwd"].update(gbs=bytes_roi / (seg[3] * 1e-3) / 1e9, bytes=bytes_roi)
    head_bytes = R * (2 * cfg.emb_dim * 4 + 8)
    kern["clip_head_fwd_bwd"].update(gbs=head_bytes / (seg[1] * 1e-3) / 1e9, bytes=head_bytes)
    dom = "roi_align_bwd" if seg[3] >= seg[0] else "roi_align_fwd"
    ach = kern[dom]["gbs"]
    step_bytes = 2 * bytes_roi + head_bytes
    traffic = None
    try:  # dram__bytes_read+write per launch of the dominant kernel, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))["kernels"]
        key = "roi_align_bwd_cl" if dom == "roi_align_bwd" else "roi_align_fwd_cl"
        traffic = next(v["dram_bytes"] for k, v in tj.items() if k.startswith(key)) if args.workload == "voc" else None
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src,
                "step_frac": (step_bytes / (ms / args.steps * 1e-3) / 1e9) / peak,
                "step_algorithmic_bytes": step_bytes}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref = synth.CONFIGS["cpu_ref"]
        cpu_reference_step(ref, 1, 32, 1)  # warm-up
        dt, r = cpu_reference_step(ref, ref.n_images, ref.rois_per_image, ref.seed)
        cpu_baseline = {"value": r / dt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                        "sample": "BASELINE.json configs[0]: 2 images x 512 RoIs (38x63 map, 1024 ch, 14x14, K=20) "
                                  "fwd+bwd once, torchvision CPU ops + ATen via oracle/torch_ref.py, all host threads"}

    line = {"metric": METRIC, "value": world * R * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kern,
            "losses": {"loss_cls": losses[0], "align_image": losses[1], "align_region": losses[2]}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
