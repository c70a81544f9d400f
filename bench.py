#!/usr/bin/env python
"""bench.py -- RGBA-VAE encode + sample + decode throughput (MPix/s), BASELINE.json's metric.

Workload (N=1): config c2 -- RGBA-VAE reconstruction, bf16, 1024x1024, batch 8 on one B200 with the
alpha-over-white PSNR validation, Qwen-Image VAE architecture, random-init weights, synthetic RGBA.
One "step" = one pass of the hot path over one batch: [0,1] RGBA -> encode -> posterior sample with
supplied noise -> decode -> clamp -> composite-over-white PSNR + alpha MAE.  Under torchrun (N>1) every
rank runs the same per-rank batch (weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arch qwen|flux]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the restatement of the
reference's diffusers path; diffusers itself is not installable offline) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rgba_vae_encode_decode_mpix_per_s"
UNIT = "MPix/s"
TFLOP_PER_IMAGE_1024 = {"qwen": 7.5657, "flux": 15.3596}  # SURVEY.md 8(d) / BASELINE.md section 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="qwen", choices=["qwen", "flux"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2: fixed 1024^2 batches (the headline); c3: mixed-aspect bucket-pure batches sharded across ranks; "
                         "c4: the rgba_vae training step (fwd + bwd + gradient all-reduce + AdamW), data parallel")
    ap.add_argument("--train-size", type=int, default=1024, help="c4: image side (SURVEY 8: 12 x 4 x 1024^2 into the encoder per GPU)")
    ap.add_argument("--train-batch", type=int, default=4, help="c4: per-GPU batch (configs/flux_vae.yaml: 4)")
    ap.add_argument("--batches", type=int, default=64, help="c3: bucket-pure batches in the whole job")
    return ap.parse_args()


def workload_config(a):
    return {"workload": f"c2: RGBA-VAE reconstruction bf16 {a.size}x{a.size} batch {a.batch} per GPU + "
                        "alpha-over-white PSNR validation (encode -> sample -> decode)",
            "arch": a.arch, "batch_per_gpu": a.batch, "height": a.size, "width": a.size,
            "weights": "random-init seed 0", "parallelism": f"batch-sharded x{a.gpus}, no collective"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_forward_factory(arch, size):
    import torch

    from oracle import vae_oracle as O

    vae = O.build_oracle(arch, seed=0)
    x = O.synthetic_rgba(1, size, size, seed=1, structured=True)
    noise = torch.randn(1, 16, size // 8, size // 8, generator=torch.Generator().manual_seed(2))

    def step():
        with torch.no_grad():
            recon, post, _ = O.rgba_vae_forward(vae, x, noise)
            return float(O.validation_metrics(recon, x, backgrounds=(1.0,))[1.0][0])

    return step


def cpu_pick_size(arch, want):
    """Largest sample (<= want) whose estimated step time stays under ~15 s on this host."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    probe = cpu_forward_factory(arch, 256)
    probe()
    t0 = time.perf_counter()
    probe()
    t256 = time.perf_counter() - t0
    size = 256
    while size * 2 <= want and t256 * ((size * 2) / 256) ** 2 * 1.3 < 10.0:
        size *= 2
    return size, t256


def run_cpu_baseline(arch, want_size, steps, warmup):
    import torch

    size, _ = cpu_pick_size(arch, want_size)
    step = cpu_forward_factory(arch, size)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    mpix = size * size * steps / 1e6 / total
    return {"value": mpix, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle fp32 (torch CPU), 1x4x{size}x{size} per step, {steps} steps after {warmup} warm-up, "
                      f"{total:.1f} s; {os.cpu_count()} logical cores"}, total / steps * 1e3


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = run_cpu_baseline(a.arch, a.size, a.steps, a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/rv_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))),
                "tflops_burst": float(p.get("bf16_tflops", 1590.0)), "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def gpu_arm(a):
    import torch
    import torch.distributed as dist

    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S = a.batch, a.size

    torch.manual_seed(0)
    vae = R.RgbaAutoencoder(a.arch)  # torch default init == the oracle's
    model = R.RgbaVAE(vae.to(dev, torch.bfloat16))

    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.rand(B, 4, S, S, generator=g).to(torch.bfloat16).pin_memory()
    n_host = torch.randn(B, 16, S // 8, S // 8, generator=torch.Generator().manual_seed(2 + rank)).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    x_dev, n_dev = x_host.to(dev), n_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    use_graph = not a.no_graph

    def step(x, noise):
        if use_graph:  # one captured CUDA graph per shape: the same kernels, replayed without host gaps
            return model.forward_graphed(x, noise, backgrounds=((1.0, 1.0, 1.0),))[2]
        recon, _ = model(x, noise=noise)
        return ops.composite_psnr(recon, x, [(1.0, 1.0, 1.0)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            flush.zero_()
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident():
        step(x_dev, n_dev)

    def end_to_end():
        xd = x_host.to(dev, non_blocking=True)
        nd = n_host.to(dev, non_blocking=True)
        out_host.copy_(step(xd, nd), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(a.warmup, 3)):
        resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(resident, a.steps)
    # kernels per step, counted on one eagerly launched step (graph replays bypass the library's launch counter)
    l0 = ops.launch_count()
    recon_e, _ = model(x_dev, noise=n_dev)
    ops.composite_psnr(recon_e, x_dev, [(1.0, 1.0, 1.0)])
    torch.cuda.synchronize()
    del recon_e
    launches = (ops.launch_count() - l0) * a.steps
    clocks = sampler.stop() if rank == 0 else None
    end_to_end()
    ms_e2e = timed(end_to_end, a.steps)
    psnr_white = float(out_host[:, 0].mean())

    # roofline of the dominant kernel (the tcgen05 implicit-GEMM conv): CUDA events on the launching
    # stream around every launch of one more step (rv_prof_*), algorithmic FLOPs / summed duration
    prof = None
    if rank == 0:
        torch.cuda.synchronize()
        ops.prof_begin()
        recon_p, _ = model(x_dev, noise=n_dev)   # eager launches: per-kernel CUDA events cannot be recorded in a replay
        ops.composite_psnr(recon_p, x_dev, [(1.0, 1.0, 1.0)])
        prof = ops.prof_end()
        del recon_p

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    mpix_step = world * B * S * S / 1e6
    value = mpix_step * a.steps / (ms_total / 1e3)
    e2e_value = mpix_step * a.steps / (ms_e2e / 1e3)
    conv = prof["conv_tc"]
    ach = conv["work"] / 1e12 / (conv["ms"] / 1e3) if conv["ms"] > 0 else 0.0
    roof = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel (CTA pairs) / conv_halo_kernel: every tcgen05 implicit-GEMM conv and GEMM launch of one step",
            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": None,
            "peak_source": f"{pk['source']} bf16_tflops_sustained", "launches_per_step": conv["launches"],
            "ms_per_step": conv["ms"], "algorithmic_tflop_per_step": conv["work"] / 1e12}
    kernels = {}
    for name, rec in prof.items():
        if rec["launches"] == 0:
            continue
        entry = {"ms": round(rec["ms"], 4), "launches": rec["launches"]}
        if name in ("conv_tc", "conv_direct", "attention"):
            entry["tflops"] = rec["work"] / 1e12 / (rec["ms"] / 1e3) if rec["ms"] > 0 else None
        else:
            gbs = rec["work"] / 1e9 / (rec["ms"] / 1e3) if rec["ms"] > 0 else None
            entry["gbs"] = gbs
            entry["hbm_frac"] = gbs / pk["hbm_gbs"] if gbs else None
        kernels[name] = entry
    whole = world * B * TFLOP_PER_IMAGE_1024[a.arch] * (S * S / 1048576.0) * a.steps / (ms_total / 1e3) if S == 1024 else None

    cpu_b = None
    if world == 1 and not a.no_cpu_baseline:
        cpu_b, _ = run_cpu_baseline(a.arch, S, 4, 1)  # ~10-15 s of host work on a bounded sample

    cfg = workload_config(a)
    cfg["launch"] = "CUDA graph replay (one capture per shape)" if use_graph else "eager launches"
    cfg["l2"] = "256 MiB buffer rewritten between timed iterations (plus multi-GB activation working set per step)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": (x_host.numel() + n_host.numel()) * 2 * world,
                    "d2h_bytes_per_step": out_host.numel() * 4 * world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels,
            "whole_step_tflops": whole, "cpu_baseline": cpu_b, "psnr_white_db": psnr_white}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c3_arm(a):
    """Config c3: bucket-pure mixed-aspect batches (<= 1 MP, sides % 32 == 0) assigned to ranks
    longest-first; every rank validates its own batches with no data-path collective; the job time is the slowest
    rank's device time, per-sample PSNR vectors are gathered once after the run."""
    import torch
    import torch.distributed as dist

    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = R.RgbaVAE(R.RgbaAutoencoder(a.arch).to(dev, torch.bfloat16))
    shapes = sharding.sample_bucket_batches(a.batches, a.batch, seed=1234)
    mine = sharding.assign_batches(shapes, world)[rank]
    data = {}
    for shp in sorted(set(shapes[i] for i in mine)):  # one synthetic batch per bucket shape, reused
        b, h, w = shp
        g = torch.Generator().manual_seed(h * 10007 + w)
        data[shp] = (torch.rand(b, 4, h, w, generator=g).to(dev, torch.bfloat16),
                     torch.randn(b, 16, h // 8, w // 8, generator=g).to(dev, torch.bfloat16))

    def run_all():
        out = []
        for i in mine:
            x, n = data[shapes[i]]
            out.append(model.forward_graphed(x, n, backgrounds=((1.0, 1.0, 1.0),))[2][:, 0].clone())
        return torch.cat(out) if out else torch.zeros(0, device=dev)

    for _ in range(max(1, min(a.warmup, 2))):
        run_all()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        psnr = run_all()
    e1.record()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    allp = sharding.gather_per_sample(psnr)
    if rank == 0:
        mpix = sum(b * h * w for (b, h, w) in shapes) / 1e6
        line = {"metric": METRIC, "value": mpix * a.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"c3: {a.batches} bucket-pure mixed-aspect batches of {a.batch} (<= 1 MP, sides % 32), "
                                       "longest-first sharding, no collective", "arch": a.arch,
                           "bucket_shapes": len(set(shapes)), "mpix_per_step": mpix},
                "samples_validated": int(allp.numel()), "psnr_white_db": float(allp.mean())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c4_arm(a):
    """Config c4: the rgba_vae training step (src/training/rgba_vae_stage.py:433-523 without LPIPS): triplet, encode,
    sample, decode, AlphaVAE loss (loss_reduce_mean like configs/flux_vae.yaml) + 1e-6 KL, hand-written backward,
    bucketed NCCL gradient all-reduce, clip_grad_norm_(1.0) and AdamW(1e-5, betas (0.5, 0.9)).  Weak scaling: every
    rank trains on its own batch.  value = trained pixels per second over all ranks (inputs resident in HBM)."""
    import torch
    import torch.distributed as dist

    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops, sharding
    from ragb_vae_b200.trainer import VaeTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    vae = R.RgbaAutoencoder("qwen").to(dev, torch.bfloat16)
    step = VaeTrainStep(vae, lr=1e-5, kl_scale=1e-6, loss_module=R.AlphaVaeLoss(reduce_mean=True))
    B, S = a.train_batch, a.train_size
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.rand(B, 4, S, S, generator=g).pin_memory()
    n_host = torch.randn(B, 16, S // 8, S // 8, generator=g).pin_memory()
    x, noise = x_host.to(dev), n_host.to(dev)
    warm = max(a.warmup, 3)
    run = step.step if a.no_graph else step.step_graphed
    for _ in range(warm):
        m = run(x, noise)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ops.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        m = run(x, noise)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    launches = ops.launch_count() - l0
    if not a.no_graph:  # replayed launches are not seen by the host-side counter: one capture's count x steps
        launches = step.launches_per_replay * a.steps
    # per-category CUDA events need eager launches: one more (eager) step outside the timed region
    ops.prof_begin()
    step.step(x, noise)
    torch.cuda.synchronize()
    prof = ops.prof_end()
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    # end to end: host batch in, loss out
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        m = run(x_host.to(dev, non_blocking=True), n_host.to(dev, non_blocking=True))
        loss_host = float(m["train/loss"])
    torch.cuda.synchronize()
    ms_e2e = sharding.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    in_sync = None
    if world > 1:  # data-parallel replicas must hold identical weights after the same number of steps
        chk = torch.stack([step.opt.master.double().sum(), step.opt.master.double().abs().sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        in_sync = all(bool(torch.equal(c, allc[0])) for c in allc)
    if rank == 0:
        mpix = world * B * S * S / 1e6
        kernels = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in prof.items() if v["launches"]}
        pk = peaks()
        conv = prof["conv_tc"]
        ach = conv["work"] / 1e12 / (conv["ms"] / 1e3) if conv["ms"] > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "every tcgen05 launch of one step: forward convs / GEMMs, data-gradient convs (conv_tc*, conv_halo) "
                                             "and conv_wgrad_kernel",
                "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": None,
                "peak_source": f"{pk['source']} bf16_tflops_sustained", "launches_per_step": conv["launches"],
                "ms_per_step": conv["ms"], "algorithmic_tflop_per_step": conv["work"] / 1e12}
        line = {"metric": "rgba_vae_train_step_mpix_per_s", "value": mpix * a.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"c4: rgba_vae training step, batch {B} x {S}x{S} per GPU, recon (reduce_mean) + 1e-6 KL, "
                                       "no LPIPS, bucketed NCCL gradient all-reduce, clip 1.0, AdamW",
                           "arch": "qwen", "parallelism": f"data parallel x{world}",
                           "launch": "eager launches" if a.no_graph else "CUDA graph replay of the whole step (incl. all-reduce + AdamW)"},
                "e2e": {"value": mpix * a.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                        "h2d_bytes_per_step": (x_host.numel() + n_host.numel()) * 4 * world, "d2h_bytes_per_step": 4 * world},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels, "loss": loss_host,
                "replicas_in_sync": in_sync}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "ours" and a.workload == "c3":
        c3_arm(a)
        return
    if a.impl == "ours" and a.workload == "c4":
        c4_arm(a)
        return
    if a.impl == "reference":
        reference_arm(a)
        return
    import __graft_entry__ as G

    B = G._load_builder()
    if not os.path.exists(B.LIB):  # normally prebuilt in-tree (it travels with the snapshot)
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            B.build()
        else:
            while not os.path.exists(B.LIB):
                time.sleep(1.0)
            time.sleep(2.0)
    gpu_arm(a)


if __name__ == "__main__":
    main()
