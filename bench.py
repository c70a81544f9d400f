#!/usr/bin/env python
"""bench.py -- RGBA-VAE encode + sample + decode throughput (MPix/s), BASELINE.json's metric.

Headline workload (config c2): RGBA-VAE reconstruction, bf16, 1024x1024, batch 8 per B200 with the alpha-over-white PSNR
validation, Qwen-Image VAE architecture, random-init weights, synthetic RGBA.  One "step" = one pass of the hot path over
one batch: [0,1] RGBA -> encode -> posterior sample with supplied noise -> decode -> clamp -> composite-over-white PSNR +
alpha MAE.  Under torchrun (N>1) every rank runs the same per-rank batch (weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arch qwen|flux] [--no-extras]

Prints ONE JSON line (rank 0).  The headline keys (`value`, `e2e`, `roofline`, `cpu_baseline`, ...) are config c2; the other
north-star configs ride along under `extra` so that the driver's records hold them too:
  extra.c4  the rgba_vae training step, data parallel, with the NCCL gradient all-reduce (bytes, exposed time per bucket)
  extra.c3  mixed-aspect bucket-pure batches sharded across the ranks
  extra.c5  Flux RGBA decode at 2048x2048 (one GPU)
  extra.gpu_eager_baseline  the oracle's nn.Modules on the same B200 in bf16 = PyTorch eager / cuDNN (a baseline, not ours)
`--workload c3|c4|c5` runs one of them alone with the full line.  `--impl reference` times the CPU oracle (the restatement
of the reference's diffusers path; diffusers itself is not installable offline) on the host cores.
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
FLUX_DECODE_2048_TFLOP = 48.496
# tolerances the parity tests hold that are looser than north_star's (tests/test_gpu_parity.py TOL_RAW_DECODE)
KNOWN_DEVIATIONS = ["bf16 raw decoder output vae.decode(z).sample in [-1,1] is held to 3e-2 relative (measured 2.1e-2 qwen / "
                    "2.2e-2 flux on random-init weights; latents and [0,1] reconstructions meet 2e-2); DESIGN.md section 2"]


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
    ap.add_argument("--no-extras", action="store_true", help="headline c2 only (skip extra.c4 / c3 / c5 / gpu_eager_baseline)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2: fixed 1024^2 batches (the headline, with the other configs under `extra`); c3 / c4 / c5: that config alone")
    ap.add_argument("--train-size", type=int, default=1024, help="c4: image side (SURVEY 8: 12 x 4 x 1024^2 into the encoder per GPU)")
    ap.add_argument("--train-batch", type=int, default=4, help="c4: per-GPU batch (configs/flux_vae.yaml: 4)")
    ap.add_argument("--batches", type=int, default=64, help="c3: bucket-pure batches in the whole job")
    return ap.parse_args()


def workload_config(a):
    """The SAME dict for our arm and the reference arm (the reference arm steps a bounded sample of it: cpu_baseline.sample)."""
    return {"workload": f"c2: RGBA-VAE reconstruction bf16 {a.size}x{a.size} batch {a.batch} per GPU + "
                        "alpha-over-white PSNR validation (encode -> sample -> decode)",
            "arch": a.arch, "batch_per_gpu": a.batch, "height": a.size, "width": a.size,
            "weights": "random-init seed 0", "parallelism": f"batch-sharded x{a.gpus}, no collective",
            "launch": "eager launches" if a.no_graph else "CUDA graph replay (one capture per shape)",
            "l2": "256 MiB buffer rewritten between timed iterations (plus a multi-GB activation working set per step)"}


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


def run_cpu_baseline(arch, want_size, steps, warmup, batch=8):
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
            "sample": f"oracle fp32 (torch CPU), ONE image 1x4x{size}x{size} per step (a bounded sample of the batch-{batch} "
                      f"workload: MPix/s does not depend on the batch on a CPU), {steps} steps after {warmup} warm-up, "
                      f"{total:.1f} s; {os.cpu_count()} logical cores"}, total / steps * 1e3


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = run_cpu_baseline(a.arch, a.size, a.steps, a.warmup, a.batch)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "one CPU process on rank 0's host cores whatever --gpus says: only the N=1 ratio against it is an anchor"}
    emit(line)


# ----------------------------------------------------------------------------------------------
# helpers of the GPU arms
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/rv_clocks_{os.getpid()}_{time.monotonic_ns()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

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


def measured_traffic():
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/r02_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


class Ctx:
    """torch.distributed plumbing shared by the arms: one process per GPU, NCCL for the barrier and the max-over-ranks time."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self._flush = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)  # > 126 MB L2
        self._flush.zero_()

    def timed(self, fn, steps, flush=True) -> float:
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks (ms)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            if flush:
                self.flush_l2()
            fn()
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1))

    def release(self):
        import gc

        gc.collect()
        self.torch.cuda.empty_cache()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def kernel_table(prof, pk):
    out = {}
    for name, rec in prof.items():
        if rec["launches"] == 0:
            continue
        entry = {"ms": round(rec["ms"], 4), "launches": rec["launches"]}
        if name in ("conv_tc", "conv_tc_upsample", "conv_direct", "attention"):
            entry["tflops"] = rec["work"] / 1e12 / (rec["ms"] / 1e3) if rec["ms"] > 0 else None
        else:
            gbs = rec["work"] / 1e9 / (rec["ms"] / 1e3) if rec["ms"] > 0 else None
            entry["gbs"] = gbs
            entry["hbm_frac"] = gbs / pk["hbm_gbs"] if gbs else None
        out[name] = entry
    return out


def conv_roofline(prof, pk, what, burst=False):
    """All tcgen05 conv / GEMM launches of one eagerly launched step (CUDA events per launch, rv_prof_*): algorithmic FLOPs
    (the up-sampling convs counted on the up-sampled grid, as the reference computes them) and EXECUTED FLOPs (those convs run
    phase-folded 2x2 kernels: 4/9 of the algorithmic MACs) over the summed launch durations."""
    conv, ups = prof["conv_tc"], prof["conv_tc_upsample"]
    ms = conv["ms"] + ups["ms"]
    n = conv["launches"] + ups["launches"]
    alg = conv["work"] + ups["work"]
    exe = conv["work"] + ups["work"] * 4.0 / 9.0
    ach = alg / 1e12 / (ms / 1e3) if ms > 0 else 0.0
    ach_x = exe / 1e12 / (ms / 1e3) if ms > 0 else 0.0
    tr = measured_traffic()
    # launches timed one by one with idle gaps between them (c5's eager decode) are held against the BURST figure of
    # MEASURED_PEAKS.json, launches of a back-to-back step against the sustained one
    peak = pk["tflops_burst"] if burst else pk["tflops"]
    return {"bound": "tensor", "kernel": what, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "achieved_executed": ach_x, "frac_executed": ach_x / peak,
            "traffic": tr.get("dram_bytes_per_launch") if tr else None, "traffic_detail": tr,
            "peak_source": f"{pk['source']} {'bf16_tflops (burst)' if burst else 'bf16_tflops_sustained'}"
                           + ("; 'achieved' counts the up-sampling convs on the up-sampled grid (9/4 of the MACs executed): frac_executed"
                              " is the utilisation figure" if ups["work"] > 0.2 * alg else ""),
            "launches_per_step": n, "ms_per_step": ms,
            "algorithmic_tflop_per_step": alg / 1e12, "executed_tflop_per_step": exe / 1e12,
            "per_launch": {"algorithmic_gflop": alg / 1e9 / n if n else None, "avg_us": ms * 1e3 / n if n else None}}


# ----------------------------------------------------------------------------------------------
# c2: the headline
# ----------------------------------------------------------------------------------------------
def c2_arm(cx: Ctx, a):
    torch = cx.torch
    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops

    world, rank, dev = cx.world, cx.rank, cx.dev
    B, S = a.batch, a.size
    torch.manual_seed(0)
    vae = R.RgbaAutoencoder(a.arch)  # torch default init == the oracle's
    model = R.RgbaVAE(vae.to(dev, torch.bfloat16))

    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.rand(B, 4, S, S, generator=g).to(torch.bfloat16).pin_memory()
    n_host = torch.randn(B, 16, S // 8, S // 8, generator=torch.Generator().manual_seed(2 + rank)).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    img_host = torch.empty(B, 4, S, S, dtype=torch.bfloat16).pin_memory()
    x_dev, n_dev = x_host.to(dev), n_host.to(dev)
    use_graph = not a.no_graph

    def step(x, noise):
        """-> (recon, metrics)"""
        if use_graph:  # one captured CUDA graph per shape: the same kernels, replayed without host gaps
            recon, _, met = model.forward_graphed(x, noise, backgrounds=((1.0, 1.0, 1.0),))
            return recon, met
        recon, _ = model(x, noise=noise)
        return recon, ops.composite_psnr(recon, x, [(1.0, 1.0, 1.0)])

    def resident():
        step(x_dev, n_dev)

    def end_to_end():  # validation: host batch in, per-sample metrics out
        xd = x_host.to(dev, non_blocking=True)
        nd = n_host.to(dev, non_blocking=True)
        out_host.copy_(step(xd, nd)[1], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def end_to_end_image():  # RgbaVAE.forward / inference: the reconstruction itself comes back too
        xd = x_host.to(dev, non_blocking=True)
        nd = n_host.to(dev, non_blocking=True)
        recon, met = step(xd, nd)
        img_host.copy_(recon, non_blocking=True)
        out_host.copy_(met, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    warm = max(a.warmup, 3)
    for _ in range(warm):
        resident()
    sampler = ClockSampler(cx.local).start() if rank == 0 else None
    ms_total = cx.timed(resident, a.steps)
    clocks = sampler.stop() if sampler else None
    # kernels per step, counted on one eagerly launched step (graph replays bypass the library's launch counter)
    l0 = ops.launch_count()
    recon_e, _ = model(x_dev, noise=n_dev)
    ops.composite_psnr(recon_e, x_dev, [(1.0, 1.0, 1.0)])
    torch.cuda.synchronize()
    del recon_e
    launches = (ops.launch_count() - l0) * a.steps
    end_to_end()
    ms_e2e = cx.timed(end_to_end, a.steps)
    psnr_white = float(out_host[:, 0].mean())
    end_to_end_image()
    ms_e2e_img = cx.timed(end_to_end_image, a.steps)

    # roofline of the dominant kernel family: CUDA events on the launching stream around every launch of one more
    # (eager) step -- per-kernel events cannot be recorded inside a graph replay
    prof = None
    if rank == 0:
        torch.cuda.synchronize()
        ops.prof_begin()
        recon_p, _ = model(x_dev, noise=n_dev)
        ops.composite_psnr(recon_p, x_dev, [(1.0, 1.0, 1.0)])
        prof = ops.prof_end()
        del recon_p
    model.reset_graphs()
    del model, vae
    cx.release()
    if rank != 0:
        return None
    pk = peaks()
    mpix_step = world * B * S * S / 1e6
    value = mpix_step * a.steps / (ms_total / 1e3)
    roof = conv_roofline(prof, pk, "conv_tc_kernel / conv_tc2_kernel (CTA pairs) / conv_halo_kernel: every tcgen05 implicit-GEMM "
                                   "conv and GEMM launch of one step")
    whole = world * B * TFLOP_PER_IMAGE_1024[a.arch] * a.steps / (ms_total / 1e3) if S == 1024 else None
    h2d = (x_host.numel() + n_host.numel()) * 2 * world
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(a),
            "e2e": {"value": mpix_step * a.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host.numel() * 4 * world,
                    "returns": "per-sample PSNR + alpha MAE (the validation loop's result)"},
            "e2e_with_image": {"value": mpix_step * a.steps / (ms_e2e_img / 1e3), "unit": UNIT, "ms_per_step": ms_e2e_img / a.steps,
                               "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": (img_host.numel() * 2 + out_host.numel() * 4) * world,
                               "returns": "the bf16 reconstruction (RgbaVAE.forward's result) + the metrics"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernel_table(prof, pk),
            "whole_step_tflops": whole, "psnr_white_db": psnr_white, "known_deviations": KNOWN_DEVIATIONS}


# ----------------------------------------------------------------------------------------------
# c3: mixed-aspect buckets
# ----------------------------------------------------------------------------------------------
def c3_arm(cx: Ctx, a, batches=None, steps=None):
    """Config c3: bucket-pure mixed-aspect batches (<= 1 MP, sides % 32 == 0) assigned to ranks
    longest-first; every rank validates its own batches with no data-path collective; the job time is the slowest
    rank's device time, per-sample PSNR vectors are gathered once after the run."""
    torch = cx.torch
    import ragb_vae_b200 as R
    from ragb_vae_b200 import sharding

    world, rank, dev = cx.world, cx.rank, cx.dev
    batches = batches or a.batches
    steps = steps or a.steps
    torch.manual_seed(0)
    model = R.RgbaVAE(R.RgbaAutoencoder(a.arch).to(dev, torch.bfloat16))
    shapes = sharding.sample_bucket_batches(batches, a.batch, seed=1234)
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
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        psnr = run_all()
    e1.record()
    torch.cuda.synchronize()
    ms = cx.max_ms(e0.elapsed_time(e1))
    allp = sharding.gather_per_sample(psnr)
    model.reset_graphs()
    del model, data
    cx.release()
    if rank != 0:
        return None
    mpix = sum(b * h * w for (b, h, w) in shapes) / 1e6
    loads = [sum(sharding.batch_cost(shapes[i]) for i in own) for own in sharding.assign_batches(shapes, world)]
    return {"metric": METRIC, "value": mpix * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": a.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"c3: {batches} bucket-pure mixed-aspect batches of {a.batch} (<= 1 MP, sides % 32), "
                                   "longest-first sharding, no collective", "arch": a.arch,
                       "bucket_shapes": len(set(shapes)), "mpix_per_step": mpix,
                       "l2": "every batch's activation working set (GBs) exceeds the 126 MB L2"},
            "load_imbalance": (max(loads) / (sum(loads) / len(loads))) if loads and sum(loads) > 0 else None,
            "samples_validated": int(allp.numel()), "psnr_white_db": float(allp.mean())}


# ----------------------------------------------------------------------------------------------
# c4: the training step
# ----------------------------------------------------------------------------------------------
def c4_arm(cx: Ctx, a, steps=None):
    """Config c4: the rgba_vae training step (src/training/rgba_vae_stage.py:433-523 without LPIPS): triplet, encode,
    sample, decode, AlphaVAE loss (loss_reduce_mean like configs/flux_vae.yaml) + 1e-6 KL, hand-written backward,
    bucketed NCCL gradient all-reduce, clip_grad_norm_(1.0) and AdamW(1e-5, betas (0.5, 0.9)).  Weak scaling: every
    rank trains on its own batch.  value = trained pixels per second over all ranks (inputs resident in HBM)."""
    torch, dist = cx.torch, cx.dist
    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops
    from ragb_vae_b200.trainer import VaeTrainStep

    world, rank, dev = cx.world, cx.rank, cx.dev
    steps = steps or a.steps
    torch.manual_seed(0)
    vae = R.RgbaAutoencoder(a.arch).to(dev, torch.bfloat16)   # --arch flux: the VAE configs/flux_vae.yaml:73 trains
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
    cx.barrier()
    l0 = ops.launch_count()
    sampler = ClockSampler(cx.local).start() if rank == 0 else None
    step.reducer.timing = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m = run(x, noise)
    e1.record()
    cx.barrier()
    step.reducer.timing = False
    exposed = step.reducer.exposed_ms()
    clocks = sampler.stop() if sampler else None
    launches = ops.launch_count() - l0
    if not a.no_graph:  # replayed launches are not seen by the host-side counter: one capture's count x steps
        launches = step.launches_per_replay * steps
    ms = cx.max_ms(e0.elapsed_time(e1))
    exposed_max = [cx.max_ms(v) for v in exposed]
    # per-category CUDA events need eager launches: one more (eager) step outside the timed region
    ops.prof_begin()
    step.step(x, noise)
    torch.cuda.synchronize()
    prof = ops.prof_end()
    # end to end: host batch in, loss out
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        m = run(x_host.to(dev, non_blocking=True), n_host.to(dev, non_blocking=True))
        loss_host = float(m["train/loss"])
    torch.cuda.synchronize()
    ms_e2e = cx.max_ms((time.perf_counter() - t0) * 1e3)
    in_sync = None
    if world > 1:  # data-parallel replicas must hold identical weights after the same number of steps
        chk = torch.stack([step.opt.master.double().sum(), step.opt.master.double().abs().sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        in_sync = all(bool(torch.equal(c, allc[0])) for c in allc)
    buckets = [(hi - lo) * 4 for lo, hi in step.reducer.buckets]
    order = sorted(range(len(buckets)), key=lambda b: (not step._bucket_is_decoder(*step.reducer.buckets[b]), b)) if world > 1 else []
    nbytes = step.reducer.bytes_per_step()
    del step, vae
    cx.release()
    if rank != 0:
        return None
    mpix = world * B * S * S / 1e6
    pk = peaks()
    kernels = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in prof.items() if v["launches"]}
    roof = conv_roofline(prof, pk, "every tcgen05 launch of one step: forward convs / GEMMs, data-gradient convs (conv_tc*, conv_halo) "
                                   "and conv_wgrad_kernel")
    lim = max(range(len(exposed_max)), key=lambda i: exposed_max[i]) if exposed_max else None
    comm = {"collective": "NCCL all-reduce (SUM) of the flat fp32 gradient, one per bucket, issued between the CUDA-graph replays "
                          "(decoder buckets right after graph 1, the deep encoder stages' after graph 2, so they overlap the rest of the backward; "
                          "only the small shallow-encoder bucket is issued after the last graph)",
            "bytes_per_step_per_rank": nbytes, "bucket_bytes": buckets, "world": world,
            "join_order_buckets": order, "exposed_ms_per_join": exposed_max,
            "exposed_ms_total": sum(exposed_max) if exposed_max else 0.0,
            "limiting_bucket": (order[lim] if lim is not None and lim < len(order) else None)}
    return {"metric": "rgba_vae_train_step_mpix_per_s", "value": mpix * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"c4: rgba_vae training step, batch {B} x {S}x{S} per GPU, recon (reduce_mean) + 1e-6 KL, "
                                   "no LPIPS, bucketed NCCL gradient all-reduce, clip 1.0, AdamW",
                       "arch": a.arch, "parallelism": f"data parallel x{world}",
                       "launch": "eager launches" if a.no_graph else "four CUDA graphs per step (forward + decoder backward | deep encoder backward | shallow encoder backward | clip + AdamW), all-reduces issued between the replays",
                       "l2": "multi-GB activation tape per step (inputs and working set exceed the 126 MB L2)"},
            "e2e": {"value": mpix * steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": (x_host.numel() + n_host.numel()) * 4 * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels, "loss": loss_host,
            "replicas_in_sync": in_sync, "allreduce": comm}


# ----------------------------------------------------------------------------------------------
# c5: Flux decode at 2048 x 2048
# ----------------------------------------------------------------------------------------------
def c5_arm(cx: Ctx, a, steps=None):
    """Config c5: Flux-arch RGBA decode of one 2048x2048 image per rank as inference_rgba_flux.py does
    (flux_kontext_textalpha.py:497-499): decode(z / scale + shift).sample -> (y + 1) / 2 -> clamp."""
    torch = cx.torch
    import ragb_vae_b200 as R
    from ragb_vae_b200 import ops

    world, rank, dev = cx.world, cx.rank, cx.dev
    steps = steps or a.steps
    S = 2048
    torch.manual_seed(0)
    vae = R.RgbaAutoencoder("flux").to(dev, torch.bfloat16)
    scale, shift = float(vae.config.scaling_factor), float(vae.config.shift_factor)
    zn = torch.randn(1, 16, S // 8, S // 8, generator=torch.Generator().manual_seed(3 + rank)).to(dev, torch.bfloat16)

    def run():
        return vae._decode_image(zn, out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0), z_scale=1.0 / scale, z_shift=shift)

    for _ in range(max(a.warmup, 3)):
        img = run()
    ms = cx.timed(run, steps)
    ok = bool(torch.isfinite(img.float()).all()) and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    prof = None
    if rank == 0:
        ops.prof_begin()
        run()
        prof = ops.prof_end()
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del vae, img
    cx.release()
    if rank != 0:
        return None
    pk = peaks()
    mpix = world * S * S / 1e6
    return {"metric": "flux_rgba_decode_mpix_per_s", "value": mpix * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "c5: Flux AutoencoderKL RGBA decode 1x16x256x256 -> 1x4x2048x2048 per GPU (eager launches)",
                       "arch": "flux", "l2": "256 MiB buffer rewritten between timed iterations"},
            "algorithmic_tflops": world * FLUX_DECODE_2048_TFLOP * steps / (ms / 1e3), "outputs_finite_in_range": ok,
            "peak_mem_gib": peak_mem, "roofline": conv_roofline(prof, pk, "every tcgen05 conv / GEMM launch of one decode", burst=True),
            "kernels": kernel_table(prof, pk)}


# ----------------------------------------------------------------------------------------------
# PyTorch eager / cuDNN on the same GPU (a baseline: the oracle's nn.Modules moved to CUDA in bf16)
# ----------------------------------------------------------------------------------------------
def gpu_eager_baseline(cx: Ctx, a):
    """What running the reference's own stack on this B200 would look like (SURVEY 2.1's bar): the oracle modules -- plain
    nn.Conv2d / F.scaled_dot_product_attention / elementwise ATen ops -- in bf16 through cuDNN.  NOT ours, not timed inside
    any of our regions; reported next to the headline."""
    torch = cx.torch
    if cx.rank != 0:
        return None
    from oracle import vae_oracle as O

    try:
        B, S = a.batch, a.size
        vae = O.build_oracle(a.arch, seed=0).to(cx.dev, torch.bfloat16)
        x = torch.rand(B, 4, S, S, generator=torch.Generator().manual_seed(1)).to(cx.dev, torch.bfloat16)
        noise = torch.randn(B, 16, S // 8, S // 8, generator=torch.Generator().manual_seed(2)).to(cx.dev, torch.bfloat16)
        torch.backends.cudnn.benchmark = True

        def run():
            with torch.no_grad():
                recon, _, _ = O.rgba_vae_forward(vae, x, noise)
                return O.validation_metrics(recon, x, backgrounds=(1.0,))[1.0]

        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        e0.record()
        for _ in range(steps):
            cx.flush_l2()
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del vae
        cx.release()
        return {"value": B * S * S / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "n_gpus": 1,
                "what": f"oracle nn.Modules on cuda:0 in bf16 (PyTorch {torch.__version__} eager, cuDNN {torch.backends.cudnn.version()}, "
                        f"cudnn.benchmark), batch {B} x {S}x{S}, {steps} steps after 2 warm-up; a baseline, none of our kernels"}
    except Exception as e:  # noqa: BLE001  (out of memory on a shared box etc.: the baseline is optional)
        cx.release()
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


def ensure_library():
    import __graft_entry__ as G

    B = G._load_builder()
    if not os.path.exists(B.LIB):  # normally prebuilt in-tree (it travels with the snapshot)
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            B.build()
        else:
            while not os.path.exists(B.LIB):
                time.sleep(1.0)
            time.sleep(2.0)


_STDOUT_FD = None


def claim_stdout():
    """ONE JSON line is the contract: whatever libraries print to fd 1 while the benchmark runs (NCCL's version banner under
    torchrun, for one) goes to stderr instead; emit() writes the line to the real stdout."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_STDOUT_FD, data)


def main():
    a = parse()
    claim_stdout()
    if a.impl == "reference":
        reference_arm(a)
        return
    ensure_library()
    cx = Ctx()
    try:
        if a.workload == "c3":
            line = c3_arm(cx, a)
        elif a.workload == "c4":
            line = c4_arm(cx, a)
        elif a.workload == "c5":
            line = c5_arm(cx, a)
        else:
            line = c2_arm(cx, a)
            extra = {}
            if not a.no_extras:
                few = max(2, min(a.steps, 5))
                extra["c4"] = c4_arm(cx, a, steps=few)
                extra["c3"] = c3_arm(cx, a, batches=max(16, 2 * cx.world), steps=1)
                if cx.world == 1:
                    extra["c5"] = c5_arm(cx, a, steps=few)
                    extra["gpu_eager_baseline"] = gpu_eager_baseline(cx, a)
            if cx.rank == 0:
                line["extra"] = extra
                if cx.world == 1 and not a.no_cpu_baseline:
                    line["cpu_baseline"], _ = run_cpu_baseline(a.arch, a.size, 4, 1, a.batch)  # ~10-15 s of host work on a bounded sample
                else:
                    line["cpu_baseline"] = None
        if cx.rank == 0 and line is not None:
            emit(line)
    finally:
        cx.close()


if __name__ == "__main__":
    main()
