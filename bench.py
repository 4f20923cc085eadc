#!/usr/bin/env python
"""SPEGNet inference throughput on B200 (BASELINE.json metric: images/sec at default resolution).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--size 512] [--dtype fp16|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's algorithm on the host CPU (oracle port)

One "step" = one SPEGNet forward over one batch of synthetic images (BASELINE config 2: batch 64, 512x512,
random-init weights).  `value` is device-timed (CUDA events, inputs resident in HBM); `e2e` is the same
metric through the public drop-in call with pinned HOST images in and HOST logits out.  `roofline` is
for the dominant kernel (the tcgen05 GEMM / implicit-GEMM conv engine): executed FLOPs of its launches /
their CUDA-event durations, against the measured sustained bf16 peak in MEASURED_PEAKS.json.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION in this image) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

GFLOP_PER_IMAGE = {512: 610.98, 1024: 2530.89}  # algorithmic, reference formulation
# DRAM traffic of one launch of the dominant kernel from an `ncu --set full` capture (profiles/
# r01_gemm_qkv_ncu_full_summary.txt): the stage-3 QKV GEMM at batch 64 (M=65536, N=1728, K=576),
# dram__bytes_read.sum 77.6 MB (algorithmic 77.5 MB: every operand byte is read exactly once) +
# dram__bytes_write.sum 173.9 MB (algorithmic 226.5 MB; the remainder is still in the 126 MB L2 at kernel end).
NCU_TRAFFIC = {"bytes_per_launch": 77_608_192 + 173_853_952, "launch": "stage-3 QKV GEMM, M=65536 N=1728 K=576",
               "algorithmic_bytes": 65536 * 576 * 2 + 1728 * 576 * 2 + 65536 * 1728 * 2,
               "source": "profiles/r01_gemm_qkv_ncu_full_summary.txt"}
FALLBACK_PEAKS = {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
CFG = {"encoder": {"config_path": "configs/sam2.1/sam2.1_hiera_l.yaml",
                   "checkpoint_path": "./checkpoints/sam2.1_hiera_large.pt", "variant": "large"}}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, maxes, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                clocks.append(float(r[1]))
                maxes.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(clocks) if clocks else None,
                "sm_max_mhz": max(maxes) if maxes else None, "reasons": sorted(reasons), "samples": len(clocks)}


def cpu_reference_rate(state_dict, size: int, budget_s: float, max_images: int):
    """The reference's algorithm on the host cores: oracle port (fp32, batch 1, all threads), the one
    place bench.py executes oracle/ (checker / baseline only, never the measured GPU path)."""
    import torch

    from oracle.spegnet import spegnet_forward

    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(1, 3, size, size, generator=torch.Generator().manual_seed(0))
    spegnet_forward(state_dict, x)  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < max_images and (time.perf_counter() < t_end or len(times) < 2):
        t0 = time.perf_counter()
        spegnet_forward(state_dict, x)
        times.append(time.perf_counter() - t0)
    return {"value": round(len(times) / sum(times), 4), "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": "port", "p50_ms": round(statistics.median(times) * 1e3, 1),
            "sample": f"{len(times)} fp32 forwards, batch 1, {size}x{size}, oracle port of SPEGNet.forward"}


def run_reference(args):
    """--impl reference: /root/reference cannot travel to the GPU box and its trunk (sam2) is not installable
    offline, so this arm times the oracle port of the same forward on the host CPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from spegnet_b200 import SPEGNet

    torch.manual_seed(0)
    sd = SPEGNet(CFG).state_dict()
    per_step, total = [], 0
    from oracle.spegnet import spegnet_forward

    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(1, 3, args.size, args.size, generator=torch.Generator().manual_seed(0))
    for _ in range(max(1, min(args.warmup, 2))):
        spegnet_forward(sd, x)
    steps = max(1, min(args.steps, 12))
    for _ in range(steps):
        t0 = time.perf_counter()
        spegnet_forward(sd, x)
        per_step.append(time.perf_counter() - t0)
        total += 1
    value = total / sum(per_step)
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": round(value, 4), "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
        "ms_per_step": round(1e3 * sum(per_step) / steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SPEGNet inference, {args.size}x{args.size}, random-init weights; reference arm: "
                               "one fp32 image per step on the host CPU (bounded sample of the batch-64 workload)"},
        "cpu_baseline": {"value": round(value, 4), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} fp32 forwards, batch 1"},
        "e2e": {"value": round(value, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_dataset(args):
    """--dataset N: BASELINE config 4 -- batch-sharded evaluation of an N-image synthetic set (COD10K-test size: 2026)
    with the five scores computed on the GPU; one all_gather of [ceil(N/W), 6] fp64 rows per dataset."""
    import torch

    from spegnet_b200 import SPEGNet, _lib, evaluate, sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        _init_nccl(dist, dev)
    torch.manual_seed(0)
    model = SPEGNet(CFG, compute_dtype=torch.float16 if args.dtype == "fp16" else torch.bfloat16).to(dev).eval()
    N, B, S = args.dataset, args.batch, args.size
    mine = sharded.shard_indices(N, rank, world)
    gen = evaluate.synthetic_batch_fn(S, dev)
    # this rank's shard, resident in HBM before the timed region (inputs are synthetic; loading is out of scope)
    cache = {}
    for s0 in range(0, len(mine), B):
        idx = mine[s0:s0 + B]
        cache[tuple(idx)] = gen(idx)
    cache[()] = gen([])

    def batch_fn(idx):
        return cache[tuple(idx)]

    with torch.no_grad():
        evaluate.score_batch(model, *batch_fn(mine[:B]))  # warm-up: weight repack, workspaces
        evaluate.score_batch(model, *batch_fn(mine[:B]))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        _lib.reset_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        result = evaluate.evaluate_dataset(model, N, B, batch_fn)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if dist is not None:
            tmax = torch.tensor([ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
    if rank == 0:
        scores = {k: round(float(result[k]), 6) for k in ("s_alpha", "weighted_f", "mae", "e_phi", "mean_f")}
        print(json.dumps({
            "metric": "images_per_sec", "value": round(N / (ms * 1e-3), 2), "unit": "images/s", "n_gpus": world,
            "steps": 1, "warmup": 2, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"batch-sharded evaluation of a {N}-image synthetic set at {S}x{S} with on-GPU "
                                   "S-alpha / weighted F-beta / E-phi / MAE / mean F-beta (BASELINE config 4)",
                       "batch_per_gpu": B, "size": S, "parallelism": f"batch-sharded x{world}",
                       "collective": "one all_gather of [ceil(N/W), 6] fp64 rows"},
            "scores": scores, "gpu_launches": int(_lib.launch_count())}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


class _StdoutToStderr:
    """File-descriptor level redirect: NCCL prints its version banner straight to fd 1 while the communicator is
    created, and stdout has to carry exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def _init_nccl(dist, dev):
    with _StdoutToStderr():
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()  # creates the communicator (and its banner) now, not inside the timed region


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step (BASELINE config 2: 64)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dtype", default=os.environ.get("SPEGNET_B200_DTYPE", "fp16"), choices=["fp16", "bf16"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of host CPU time for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--dataset", type=int, default=0, help="evaluate an N-image synthetic set instead (BASELINE config 4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.dataset > 0:
        return run_dataset(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch

    from spegnet_b200 import SPEGNet, _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: spegnet_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        _init_nccl(dist, dev)
    lib = _lib.load(args.dtype)
    if lib.spg_device_check() != 0:
        raise SystemExit(lib.spg_last_error().decode())

    torch.manual_seed(0)  # identical random-init weights on every rank
    model = SPEGNet(CFG, compute_dtype=torch.float16 if args.dtype == "fp16" else torch.bfloat16)
    cpu_sd = {k: v.clone() for k, v in model.state_dict().items()} if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    model = model.to(dev).eval()

    B, S = args.batch, args.size
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # three different resident batches, rotated, so that no step re-reads its predecessor's inputs; the
    # per-step activation working set (~13 GB at B=64) is two orders of magnitude beyond the 126 MB L2.
    batches = [torch.randn(B, 3, S, S, device=dev, generator=gen) for _ in range(3)]

    gt = (torch.rand(B, S, S, device=dev, generator=gen) > 0.75).to(torch.uint8) * 255  # synthetic ground truth

    def step(i):
        out = model(batches[i % 3])
        # per-image predictions -> uint8 masks + integer MAE partials on the GPU (utils/metrics.py:205-210), and the
        # only collective of the path: one all_gather of 5 integers per image
        _, stats = ops.mask_stats(out["predictions"][-1], gt, True)
        if dist is not None:
            gathered = torch.empty(world * B, stats.shape[1], dtype=stats.dtype, device=dev)
            dist.all_gather_into_tensor(gathered, stats)
            return gathered
        return stats

    # ---- instrumented GEMM/conv timing (dominant kernel) -------------------------------------------------
    gemm_events = []
    real_linear, real_conv, real_conv_up2 = ops.linear, ops.conv3x3, ops.conv3x3_up2

    def timed_linear(a, w, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_linear(a, w, out, **kw)
        e1.record()
        k_true = 147 if w.shape[1] == 168 else w.shape[1]
        gemm_events.append((e0, e1, 2.0 * a.shape[0] * w.shape[0] * k_true))

    def timed_conv(x, w, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_conv(x, w, out, **kw)
        e1.record()
        gemm_events.append((e0, e1, 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * w.shape[0] * w.shape[1]))

    def timed_conv_up2(x, w_phase, corr, bias4, out):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_conv_up2(x, w_phase, corr, bias4, out)
        e1.record()
        # 4 phases x Cout outputs per low-resolution pixel over K = 9*Cin: the FLOPs of the 3x3 conv on the 2x grid
        gemm_events.append((e0, e1, 2.0 * x.shape[0] * x.shape[1] * x.shape[2] * (w_phase.shape[0] // 3) * w_phase.shape[1]))

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        _lib.reset_launch_count()
        ops.linear, ops.conv3x3, ops.conv3x3_up2 = timed_linear, timed_conv, timed_conv_up2
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for i in range(args.steps):
            step(i)
        t1.record()
        torch.cuda.synchronize()
        ops.linear, ops.conv3x3, ops.conv3x3_up2 = real_linear, real_conv, real_conv_up2
        if dist is not None:
            dist.barrier()
        launches = _lib.launch_count()
        clocks = sampler.stop() if rank == 0 else None
        elapsed_ms = t0.elapsed_time(t1)
        if dist is not None:
            tmax = torch.tensor([elapsed_ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            elapsed_ms = float(tmax.item())
        gemm_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in gemm_events)
        gemm_flops = sum(f for _, _, f in gemm_events)
        n_gemm = len(gemm_events)

        # ---- e2e: pinned host images in, host logits out, through the public host-to-host call -----------------
        # spegnet_b200.HostPipeline: every step's 201 MB host->device copy and 68 MB device->host read are inside the
        # timed region, on their own streams (copy of batch i+1 / read-back of batch i-1 overlap the kernels of batch i)
        from spegnet_b200.pipeline import HostPipeline

        host_in = [torch.randn(B, 3, S, S).pin_memory() for _ in range(2)]
        pipe = HostPipeline(model, dev)
        checksum = 0.0
        for out in pipe.run(host_in[i % 2] for i in range(3)):
            checksum += float(out["prediction"][0, 0, 0, 0])  # touch the host result
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e2e_steps = max(3, min(args.steps, 10))
        w0 = time.perf_counter()
        for out in pipe.run(host_in[i % 2] for i in range(e2e_steps)):
            checksum += float(out["prediction"][0, 0, 0, 0])
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - w0) * 1e3
        if dist is not None:
            tmax = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_ms = float(tmax.item())

        # ---- p50 batch-1 latency (the second half of BASELINE's metric) ---------------------------------
        latency = None
        if rank == 0 and not args.no_latency:
            x1 = batches[0][:1].contiguous()
            for _ in range(5):
                model(x1)
            torch.cuda.synchronize()
            lat = []
            for _ in range(30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                model(x1)
                b.record()
                torch.cuda.synchronize()
                lat.append(a.elapsed_time(b))
            lat.sort()
            latency = {"p50_ms": round(lat[len(lat) // 2], 3), "p90_ms": round(lat[int(len(lat) * 0.9)], 3), "batch": 1}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    images = world * B * args.steps
    value = images / (elapsed_ms * 1e-3)
    ms_per_step = elapsed_ms / args.steps
    gflop_img = GFLOP_PER_IMAGE.get(S, GFLOP_PER_IMAGE[512] * (S / 512.0) ** 2)
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) * 1e-12
    model_tf = gflop_img * 1e9 * B / (ms_per_step * 1e-3) * 1e-12  # per GPU
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)

    line = {
        "metric": "images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"SPEGNet inference forward, batch {B}/GPU, {S}x{S}, Hiera-L trunk + CFI/EFE/PED head, "
                               "random-init weights (BASELINE config 2)",
                   "batch_per_gpu": B, "size": S, "parallelism": f"batch-sharded x{world}",
                   "l2": "3 rotating input batches; per-step activation working set >> 126 MB L2",
                   "storage_dtype": args.dtype, "accumulate": "fp32 (TMEM)", "residual_stream": "fp32"},
        "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S * 4,
                "d2h_bytes_per_step": B * (S * S + (S // 8) ** 2) * 4, "steps": e2e_steps,
                "call": "HostPipeline(model).run(pinned host batches) -> pinned host logits + edge maps, copies on their own streams"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all Linear / 1x1 / 3x3-conv launches)",
                     "achieved": round(achieved_tf, 1), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved_tf / peak_tf, 4),
                     "traffic": NCU_TRAFFIC["bytes_per_launch"] if (B == 64 and S == 512) else None,
                     "traffic_detail": NCU_TRAFFIC, "peak_source": f"{peak_src} sustained bf16",
                     "launches_per_step": n_gemm // max(args.steps, 1), "kernel_ms_per_step": round(gemm_ms / args.steps, 3),
                     "share_of_step": round(gemm_ms / elapsed_ms, 4)},
        "model_roofline": {"gflop_per_image": gflop_img, "achieved_tflops_per_gpu": round(model_tf, 1),
                           "frac_of_sustained_peak": round(model_tf / peak_tf, 4),
                           "note": "algorithmic FLOPs of the reference formulation (SURVEY.md 8(d)) / whole step time"},
        "clocks": clocks,
    }
    if latency is not None:
        line["latency_b1"] = latency
    if cpu_sd is not None:
        line["cpu_baseline"] = cpu_reference_rate(cpu_sd, S, args.cpu_budget, 12)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
