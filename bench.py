#!/usr/bin/env python
"""Benchmark of the hot path: one DDPM + output-KD + feature-KD training step of the APTP-pruned SD-2.1 U-Net
(BASELINE.json config[1]): student r=0.55 (508.2 M params) + frozen full teacher (865.9 M), batch 16 per GPU,
64x64 latents (512 px), 77x1024 text context, bf16 kernels / fp32 masters, fused AdamW.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same step driven from
pinned host buffers through the public API (H2D of the batch + D2H of the loss inside the timed region).
`--impl reference` times the reference's own CPU path (the oracle restatement: the reference itself needs diffusers,
which cannot be installed here) on the host cores, on a bounded sample of the same workload.
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
try:        # DRAM bytes per launch of the roofline kernels, from the committed `ncu --set full` capture (not measurable in-process)
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _f:
        NCU_TRAFFIC = json.load(_f)
except OSError:
    NCU_TRAFFIC = {}


def ncu_traffic(key):
    t = NCU_TRAFFIC.get(key)
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) if t else None
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/sec @512px pruned U-Net (DDPM+KD)"
UNIT = "samples/s"
# Algorithmic matmul/conv/attention FLOPs per sample for the both-loss step (SURVEY.md section 8d / BASELINE.md 3):
# teacher fwd 0.804 + student fwd 0.464 + student bwd 2 * 0.464 (r = 0.55)
STEP_TFLOP_PER_SAMPLE = {0.55: 2.196, 0.82: 2.823}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--ratio", type=float, default=0.55)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip the stock-torch (cuDNN / cuBLAS / SDPA / fused AdamW) run of the oracle on the same GPU")
    ap.add_argument("--no-graph", action="store_true", help="single GPU: time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--sampling", action="store_true",
                    help="BASELINE config 5 instead of the default config 2: forward-only CFG sampling loop (U-Net only), "
                         "--images per GPU, 50 DDIM steps, guidance 7.5; a 'step' is one complete 50-step sampling call")
    ap.add_argument("--images", type=int, default=32, help="--sampling: images per GPU (U-Net batch = 2x)")
    ap.add_argument("--cpu-steps", type=int, default=4,
                    help="cpu_baseline: timed batch-1 oracle steps after one untimed step (about 10-15 s on the GPU box's host cores)")
    ap.add_argument("--bilevel", action="store_true",
                    help="BASELINE config 4: bilevel fine-tuning + concept suppression -- every --upper-freq lower (DDPM+KD) "
                         "steps one upper (ESD-style unlearning) step with its own AdamW state; samples/s over whole cycles")
    ap.add_argument("--upper-freq", type=int, default=10)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# clocks sampling (profiling guide recipe)
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference step) on the host cores
# --------------------------------------------------------------------------------------------------------------
def build_oracle_cpu(ratio, seed=43):
    import torch

    from oracle import diffusers_restated as D
    from oracle import pdm_restated as P

    def fast_init(m):
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() >= 2:
                    p.normal_(0.0, 0.02, generator=g)
                elif "norm" in n and n.endswith("weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()

    with torch.device("meta"):
        teacher = D.UNet2DConditionModel(**D.SD21_UNET_CONFIG)
        student = P.UNetGated()
    torch.manual_seed(seed)
    av = P.get_random_arch_vector(ratio, student.get_structure())
    teacher = teacher.to_empty(device="cpu")
    student = student.to_empty(device="cpu")
    fast_init(teacher), fast_init(student)
    student.set_structure(P.transform_arch_vector(av, student.get_structure()))
    student.prune()
    return teacher.eval().requires_grad_(False), student.eval(), av


def cpu_reference_run(ratio, latent, steps, warmup, batch=1):
    """Times `steps` full training steps (teacher fwd, student fwd+bwd, AdamW) of the oracle on all host cores.
    Bounded sample: batch 1 of the bench workload (same model, resolution, context, loss, optimiser)."""
    import torch

    from oracle import diffusers_restated as D
    from oracle import pdm_restated as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    teacher, student, _ = build_oracle_cpu(ratio)
    fs, ft = {}, {}
    P.cast_block_act_hooks(student, fs), P.cast_block_act_hooks(teacher, ft)
    sched = D.DDIMSchedulerLite()
    opt = P.adamw_reference(student.parameters())
    g = torch.Generator().manual_seed(43)
    times = []
    for i in range(warmup + steps):
        lat, noise = torch.randn(batch, 4, latent, latent, generator=g), torch.randn(batch, 4, latent, latent, generator=g)
        t = torch.randint(0, 1000, (batch,), generator=g)
        ehs = torch.randn(batch, 77, 1024, generator=g)
        t0 = time.perf_counter()
        loss, _, _, _ = P.finetune_step(student, teacher, sched, lat, noise, t, ehs, fs, ft)
        loss.backward()
        opt.step()
        opt.zero_grad()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} step(s) of batch {batch} (same model/resolution/loss/optimiser as the GPU workload; "
                      f"fp32, torch {torch.__version__} CPU ops, {torch.get_num_threads()} threads, anomaly detection off)",
            "ms_per_step": 1e3 * total / len(times)}


# --------------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------------
def gemm_roofline(torch, K, batch, latent):
    """Dominant kernel = the tcgen05 implicit-GEMM convolution.  Time the largest-FLOP student conv shape of the step
    (up_blocks.3 resnets.0 conv1: 960 -> 170 at 64x64, r = 0.55) in isolation with CUDA events on the launching
    stream, rotating inputs through > L2 (126 MB) worth of buffers."""
    B, H, W, Ci, Co = batch, latent, latent, 960, 170
    nbuf = 6
    xs = [K.alloc2d(B * H * W, Ci) for _ in range(nbuf)]     # 6 x 126 MB > L2
    for x in xs:
        x.normal_()
    w = torch.randn(Co, 9, Ci, device="cuda", dtype=torch.bfloat16) * 0.02
    bias = torch.zeros(Co, device="cuda")
    out = K.alloc2d(B * H * W, Co)
    for i in range(3):
        K.conv_fwd(xs[i % nbuf], w, B, H, W, Co, 3, 1, bias=bias, out=out)
    torch.cuda.synchronize()
    iters = 24
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record()
        K.conv_fwd(xs[i % nbuf], w, B, H, W, Co, 3, 1, bias=bias, out=out)
        ev[i][1].record()
    torch.cuda.synchronize()
    ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    flops = 2.0 * B * H * W * Co * 9 * Ci
    return flops / (ms * 1e-3) / 1e12, ms, f"conv3x3 {Ci}->{Co} @ {H}x{W} batch {B} (implicit GEMM M={B*H*W} N={Co} K={9*Ci})"


def attention_roofline(torch, K, batch, peaks):
    """Second tensor-class kernel: the fused tcgen05 attention forward at the step's dominant attention shape (teacher
    self-attention at 64x64: 5 heads, L = 4096, head_dim 64), timed alone with CUDA events.  FLOPs = 4 B h L^2 64."""
    B, H, L = batch, 5, 4096
    q, k, v = (K.alloc2d(B * L, H * 64).normal_() for _ in range(3))
    out = K.alloc2d(B * L, H * 64)
    for _ in range(3):
        K.attention_fwd(q, k, v, B, H, L, L, 0.125, out=out)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for e0, e1 in ev:
        e0.record()
        K.attention_fwd(q, k, v, B, H, L, L, 0.125, out=out)
        e1.record()
    torch.cuda.synchronize()
    ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    tf = 4.0 * B * H * L * L * 64 / (ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops", 1590.0)
    return {"bound": "tensor", "kernel": "b200::attn_fwd2_kernel (tcgen05 flash-style attention forward)",
            "shape": f"batch {B}, {H} heads, Lq = Lk = {L}, head_dim 64", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
            "frac": tf / peak, "ms_per_launch": ms, "traffic": ncu_traffic("attn_fwd_L4096_h5_b16") if batch == 16 else None}


def gemm_class_roofline(torch, _lib, one_eager_step, peaks):
    """Launch-weighted roofline of the whole GEMM class (every tcgen05 GEMM / implicit-GEMM conv launch of ONE step, forward
    and backward, teacher and student): sum of algorithmic FLOPs (2 M N K) / sum of per-launch CUDA-event times, measured
    live by the library's per-launch trace (serialised launches, so launch gaps are included: a lower bound)."""
    import ctypes as C
    L = _lib.lib()
    L.b200pdm_gemm_trace_enable(1)
    try:
        one_eager_step()
        torch.cuda.synchronize()
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        L.b200pdm_gemm_trace_totals(C.byref(ms), C.byref(fl), C.byref(n))
    finally:
        L.b200pdm_gemm_trace_enable(0)
    if ms.value <= 0:
        return None
    tf = fl.value / (ms.value * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops", 1590.0)
    return {"bound": "tensor", "kernel": "b200::gemm_kernel, all launches of one step (launch-weighted)", "achieved": tf,
            "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "gemm_ms_per_step": ms.value, "gemm_launches_per_step": n.value,
            "gemm_tflop_per_step": fl.value / 1e12, "traffic": None,
            "how": "per-launch CUDA events on the launching stream, launches serialised (one eager step after the timed region)"}


def library_baseline_run(args):
    """SURVEY section 8d 'library' reference point, driver-visible: the oracle (torch restatement of the reference step) on
    the SAME GPU through stock torch ops (cuDNN / cuBLAS / SDPA / fused torch AdamW, bf16 autocast, eager), in a
    subprocess after the timed region (tests/library_baseline.py).  Not the product path; never part of `value`."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "library_baseline.py"), "--batch", str(args.batch),
                            "--ratio", str(args.ratio), "--steps", "3", "--warmup", "2"], capture_output=True, text=True,
                           timeout=600, cwd=ROOT)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"value": None, "note": f"failed rc={r.returncode}: {r.stderr[-300:]}"}
        d = json.loads(line[-1])
        return {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "kind": "torch-library oracle on the same GPU",
                "what": d["what"], "steps": 3, "warmup": 2, "torch": d["torch"], "cudnn": d["cudnn"]}
    except Exception as e:  # pragma: no cover
        return {"value": None, "note": f"failed: {e!r}"}


def adamw_roofline(torch, K, student):
    """Second roofline point (HBM-bound class): the fused AdamW over the student's flat arena, timed alone with CUDA
    events.  Algorithmic bytes per parameter: 16 B read (p, g, m, v) + 12 B write (p, m, v) + 2 B bf16 shadow + 4 B
    gradient zeroing = 34 B."""
    a = student.arena
    m, v = torch.zeros_like(a.master), torch.zeros_like(a.master)
    p = a.master.detach().clone()
    g = torch.zeros_like(p)
    sh = torch.empty_like(a.shadow)
    for i in range(2):
        K.adamw_step(p, g, m, v, sh, 1e-6, 0.9, 0.999, 1e-8, 0.0, i + 1)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for i, (e0, e1) in enumerate(ev):
        e0.record()
        K.adamw_step(p, g, m, v, sh, 1e-6, 0.9, 0.999, 1e-8, 0.0, i + 3)
        e1.record()
    torch.cuda.synchronize()
    ms = statistics.median(e0.elapsed_time(e1) for e0, e1 in ev)
    bytes_ = 34.0 * a.numel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    gbs = bytes_ / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "b200::adamw_kernel (flat multi-tensor AdamW + bf16 shadow + grad zeroing)",
            "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "ms_per_launch": ms,
            "bytes_per_launch": bytes_, "traffic": None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs, of measured" if peaks else "fallback 6650 GB/s, of fallback"}


def sampling_bench(args, rank, world, local_rank):
    """BASELINE config 5: images/s of the guided denoising loop, U-Net only.  Ranks are independent
    replicas (each samples its own images; no collective on this path)."""
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from unlearn_ft_b200 import _lib
    from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
    from unlearn_ft_b200.pdm.pipelines import CFGSampler

    torch.manual_seed(43)
    av = HyperStructure.get_random_arch_vector(args.ratio, structure_from_config(SD21_CONFIG))
    unet = UNet2DConditionModelPruned(arch_vector=av, trainable=False, seed=43)
    n, L, steps_inf = args.images, args.latent, 50
    sampler = CFGSampler(unet, num_inference_steps=steps_inf, guidance_scale=7.5, use_cuda_graph=not args.no_graph)
    g = torch.Generator().manual_seed(2000 + rank)
    host = [(torch.randn(n, 4, L, L, generator=g).pin_memory(), torch.randn(n, 77, 1024, generator=g).bfloat16().pin_memory(),
             torch.randn(1, 77, 1024, generator=g).bfloat16().expand(n, -1, -1).contiguous().pin_memory()) for _ in range(2)]
    dev = [tuple(t.cuda() for t in hb) for hb in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n0 = _lib.launch_count()
    sampler.use_cuda_graph, keep = False, sampler.use_cuda_graph
    sampler.evals, full = 1, sampler.evals
    sampler.sample(*dev[0])                                       # one eager step: kernels per denoising step
    launches_per_call = (_lib.launch_count() - n0) * full
    sampler.evals, sampler.use_cuda_graph = full, keep
    for i in range(max(1, min(args.warmup, 2))):
        sampler.sample(*dev[i % 2])
    barrier()
    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        sampler.sample(*dev[i % 2])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    barrier()
    e0.record()
    for i in range(args.steps):
        lat, pos, neg = host[i % 2]
        out = sampler.sample(lat.cuda(non_blocking=True), pos.cuda(non_blocking=True), neg.cuda(non_blocking=True))
        res = out.cpu()                                           # final latents back to the host (what the VAE would take)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    clocks = clk.stop() if rank == 0 else None
    if rank == 0:
        images = world * n * args.steps
        fl = 0.464e12 * 2 * steps_inf if abs(args.ratio - 0.55) < 1e-6 else None      # per image
        line = {"metric": "sampled images/sec (U-Net only, 50-step CFG DDIM, 512px)", "value": images / (float(ms) * 1e-3),
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(ms) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE config 5: APTP r={args.ratio} pruned SD-2.1 U-Net, forward only, {n} images/GPU "
                                       f"(U-Net batch {2 * n}), 50 DDIM steps, guidance 7.5, {L}x{L} latent; VAE / text encoder "
                                       "excluded", "parallelism": f"replicas x{world}",
                           "step_launch": "one CUDA graph replay per denoising step" if sampler.use_cuda_graph else "eager",
                           "l2": "no flush: weights + activations of a batch-64 forward >> 126 MB L2"},
                "e2e": {"value": images / (float(ms2) * 1e-3), "unit": "images/s",
                        "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]),
                        "d2h_bytes_per_step": n * 4 * L * L * 4},
                "gpu_launches": int(launches_per_call * args.steps), "clocks": clocks,
                "algorithmic_tflops": None if fl is None else fl * images / (float(ms) * 1e-3) / 1e12}
        emit(line)
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def own_stdout():
    """stdout carries exactly ONE JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr, so that whatever
    libraries print from C (NCCL writes "NCCL version ..." to stdout at the first collective) cannot precede the result."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    args = parse()
    own_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.sampling and args.impl != "reference":
        return sampling_bench(args, rank, world, local_rank)
    workload = (f"APTP r={args.ratio} pruned SD-2.1 U-Net (random arch vector, random init) + frozen SD-2.1 teacher, "
                f"DDPM+output-KD+feature-KD step, batch {args.batch}/GPU, {args.latent}x{args.latent} latent, bf16")

    if args.impl == "reference":
        if rank != 0:
            return
        # a "step" of this arm = one batch-1 training step of the same workload (bounded sample: ~2 s of CPU work each); the
        # requested K / W are honoured up to a 60-step total so the run stays within a few minutes, and the line reports
        # what actually ran
        k_run = max(1, min(args.steps, 40))
        w_run = max(0, min(args.warmup, 60 - k_run, 5))
        res = cpu_reference_run(args.ratio, args.latent, k_run, w_run)
        ref_workload = (f"bounded sample of: {workload} -- here ONE sample per step (batch 1), fp32, CPU: "
                        f"{res['cores']} host threads, oracle restatement of the reference step")
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": k_run, "warmup": w_run, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": ref_workload, "batch_per_step": 1, "requested_steps": args.steps,
                           "requested_warmup": args.warmup,
                           "note": "reference CPU path = oracle restatement (the reference itself needs diffusers, not "
                                   "installable offline)"},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from unlearn_ft_b200 import _lib
    from unlearn_ft_b200 import kernels as K
    from unlearn_ft_b200.pdm.models import HyperStructure, UNet2DConditionModel, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.models.unet.unet_2d_conditional import SD21_CONFIG, structure_from_config
    from unlearn_ft_b200.pdm.training import BilevelUnetFineTuner, UnetFineTuner

    torch.manual_seed(43)                                           # same arch vector on every rank
    av = HyperStructure.get_random_arch_vector(args.ratio, structure_from_config(SD21_CONFIG))
    student = UNet2DConditionModelPruned(arch_vector=av, seed=43)   # random_init path (configs/*_random.yaml:27)
    teacher = UNet2DConditionModel(seed=44)
    B, L = args.batch, args.latent
    freq = max(1, args.upper_freq)
    if args.bilevel:
        # reference bilevel config: lower lr 1e-6 with warm-up, upper lr 5e-6, upper step every 10th lower step
        tuner = BilevelUnetFineTuner(student, teacher, lr=1e-6, warmup_steps=250, upper_lr=5e-6, upper_step_freq=freq)
        args.steps = -(-args.steps // freq) * freq                 # whole cycles: any K = n*freq consecutive steps hold n upper steps
        workload = (f"bilevel fine-tuning + concept suppression: APTP r={args.ratio} pruned SD-2.1 U-Net + frozen SD-2.1 teacher, "
                    f"{freq} lower DDPM+KD steps then 1 upper step (teacher cond/uncond, target 2*uncond-cond, second AdamW), "
                    f"batch {args.batch}/GPU, {args.latent}x{args.latent} latent, bf16")
    else:
        tuner = UnetFineTuner(student, teacher, lr=1e-6, warmup_steps=250)
    g = torch.Generator().manual_seed(1000 + rank)                  # independent data per rank

    def host_batch():
        return dict(latents=torch.randn(B, 4, L, L, generator=g).pin_memory(),
                    noise=torch.randn(B, 4, L, L, generator=g).pin_memory(),
                    timesteps=torch.randint(0, 1000, (B,), generator=g).pin_memory(),
                    prompt_embeds=torch.randn(B, 77, 1024, generator=g).bfloat16().pin_memory())

    host_batches = [host_batch() for _ in range(4)]
    dev_batches = [{k: v.cuda(non_blocking=True) for k, v in hb.items()} for hb in host_batches]
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
    host_upper = dev_upper = None
    if args.bilevel:                                                # upper batches: own data + the "" prompt embedding for every sample
        empty = torch.randn(1, 77, 1024, generator=g).bfloat16()
        host_upper = [dict(host_batch(), empty_prompt_embeds=empty.expand(B, 77, 1024).contiguous().pin_memory()) for _ in range(2)]
        dev_upper = [{k: v.cuda(non_blocking=True) for k, v in hb.items()} for hb in host_upper]
        h2d_upper = sum(v.numel() * v.element_size() for v in host_upper[0].values())

    def train_step(i, batches, upper):
        if args.bilevel:
            return tuner.train_step(batches[i % len(batches)], upper[i % len(upper)])
        return tuner.train_step(batches[i % len(batches)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (eager), and the number of kernels one step launches
    for i in range(max(args.warmup, 3)):
        n_before = _lib.launch_count()
        train_step(i, dev_batches, dev_upper)
        launches_per_step = _lib.launch_count() - n_before
    if args.bilevel:                                                # one whole eager cycle: launches of freq lower + 1 upper step
        n_before = _lib.launch_count()
        for i in range(freq):
            train_step(i, dev_batches, dev_upper)
        launches_per_step = (_lib.launch_count() - n_before) / freq
    barrier()
    # The whole step (forward x2, backward with its overlapped per-block NCCL all-reduces, AdamW: ~2300 launches) is replayed
    # from ONE CUDA graph, so the host costs microseconds per step and a per-step result read-back cannot starve the GPU.
    # --no-graph measures the eager path.
    use_graph = not args.no_graph
    if use_graph:
        try:
            if args.bilevel:
                tuner.capture_cuda_graph(dev_batches[0], dev_upper[0])
            else:
                tuner.capture_cuda_graph(dev_batches[0])
            for i in range(freq if args.bilevel else 2):
                train_step(i, dev_batches, dev_upper)
        except Exception as e:                                     # report it instead of silently changing the metric
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); using the eager step", file=sys.stderr)
            tuner._graph = None
            use_graph = False
    barrier()

    # ---- timed: device-resident inputs (activations + parameters >> L2, so no L2 flush is needed between steps)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        train_step(i, dev_batches, dev_upper)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    launches = launches_per_step * args.steps      # (graph replays do not pass through the C entry points' counter)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)

    # ---- timed: end to end from pinned host memory through the public API
    def e2e_step(i):
        hb = host_batches[i % len(host_batches)]
        to_dev = (lambda b: b) if use_graph else (lambda b: {k: v.cuda(non_blocking=True) for k, v in b.items()})
        if args.bilevel:                                        # (the upper batch is only copied on the steps that use it)
            ub = to_dev(host_upper[i % len(host_upper)]) if (tuner.global_step + 1) % freq == 0 else None
            loss, _, _, _ = tuner.train_step(to_dev(hb), ub)
            if tuner.last_upper is not None:
                return float(loss) + 0.0 * float(tuner.last_upper[0])   # both steps' results are read back
        else:
            loss, _, _, _ = tuner.train_step(to_dev(hb))        # graph: H2D straight from pinned memory into its input buffers
        return float(loss)                                      # D2H read of the step's result (4 bytes) + sync

    e2e_step(0)                                                 # untimed: first-use effects of this code path (allocator)
    barrier()
    e0.record()
    last = None
    for i in range(args.steps):
        last = e2e_step(i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2)
    clocks = sampler.stop() if rank == 0 else None

    def shutdown():
        """Drop the captured graph (it references the NCCL communicator), then tear the process group down; a teardown that
        does not return within 20 s is abandoned (the result line has been printed by then)."""
        tuner.release_cuda_graph()
        torch.cuda.synchronize()
        if world > 1:
            import threading
            t = threading.Thread(target=dist.destroy_process_group, daemon=True)
            t.start()
            t.join(20)
            if t.is_alive():
                sys.stdout.flush()
                os._exit(0)

    if rank != 0:
        shutdown()
        return

    samples = world * B * args.steps
    if args.bilevel:
        samples += world * B * (args.steps // freq)             # the upper steps' batches
        h2d = h2d + h2d_upper / freq
    value = samples / (ms_total * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf, conv_ms, conv_desc = gemm_roofline(torch, K, B, L)
    hbm = adamw_roofline(torch, K, student)
    attn_rf = attention_roofline(torch, K, B, peaks)
    gemm_class = None
    if world == 1:      # (an eager step on rank 0 alone would wait for the other ranks' all-reduces)
        tuner.release_cuda_graph()
        gemm_class = gemm_class_roofline(torch, _lib, lambda: train_step(0, dev_batches, dev_upper), peaks)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    step_tf = STEP_TFLOP_PER_SAMPLE.get(round(args.ratio, 2), 2.196) * B / (ms_total / args.steps * 1e-3) if ms_total else 0
    if args.bilevel:   # upper step = 2 teacher fwd + student fwd + bwd = 2*0.804 + 3*fwd_s TFLOP/sample
        fwd_s = (STEP_TFLOP_PER_SAMPLE.get(round(args.ratio, 2), 2.196) - 0.804) / 3
        cyc = freq * STEP_TFLOP_PER_SAMPLE.get(round(args.ratio, 2), 2.196) + 2 * 0.804 + 3 * fwd_s
        step_tf = cyc * B * (args.steps // freq) / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": world * B, "parallelism": f"dp{world}",
                   "student_params": student.num_parameters(), "teacher_params": teacher.num_parameters(),
                   "l2": "no flush: per-step working set (1.4 B params x 2-4 B + activations) >> 126 MB L2",
                   "last_loss": last,
                   "step_launch": "one CUDA graph replay per step (capture of step+backward+AdamW)" if use_graph
                   else "eager: one launch per kernel"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline_hbm": hbm,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": tf / peak_burst,
                     # not measurable from inside the process: dram__bytes_read.sum + dram__bytes_write.sum of this launch
                     # from the committed `ncu --set full` capture (profiles/ncu_traffic.json names its source file)
                     "traffic": ncu_traffic("conv3x3_960_170_b16") if B == 16 else None,
                     "traffic_source": NCU_TRAFFIC.get("source"),
                     "algorithmic_bytes": 2.0 * (B * L * L * 960 + 170 * 9 * 960 + B * L * L * 170),
                     "kernel": "b200::gemm_kernel<0,0,true> (tcgen05 cta_group::2 implicit-GEMM conv)",
                     "shape": conv_desc,
                     "ms_per_launch": conv_ms,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone), of measured"
                     if peaks else "fallback 1590 TFLOP/s, of fallback",
                     "step_algorithmic_tflops": step_tf,
                     "step_frac_of_sustained": step_tf / peaks.get("bf16_tflops_sustained", 1400.0)},
        "roofline_step": gemm_class,
        "roofline_attention": attn_rf,
    }
    if args.bilevel:
        line["config"]["cycle"] = (f"{freq} lower + 1 upper step per cycle, {args.steps // freq} cycle(s) timed; value counts the "
                                   f"samples of both kinds of step; ms_per_step = cycle time / {freq}")
        line["config"]["step_launch"] += " (lower and upper step are two graphs sharing one memory pool)" if use_graph else ""
    if not args.no_cpu_baseline and world == 1 and not args.bilevel:
        try:
            res = cpu_reference_run(args.ratio, L, args.cpu_steps, 1)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    if not args.no_library_baseline and world == 1 and not args.bilevel:
        line["library_baseline"] = library_baseline_run(args)
    emit(line)
    shutdown()


if __name__ == "__main__":
    main()
