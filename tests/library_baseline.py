"""'Library' reference point asked for by SURVEY.md §8(d): the ORACLE (torch restatement of the reference step) run on the
B200 through stock torch ops -- cuDNN convolutions, cuBLAS linears, SDPA attention, torch.optim.AdamW(fused) -- under the
bf16 autocast wrapper accelerate would apply.  This is the closest stand-in for "the reference itself on a
B200" that can run offline (diffusers is not installable).  It is test infrastructure: NOT collected by pytest, not the
product path, and clearly labelled in its output.  Same workload as bench.py (config 2: r = 0.55 student + teacher, batch
16, 64x64 latents, both-loss step + AdamW).

    python tests/library_baseline.py [--batch 16] [--steps 5] [--channels-last]
prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--ratio", type=float, default=0.55)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--channels-last", action="store_true")
    args = ap.parse_args()
    from bench import build_oracle_cpu
    from oracle import diffusers_restated as D
    from oracle import pdm_restated as P
    teacher, student, _ = build_oracle_cpu(args.ratio)
    teacher = teacher.cuda().to(torch.bfloat16)          # frozen teacher in bf16 (trainer.py:2145-2149 weight_dtype)
    student = student.cuda()                             # fp32 master weights, bf16 autocast compute
    if args.channels_last:
        teacher = teacher.to(memory_format=torch.channels_last)
        student = student.to(memory_format=torch.channels_last)
    fs, ft = {}, {}
    P.cast_block_act_hooks(student, fs), P.cast_block_act_hooks(teacher, ft)
    sched = D.DDIMSchedulerLite()
    for name in ("alphas_cumprod",):
        if hasattr(sched, name):
            setattr(sched, name, getattr(sched, name).cuda())
    opt = torch.optim.AdamW(student.parameters(), lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused=True)

    class AC(torch.nn.Module):                            # accelerate's autocast wrapper: bf16 inside, fp32 .sample out
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, *a):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = self.m(*a)
            out.sample = out.sample.float()
            return out

    s_ac, t_ac = AC(student), AC(teacher)
    B, L = args.batch, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    batches = [(torch.randn(B, 4, L, L, device="cuda", generator=g), torch.randn(B, 4, L, L, device="cuda", generator=g),
                torch.randint(0, 1000, (B,), device="cuda", generator=g),
                torch.randn(B, 77, 1024, device="cuda", generator=g)) for _ in range(4)]

    def step(i):
        lat, noise, t, ehs = batches[i % len(batches)]
        loss, _, _, _ = P.finetune_step(s_ac, t_ac, sched, lat, noise, t, ehs, fs, ft)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "LIBRARY reference point, not the product: oracle restatement on the GPU via stock torch ops "
                              "(cuDNN / cuBLAS / SDPA / fused torch AdamW), bf16 autocast, eager",
                      "metric": "train samples/sec @512px pruned U-Net (DDPM+KD)", "value": B / (ms * 1e-3), "unit": "samples/s",
                      "ms_per_step": ms, "batch": B, "ratio": args.ratio, "channels_last": args.channels_last,
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "last_loss": float(loss),
                      "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
