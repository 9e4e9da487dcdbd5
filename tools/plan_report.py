"""Offline (no GPU) report of the GEMM tile plans for the U-Net's main shape families at the bench workload (B = 16,
64x64 latents): block_n, split-K, pair / tall tiles, stages, waves over the CTA slots and the slot utilisation of the
last wave.  Uses the host-only planner query b200pdm_gemm_plan.     python tools/plan_report.py [batch]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unlearn_ft_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rows = []
for hw, C in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    M = B * hw
    tm = math.ceil(M / 128)
    fam = [
        ("conv3x3 C->C fprop", C, 9 * math.ceil(C / 64), False, False),
        ("conv3x3 2C->C fprop (up)", C, 9 * math.ceil(2 * C / 64), False, False),
        ("conv3x3 dgrad (weights MN-major)", C, 9 * math.ceil(C / 64), True, False),
        ("linear C->C", C, math.ceil(C / 64), False, False),
        ("linear C->3C (qkv)", 3 * C, math.ceil(C / 64), False, False),
        ("linear C->8C (GEGLU proj)", 8 * C, math.ceil(C / 64), False, False),
        ("linear 4C->C (FF out)", C, math.ceil(4 * C / 64), False, False),
    ]
    for name, n, kb, b_mn, _ in fam:
        p = _lib.gemm_plan(n, tiles_m=tm, kblocks=kb, b_mn=b_mn, can_split=(M * n * 4 <= 64 << 20), split_needs_finalize=True)
        waves = p["tiles"] / p["slots"]
        rows.append((f"{hw:>5} px  C={C:<5} {name:<34}", M, n, kb * 64, p, waves))
    # conv wgrad: M = C_out, N = C_in per tap (9 groups), K = pixels
    p = _lib.gemm_plan(C, n_groups=9, b_mn=True, tiles_m=math.ceil(C / 128), kblocks=math.ceil(M / 64), can_split=True,
                       split_needs_finalize=False)
    rows.append((f"{hw:>5} px  C={C:<5} {'conv3x3 wgrad (K = pixels)':<34}", C, C, M, p, p["tiles"] / p["slots"]))
print(f"{'shape':<58} {'M':>7} {'N':>6} {'K':>7}  bn  split pair msub stg   tiles/slots   last-wave fill")
for name, M, n, K, p, waves in rows:
    full = math.ceil(waves)
    fill = waves / full
    print(f"{name:<58} {M:>7} {n:>6} {K:>7} {p['block_n']:>4} {p['splits']:>5} {p['pair']:>4} {p['m_sub']:>4} {p['stages']:>3} "
          f"{p['tiles']:>6}/{p['slots']:<4} = {waves:5.2f} waves   {fill:4.2f}")
