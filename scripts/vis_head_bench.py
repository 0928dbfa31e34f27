"""Visibility head (bhnerf_vis_head) at the cfg3 shape: time per step for different frame-group sizes (L2 reuse of A in the
backward pass vs fewer launches), against the HBM roofline of its algorithmic bytes (two passes over A).
   python scripts/vis_head_bench.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bhnerf_b200 import engine  # noqa: E402

Bt, V, P = 64, 190, 16384
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.view_as_complex(torch.randn(Bt, V, P, 2, device='cuda', generator=g))
img = torch.rand(Bt, 1, P, device='cuda', generator=g)
tgt = torch.view_as_complex(torch.randn(Bt, V, 2, device='cuda', generator=g))
sig = torch.full((Bt, V), 0.01, device='cuda')
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6545.3) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6545.3
alg = 2 * 8.0 * V * P * Bt
for gb in (24 << 20, 48 << 20, 100 << 20, 400 << 20, 1 << 40):
    for _ in range(3):
        engine.vis_head(A, img, tgt, sig, 1.0, 'vis', group_bytes=gb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        loss, vis, dI = engine.vis_head(A, img, tgt, sig, 1.0, 'vis', group_bytes=gb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('group %8.0f MB of A: %.3f ms per step, %.0f GB/s of algorithmic bytes = %.2f of HBM peak (%.0f)' % (
        min(gb, A.numel() * 8) / 2 ** 20, ms, alg / ms / 1e6, alg / ms / 1e6 / peak, peak), flush=True)
# forward and backward alone (whole batch)
for name, fn in (('vis_fwd', lambda: engine.vis_fwd(A, img)), ('vis_bwd', lambda: engine.vis_bwd(A, tgt, P))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%s alone: %.3f ms, %.0f GB/s = %.2f of HBM peak' % (name, ms, alg / 2 / ms / 1e6, alg / 2 / ms / 1e6 / peak), flush=True)

# ---- separable tensor-core DFT head (opt-in: the caller hands over (u, v) instead of the matrix) at the same shape ----
NA = NB = 128
uv = (torch.rand(Bt, V, 2, device='cuda', generator=g) - 0.5) * 0.4
grid = (-20.0, 40.0 / NA, -20.0, 40.0 / NB)
imgs = img.view(Bt, NA, NB)
dvis = torch.view_as_complex(torch.randn(Bt, V, 2, device='cuda', generator=g))
for name, fn in (('vis_dft_fwd', lambda: engine.vis_dft_fwd(uv, imgs, grid)),
                 ('vis_dft_bwd', lambda: engine.vis_dft_bwd(uv, dvis, grid, NA, NB))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flop = 2.0 * 2 * Bt * V * NA * NB            # one complex-by-real GEMM: cos and sin parts
    print('%s: %.3f ms for %d frames (explicit-matrix pass: %.3f ms of HBM time at peak), %.1f TFLOP/s algorithmic'
          % (name, ms, Bt, 8.0 * V * P * Bt / peak / 1e6, flop / ms / 1e9), flush=True)
