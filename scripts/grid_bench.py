"""Throughput of the voxel-grid renderers (csrc/grid.cu) at BASELINE.json shapes: emission.image_plane_dynamics (mode 0,
64^3 grid) and the GRID_Predictor forward + pull-back (mode 1).  Per evaluated sample-frame the kernel reads the packed
sample (24 + 4 S bytes, L2-resident across frames) and 8 grid corners (32 B, L2/L1): it is bound by L2 gathers, not HBM.
Usage (GPU box): python scripts/grid_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

for name in ('cfg1_tutorial3', 'cfg2_lp_flare'):
    c = synthetic.make_config(name)
    rt, pr = c['rt'], c['predictor']
    fov = 2.0 * pr['rmax']
    res = 64
    ax = np.linspace(-fov / 2, fov / 2, res, dtype=np.float32)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    grid = torch.as_tensor(np.exp(-((X - 0.5 * pr['rmax']) ** 2 + Y ** 2 + Z ** 2) / 8.0).astype(np.float32)).cuda()
    tf = torch.as_tensor(c['t_frames']).cuda()
    Bt = tf.numel()
    for mode, label in ((0, 'image_plane_dynamics'), (1, 'GRID_Predictor')):
        rmax = 0.5 * fov * np.sqrt(3.0) * 1.00001 if mode == 0 else pr['rmax']
        scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                                   rt['t_start_obs'], rt['t_injection'], pr['scale'], 0.0 if mode == 0 else pr['rmin'], rmax,
                                   0.5 * fov * 1.00001 if mode == 0 else pr['z_width'], constants.GM_c3(t_units='hr'))
        g = grid if mode == 0 else (grid * 4.0 + 8.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for _ in range(3):
            images, _ = engine.grid_render_fwd(scene, g, fov, tf, mode)
        dI = torch.ones_like(images)
        if mode == 1:
            engine.grid_render_bwd(scene, g, fov, tf, dI)
        torch.cuda.synchronize()
        n = 5
        ev[0].record()
        for _ in range(n):
            images, _ = engine.grid_render_fwd(scene, g, fov, tf, mode)
        ev[1].record()
        if mode == 1:
            for _ in range(n):
                engine.grid_render_bwd(scene, g, fov, tf, dI)
        ev[2].record()
        torch.cuda.synchronize()
        fwd = ev[0].elapsed_time(ev[1]) / n
        sf = Bt * scene.n_active
        r = dict(config=name, renderer=label, frames=Bt, grid=res, active_fraction=scene.n_active / (c['P'] * c['G']),
                 fwd_ms=fwd, evaluated_sample_frames_per_s=sf / fwd * 1e3, dense_samples_per_s=Bt * c['P'] * c['G'] / fwd * 1e3,
                 gather_GBps=sf * (32 + 24 + 4 * scene.S) / fwd * 1e3 / 1e9)
        if mode == 1:
            bwd = ev[1].elapsed_time(ev[2]) / n
            r.update(bwd_ms=bwd, bwd_evaluated_sample_frames_per_s=sf / bwd * 1e3)
        print(json.dumps(r), flush=True)
