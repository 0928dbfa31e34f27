"""Gradient of the tcgen05 family against the fp32 SIMT family at a bench-like size (cfg2, `frames` frames, one-plane
plan), for A/B runs of the backward's precision plans:
   python scripts/grad_ab.py [frames]                       (default plan)
   BHNERF_TC_DGRAD=bf16 python scripts/grad_ab.py [frames]  (two-product bf16 dgrad chain)
Prints max-relative (the test metric) and norm-relative errors, whole vector and per layer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for name, nt in (('cfg2_lp_flare', frames), ('cfg1_tutorial3', 8)):
    c = synthetic.make_config(name, nt=nt)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    kind = c['cfg']['loss']
    off = np.zeros_like(c['target'])
    ls, is_, gs = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], off, 1.0, kind, 'simt')
    gs = gs.double().clone()
    lt, it_, gt = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], off, 1.0, kind, 'tc')
    gt = gt.double()
    print('%s x%d frames (%d sample-frames)  dgrad plan: %s  status %s' % (
        name, nt, scene.n_active * nt, os.environ.get('BHNERF_TC_DGRAD', 'default (bf16 two-product chain)'),
        engine.workspace_status()[:5]))
    print('  whole gradient: max-rel %.3e   norm-rel %.3e' % (
        ((gt - gs).abs().max() / gs.abs().max()).item(), ((gt - gs).norm() / gs.norm()).item()))
    o = 0
    for nm, n in (('W0', 21 * 128), ('b0', 128), ('W1', 128 * 128), ('b1', 128), ('W2', 128 * 128), ('b2', 128),
                  ('W3', 149 * 128), ('b3', 128), ('W4', 128), ('b4', 1)):
        a, b = gt[o:o + n], gs[o:o + n]
        print('    %-3s max-rel %.3e  norm-rel %.3e' % (nm, ((a - b).abs().max() / b.abs().max()).item(),
                                                       ((a - b).norm() / b.norm()).item()))
        o += n
