"""Iterations/s of the reference's own training loop shape (Tutorial3: 64x64 rays, 64 frames, batchsize 6, 'full' image
loss; Fit_Synthetic_LP_Flares: cfg5-like 64x64x100, Q/U lightcurves, batchsize 6) through the reference-facing API:
TemporalBatchedArgs.sample -> TrainStep.__call__ -> gradient_step_image (render, loss, gradient, Adam) per iteration,
host targets indexed and shipped every step as the reference does.  Usage (GPU box): python scripts/train_loop_bench.py [iters]"""
import json
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import network, optimization, synthetic  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, batch in (('cfg1_tutorial3', 6), ('cfg5_alma', 6), ('cfg2_lp_flare', 8)):
    c = synthetic.make_config(name)
    rt, pr, kind = c['rt'], c['predictor'], c['cfg']['loss']
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    rta = OrderedDict((k, rt[k]) for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos', 't_injection'))
    nt, S, A, B = len(c['t_frames']), c['S'], c['A'], c['B']
    shape = (nt, S) if kind == 'lc' else ((nt, A, B) if S == 1 else (nt, S, A, B))
    ts = optimization.TrainStep.image(c['t_frames'], c['target'].reshape(shape), sigma=c['sigma'].reshape(shape), dtype=kind)
    state = pred.init_state(pred.init_params(seed=1), num_iters=iters * 2, lr_init=1e-4, lr_final=1e-6)
    np.random.seed(0)
    for _ in range(10):
        loss, state, _ = ts(state, rta, ts.args[0].sample(batch))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        loss, state, images = ts(state, rta, ts.args[0].sample(batch))
    l = float(loss.item())
    dt = time.perf_counter() - t0
    scene = network._scene_for(pred, *rta.values(), 'hr', device=state.flat.device)
    print(json.dumps(dict(config=name, batchsize=batch, iters=iters, it_per_s=iters / dt, ms_per_it=1e3 * dt / iters,
                          dense_samples_per_s=batch * c['P'] * c['G'] * iters / dt,
                          evaluated_samples_per_it=batch * scene.n_active, loss=l)), flush=True)
