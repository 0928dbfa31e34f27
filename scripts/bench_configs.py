"""One train step of every BASELINE.json config on ONE GPU (the per-GPU share where the config names 8 GPUs):
device-timed ms/step, dense and evaluated samples/s.  Usage (GPU box): python scripts/bench_configs.py [steps]
cfg1/2/5: image / lightcurve losses through the fused C-ABI step; cfg3: visibility loss through
network.gradient_step_eht (render fwd -> A_t vec(I_t) -> chi^2 -> adjoint -> render bwd, Adam included);
cfg4: 16 of its 128 frames (the 8-GPU frame shard)."""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import engine, network, synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda', 0)
rows = []
for name, nt in (('cfg1_tutorial3', None), ('cfg2_lp_flare', None), ('cfg3_ngeht', None), ('cfg4_highres', 16),
                 ('cfg5_alma', None)):
    c = synthetic.make_config(name, nt=nt)
    rt, pr, kind = c['rt'], c['predictor'], c['cfg']['loss']
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    rta = OrderedDict((k, rt[k]) for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos', 't_injection'))
    scene = network._scene_for(pred, *rta.values(), 'hr', device=dev)
    state = pred.init_state(network.unflatten_params(synthetic.trained_like_flat_params(7)), num_iters=10000, device=dev)
    Bt = len(c['t_frames'])
    tf = torch.as_tensor(c['t_frames'], device=dev)
    if kind == 'vis':
        A = torch.as_tensor(c['Amat'], device=dev)
        tgt = torch.as_tensor(c['target'], device=dev); sig = torch.as_tensor(c['sigma'], device=dev)
        step = lambda: network.gradient_step_eht(state, 'hr', 'vis', tgt, sig, A, tf, *rta.values(), 1.0)[0]
    else:
        tgt, sig, off = [torch.as_tensor(c[k], device=dev) for k in ('target', 'sigma', 'offset')]
        step = lambda: engine.train_step_image(scene, state.flat, tf, tgt, sig, off, 1.0, kind, max_workspace=40 * 2 ** 30)[0]
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    dense = Bt * c['P'] * c['G']
    r = dict(config=name, frames=Bt, rays=c['P'], samples_per_ray=c['G'], stokes=c['S'], loss=kind,
             active_fraction=scene.n_active / (c['P'] * c['G']), ms_per_step=ms, dense_samples_per_s=dense / ms * 1e3,
             evaluated_samples_per_s=Bt * scene.n_active / ms * 1e3,
             algorithmic_tflops=Bt * scene.n_active * 317184 / ms * 1e3 / 1e12, loss_value=float(loss.item()),
             status=engine.workspace_status(dev)[:5])
    if kind == 'vis':
        r['A_bytes_per_step'] = int(2 * A.numel() * 8)      # streamed once forward, once for the adjoint
        r['V'] = int(A.shape[1])
    rows.append(r)
    print(json.dumps(r), flush=True)
    del scene, state
    engine._workspaces.clear(); network._scene_cache.clear(); torch.cuda.empty_cache()
