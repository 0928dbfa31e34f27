"""Where a SMALL train step (Tutorial3 shape, batchsize 6) spends its device time: per-category CUDA-event times of the
library (direct calls, no graph) next to the wall time per iteration.  Usage: python scripts/small_step_profile.py [planes]"""
import ctypes
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import _lib, engine, network, synthetic  # noqa: E402

c = synthetic.make_config('cfg1_tutorial3')
rt, pr = c['rt'], c['predictor']
pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
rta = OrderedDict((k, rt[k]) for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs', 't_geos', 't_injection'))
dev = torch.device('cuda')
scene = network._scene_for(pred, *rta.values(), 'hr', device=dev)
params = torch.as_tensor(synthetic.trained_like_flat_params(7), device=dev)
B = 6
tf = torch.as_tensor(c['t_frames'][:B], device=dev)
tgt, sig, off = [torch.as_tensor(c[k][:B], device=dev) for k in ('target', 'sigma', 'offset')]
lib = _lib.load()
out = None
for it in range(20):
    out = engine.train_step_image(scene, params, tf, tgt, sig, off, 1.0, 'full', out=out)
torch.cuda.synchronize()
lib.bhnerf_profile_begin()
t0 = time.perf_counter()
N = 200
for it in range(N):
    engine.train_step_image(scene, params, tf, tgt, sig, off, 1.0, 'full', out=out)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = (ctypes.c_double * 7)(); sc = (ctypes.c_int64 * 7)(); ln = (ctypes.c_int64 * 7)()
lib.bhnerf_profile_end(ms, sc, ln)
names = ['render_fwd', 'render_bwd', 'wgrad', 'heads', 'misc']
print('planes env=%s  n_active=%d  sample-frames=%d  wall %.3f ms/step' % (os.environ.get('BHNERF_TC_PLANES'), scene.n_active,
                                                                           scene.n_active * B, 1e3 * wall / N))
for n, m, l in zip(names, ms, ln):
    print('   %-11s %.4f ms/step  (%d launches/step)' % (n, m / N, l // N))
