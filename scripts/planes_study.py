"""Gradient error of the one-plane (bf16 hi only) backward plan on the small golden cases (5.5e3 sample-frames) and on
cfg1 with few frames, against the float64 oracle / the fp32 SIMT family.  Run with BHNERF_TC_PLANES=1."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import constants, engine, synthetic, testing  # noqa: E402

print('BHNERF_TC_PLANES =', os.environ.get('BHNERF_TC_PLANES'))
for case in ('case_image_full', 'case_lc_QU', 'case_lc_IQU', 'case_vis', 'case_amp', 'case_cphase'):
    r = testing.run_golden_case(case, impl='tc')
    print('%-16s n_active*Bt=%6d  img %.2e  grad max-rel %.2e  l2 %.2e' % (case, r['n_active'] * 4, r['img_err'], r['grad_err'],
                                                                        r['grad_err_l2']))
c = synthetic.make_config('cfg1_tutorial3', nt=8)
rt, pr = c['rt'], c['predictor']
params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                           rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                           constants.GM_c3(t_units='hr'))
for nb in (1, 2, 6):
    tf = torch.as_tensor(c['t_frames'][:nb]).cuda()
    tg, sg = c['target'][:nb], c['sigma'][:nb]
    off = np.zeros_like(tg)
    _, _, gs = engine.train_step_image(scene, params, tf, tg, sg, off, 1.0, 'full', 'simt')
    gs = gs.double().clone()
    _, _, gt = engine.train_step_image(scene, params, tf, tg, sg, off, 1.0, 'full', 'tc')
    gt = gt.double()
    print('cfg1 x %d frames (%7d sample-frames): grad max-rel %.2e  l2 %.2e' % (
        nb, scene.n_active * nb, ((gt - gs).abs().max() / gs.abs().max()).item(), ((gt - gs).norm() / gs.norm()).item()))
