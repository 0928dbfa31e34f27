"""The STATED fast precision mode (BHNERF_PRECISION=fast: one fp16 product per layer in the forward and the dgrad chain)
against the fp32 SIMT family: its image / gradient errors are RECORDED, not asserted to meet the parity tolerances (they do
not: SURVEY.md s0.9), and its speed is printed next to the default mode's.   python scripts/fast_mode_check.py [frames]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from bhnerf_b200 import constants, engine, synthetic
    frames = int(sys.argv[2])
    c = synthetic.make_config('cfg2_lp_flare', nt=frames)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    ls, is_, gs = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], c['offset'], 1.0, 'lc', 'simt')
    is_, gs = is_.clone().double(), gs.clone().double()
    for _ in range(3):
        lt, it_, gt = engine.train_step_image(scene, params, tf, c['target'], c['sigma'], c['offset'], 1.0, 'lc', 'tc')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        engine.train_step_image(scene, params, tf, c['target'], c['sigma'], c['offset'], 1.0, 'lc', 'tc')
    e1.record(); torch.cuda.synchronize()
    it_, gt = it_.double(), gt.double()
    lc_t, lc_s = it_.sum(-1), is_.sum(-1)
    print('mode %-8s frames %d: %.3f ms/step   image err %.2e   lightcurve err %.2e   gradient err max-rel %.2e norm-rel %.2e' % (
        os.environ.get('BHNERF_PRECISION', 'default'), frames, e0.elapsed_time(e1) / 5,
        ((it_ - is_).abs().max() / is_.abs().max()).item(), ((lc_t - lc_s).abs().max() / lc_s.abs().max()).item(),
        ((gt - gs).abs().max() / gs.abs().max()).item(), ((gt - gs).norm() / gs.norm()).item()), flush=True)
    sys.exit(0)
frames = sys.argv[1] if len(sys.argv) > 1 else '25'
for mode in ('default', 'fast'):
    env = dict(os.environ)
    env.pop('BHNERF_PRECISION', None)
    if mode == 'fast':
        env['BHNERF_PRECISION'] = 'fast'
    subprocess.run([sys.executable, os.path.abspath(__file__), 'child', frames], env=env, check=True)
