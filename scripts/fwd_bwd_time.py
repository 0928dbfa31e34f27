"""CUDA-event times of the tcgen05 forward (with saved activations) and backward at a given frame count, production build
(scripts/tc_timing_bwd.py measures the instrumented build).  Usage: python scripts/fwd_bwd_time.py [frames ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

for frames in [int(a) for a in sys.argv[1:]] or [25, 50]:
    c = synthetic.make_config('cfg2_lp_flare', nt=frames)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best = [1e9, 1e9]
    for it in range(5):
        ev[0].record()
        images, e, acts = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
        ev[1].record()
        _, dI = engine.loss_image(images, c['target'], c['sigma'], c['offset'], 1.0, c['cfg']['loss'])
        torch.cuda.synchronize()
        ev[1].record()
        g = engine.render_bwd(scene, params, tf, dI, e, acts, 'tc', max_workspace=60 * 2 ** 30)
        ev[2].record()
        torch.cuda.synchronize()
        if it:
            best = [min(best[0], ev[0].elapsed_time(ev[1])), min(best[1], ev[1].elapsed_time(ev[2]))]
    print('frames=%d  fwd %.3f ms  bwd (dout + fused) %.3f ms   per 25 frames: %.3f / %.3f' % (
        frames, best[0], best[1], best[0] * 25 / frames, best[1] * 25 / frames), flush=True)
    del acts, e, images
    engine._workspaces.clear(); torch.cuda.empty_cache()
