"""How do the forward and the fused backward scale when they get only part of the GPU (BHNERF_TC_FWD_GRID /
BHNERF_TC_BWD_CLUSTERS)?  If a kernel on half the SMs takes less than twice as long, it is bound by a chip-wide resource
(HBM write path, L2) and would gain from sharing the chip with a kernel bound by something else.
   python scripts/partial_chip_probe.py [frames]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT)
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
    img, e, acts = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
    _, dI = engine.loss_image(img, c['target'], c['sigma'], c['offset'], 1.0, 'lc')
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):
        ev[0].record()
        engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
        ev[1].record()
        engine.render_bwd(scene, params, tf, dI, e, acts, 'tc', max_workspace=40 * 2 ** 30)
        ev[2].record()
        torch.cuda.synchronize()
    print('fwd CTAs %-4s bwd clusters %-4s: fwd %.3f ms  bwd %.3f ms' % (os.environ.get('BHNERF_TC_FWD_GRID', '148'),
          os.environ.get('BHNERF_TC_BWD_CLUSTERS', '74'), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])), flush=True)
    sys.exit(0)
frames = sys.argv[1] if len(sys.argv) > 1 else '25'
for fg, bc in ((None, None), ('111', '55'), ('74', '37'), ('37', '18')):
    env = dict(os.environ)
    if fg:
        env['BHNERF_TC_FWD_GRID'] = fg; env['BHNERF_TC_BWD_CLUSTERS'] = bc
    subprocess.run(['timeout', '60', sys.executable, os.path.abspath(__file__), 'child', frames], env=env, check=False)
