"""Per-hop timeline of the fused backward's dgrad CTA (instrumented build, -DBH_TC_TRACE): CTA 0 logs the clock at every hop of
rounds 10 and 11 -- MMA warp: operand seen / products issued + committed; epilogue warp 0: accumulator seen / in registers /
operand stored / handed over; tops; hand-over to the wgrad CTA.  Prints the events of the two rounds in time order.
   python scripts/tc_trace_bwd.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

out = os.path.join(ge.LIBDIR, 'libbhnerf_b200_trace.so')
srcs = [os.path.join(ge.CSRC, f) for f in ge.LIB_SOURCES]
hdrs = [os.path.join(ge.CSRC, f) for f in os.listdir(ge.CSRC) if f.endswith('.cuh')]
if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs + hdrs):
    subprocess.check_call(['/usr/local/cuda/bin/nvcc'] + ge.NVCC_FLAGS + ['-DBH_TC_TRACE', '-o', out] + srcs)
if 'build' in sys.argv:
    sys.exit(0)
os.environ['BHNERF_B200_LIB'] = out
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

c = synthetic.make_config('cfg2_lp_flare', nt=25)
rt, pr = c['rt'], c['predictor']
params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                           rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                           constants.GM_c3(t_units='hr'))
tf = torch.as_tensor(c['t_frames']).cuda()
for it in range(3):
    images, e, acts = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
    _, dI = engine.loss_image(images, c['target'], c['sigma'], c['offset'], 1.0, c['cfg']['loss'])
    g = engine.render_bwd(scene, params, tf, dI, e, acts, 'tc', max_workspace=40 * 2 ** 30)
    torch.cuda.synchronize()
ws = engine._workspaces[torch.cuda.current_device()]
w = ws[256 + 710 * 4:256 + (710 + 92) * 4].view(torch.int32).cpu().numpy().astype(np.int64) & 0xffffffff
ev = []
for r in range(2):
    for li, l in enumerate((3, 2, 1)):
        for s in range(2):
            k = ((r * 6 + li * 2 + s) * 2)
            ev.append((w[k], 'MMA  r%d l%d s%d  operand seen' % (10 + r, l, s)))
            ev.append((w[k + 1], 'MMA  r%d l%d s%d  products issued + committed' % (10 + r, l, s)))
    base = 24 + r * 30
    for s in range(2):
        ev.append((w[base + s * 2], 'EPI  r%d top s%d  start' % (10 + r, s)))
        ev.append((w[base + s * 2 + 1], 'EPI  r%d top s%d  handed to the MMA warp' % (10 + r, s)))
    for li, l in enumerate((3, 2, 1)):
        for s in range(2):
            k = base + 4 + (li * 2 + s) * 4
            ev.append((w[k], 'EPI  r%d l%d s%d  accumulator seen' % (10 + r, l, s)))
            ev.append((w[k + 1], 'EPI  r%d l%d s%d  accumulator in registers' % (10 + r, l, s)))
            if l > 1:
                ev.append((w[k + 2], 'EPI  r%d l%d s%d  packed, stored, operand in TMEM' % (10 + r, l, s)))
            ev.append((w[k + 3], 'EPI  r%d l%d s%d  %s' % (10 + r, l, s, 'handed to the MMA warp' if l > 1 else 'delta_0 stored')))
    ev.append((w[84 + r * 4], 'EPI  r%d loop top' % (10 + r)))
    ev.append((w[84 + r * 4 + 1], 'EPI  r%d next round\'s inputs requested' % (10 + r)))
    ev.append((w[84 + r * 4 + 2], 'EPI  r%d top s0: before the ring-set wait' % (10 + r)))
    ev.append((w[base + 28], 'EPI  r%d hand-over to the wgrad CTA: start' % (10 + r)))
    ev.append((w[base + 29], 'EPI  r%d hand-over to the wgrad CTA: done' % (10 + r)))
ev = [(t, n) for t, n in ev if t != 0]
t0 = min(t for t, _ in ev)
ev.sort(key=lambda x: (x[0] - t0) & 0xffffffff)
prev = 0
print('cycles since the first event | since the previous event | event   (CTA 0 = dgrad CTA of pair 0, cfg2 x 25 frames)')
for t, n in ev:
    dt = (t - t0) & 0xffffffff
    print('%8d  %6d  %s' % (dt, dt - prev, n))
    prev = dt
