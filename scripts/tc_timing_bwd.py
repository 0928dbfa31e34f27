"""Cycle accounting of the fused tcgen05 backward (instrumented build, -DBH_TC_TIMING): where pair 0's dgrad CTA
(MMA issuer, epilogue thread 0) and wgrad CTA (producer, MMA issuer) spend their cycles, next to CUDA-event times
of the forward (with saved activations) and the backward.  Usage (GPU box):
   python scripts/tc_timing_bwd.py [frames] [extra nvcc -D flags ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 25
flags = [a for a in sys.argv[1:] if a.startswith('-D')]
tag = ''.join(c if c.isalnum() else '_' for c in ''.join(flags))
out = os.path.join(ge.LIBDIR, 'libbhnerf_b200_timing%s.so' % tag)
srcs = [os.path.join(ge.CSRC, f) for f in ge.LIB_SOURCES]
hdrs = [os.path.join(ge.CSRC, f) for f in os.listdir(ge.CSRC) if f.endswith('.cuh')]
if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs + hdrs):
    subprocess.check_call(['/usr/local/cuda/bin/nvcc'] + ge.NVCC_FLAGS + ['-DBH_TC_TIMING'] + flags + ['-o', out] + srcs)
if 'build' in sys.argv:
    sys.exit(0)
os.environ['BHNERF_B200_LIB'] = out
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

c = synthetic.make_config('cfg2_lp_flare', nt=frames)
rt, pr = c['rt'], c['predictor']
params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                           rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                           constants.GM_c3(t_units='hr'))
tf = torch.as_tensor(c['t_frames']).cuda()
kind = c['cfg']['loss']
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    ev[0].record()
    images, e, acts = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
    ev[1].record()
    _, dI = engine.loss_image(images, c['target'], c['sigma'], c['offset'], 1.0, kind)
    torch.cuda.synchronize()
    engine._workspaces[torch.cuda.current_device()][200:216] = 0      # the MAX-over-CTAs words
    engine._workspaces[torch.cuda.current_device()][256 + 710 * 4:256 + 860 * 4] = 0
    torch.cuda.synchronize()
    ev[1].record()
    g = engine.render_bwd(scene, params, tf, dI, e, acts, 'tc', max_workspace=40 * 2 ** 30)
    ev[2].record()
    torch.cuda.synchronize()
tiles = frames * scene.n_pad // 128
print('%s frames=%d tiles=%d  fwd %.3f ms  bwd %.3f ms  (%.0f / %.0f cycles per tile-pair per SM-pair at 1.85 GHz)' % (
    tag or 'base', frames, tiles, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
    ev[0].elapsed_time(ev[1]) * 1.85e6 / (tiles / 2 / 148), ev[1].elapsed_time(ev[2]) * 1.85e6 / (tiles / 2 / 74)))
ws = engine._workspaces[torch.cuda.current_device()]
w = ws[:256].view(torch.int32).cpu().numpy().astype(np.int64)
val = lambda i: int((w[i] & 0xffffffff) | (w[i + 1] << 32))
rounds = (tiles + 2 * 74 - 1) // (2 * 74)
names = [('dgrad mma: wait A', 20), ('dgrad mma: issue', 22), ('dgrad epi: top (delta3)', 24), ('dgrad epi: wait ring free', 26),
         ('wgrad gen: wait buffer free', 28), ('dgrad epi: wait D', 30), ('dgrad epi: layer epilogue', 32), ('dgrad epi: loop total', 34),
         ('wgrad mma: wait X (activations)', 36), ('wgrad prod: wait ring full', 38), ('wgrad prod: wait stage empty', 40),
         ('wgrad mma: wait delta_3 (rebuilt)', 42), ('wgrad mma: wait Y (cotangents)', 44), ('wgrad mma: issue', 46),
         ('wgrad mma: loop total', 48), ('dgrad epi: loop total, MAX over pairs', 50), ('kernel entry -> exit, MAX over CTAs', 52)]
print('  rounds (tile pairs) per CTA pair = %d' % rounds)
for n, i in names:
    print('   %-30s %12d cycles  (%8.0f per round)' % (n, val(i), val(i) / rounds))
print('   grad norm %.6e' % float(g.norm()))

rec = ws[256 + 710 * 4:256 + 858 * 4].view(torch.int32).cpu().numpy().astype(np.int64).reshape(74, 2)
print('  per pair: (dgrad SM, wgrad SM) rounds, cycles per round')
for p in range(74):
    cyc, info = int(rec[p, 0]) << 6, int(rec[p, 1])
    rd = (info >> 20) & 0xfff
    print('   pair %2d  SM %3d %3d  rounds %4d  %7.0f cycles/round' % (p, info & 0x3ff, (info >> 10) & 0x3ff, rd, cyc / max(rd, 1)))
