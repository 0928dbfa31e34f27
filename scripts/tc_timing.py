"""Cycle accounting of the tcgen05 forward kernel (instrumented build, -DBH_TC_TIMING): where block 0's MMA-issuer
thread and one epilogue thread spend their cycles.  Usage (GPU box):
   python scripts/tc_timing.py            # builds lib/libbhnerf_b200_timing.so and runs cfg2 x 16 frames"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

flags = [a for a in sys.argv[1:] if a.startswith('-D')]
tag = ''.join(c if c.isalnum() else '_' for c in ''.join(flags))
out = os.path.join(ge.LIBDIR, 'libbhnerf_b200_timing%s.so' % tag)
srcs = [os.path.join(ge.CSRC, f) for f in ge.LIB_SOURCES]
hdrs = [os.path.join(ge.CSRC, f) for f in os.listdir(ge.CSRC) if f.endswith('.cuh')]
if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs + hdrs):
    subprocess.check_call(['/usr/local/cuda/bin/nvcc'] + ge.NVCC_FLAGS + ['-DBH_TC_TIMING'] + flags + ['-o', out] + srcs)
if 'build' in sys.argv:
    sys.exit(0)
print('variant:', tag or 'base')
os.environ['BHNERF_B200_LIB'] = out
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bhnerf_b200 import constants, engine, synthetic  # noqa: E402

c = synthetic.make_config('cfg2_lp_flare', nt=16)
rt, pr = c['rt'], c['predictor']
params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                           rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                           constants.GM_c3(t_units='hr'))
tf = torch.as_tensor(c['t_frames']).cuda()
for save in (False, True):
    for _ in range(2):
        engine.render_fwd(scene, params, tf, 'tc', save_acts=save)
    torch.cuda.synchronize()
    ws = engine._workspaces[torch.cuda.current_device()]
    w = ws[:256].view(torch.int32).cpu().numpy().astype(np.int64)
    val = lambda i: int((w[i] & 0xffffffff) | (w[i + 1] << 32))
    names = [('mma: wait weights', 8), ('mma: wait A', 10), ('mma: issue+commit', 12), ('epi: features', 14),
             ('epi: wait D', 16), ('epi: epilogue', 18), ('epi: wait staging buffer (inside epilogue)', 36)]
    rounds = (16 * scene.n_pad // 128 + 2 * 148 - 1) // (2 * 148)
    print('save=%d rounds/CTA=%d' % (save, rounds))
    for n, i in names:
        print('   %-44s %12d cycles  (%8.0f per round)' % (n, val(i), val(i) / rounds))
    rec = ws[256 + 710 * 4:256 + 858 * 4].view(torch.int32).cpu().numpy().astype(np.int64) & 0xffffffff
    order = np.argsort(rec & 0x3ff)
    print('   per CTA (sorted by SM id): k-cycles per round')
    line = ''
    for i in order:
        line += ' %3d:%5.2f' % (rec[i] & 0x3ff, (rec[i] >> 10) * 1.024 / rounds)
        if len(line) > 150:
            print('    ' + line); line = ''
    print('    ' + line)
