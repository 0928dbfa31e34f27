"""Compare the tcgen05 dgrad kernel's cotangent images (hi [+ lo] bf16 planes) with the fp32 SIMT chain's."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import engine, testing

case = sys.argv[1] if len(sys.argv) > 1 else 'case_lc_QU'
scene, d = testing.load_golden_scene(case)
params = torch.as_tensor(d['params_flat'], device='cuda')
tf = torch.as_tensor(d['t_frames'].astype(np.float32), device='cuda')
kind = 'full' if case == 'case_image_full' else 'lc'
tgt = d['target']; sig = d['sigma'] if 'sigma' in d.files else np.ones_like(tgt); off = np.zeros_like(tgt)
Bt = tf.numel(); npad = scene.n_pad
img_s, e_s, acts_s = engine.render_fwd(scene, params, tf, 'simt', save_acts=True)
img_t, e_t, acts_t = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
_, dI = engine.loss_image(img_s, tgt, sig, off, 1.0, kind)
g_s = engine.render_bwd(scene, params, tf, dI, e_s, acts_s, 'simt')
torch.cuda.synchronize()
ws = engine._workspaces[torch.cuda.current_device()]
fixed = 196608
ds = ws[fixed: fixed + Bt * 512 * npad * 4].view(torch.float32).reshape(Bt, 4, 128, npad).clone()
g_t = engine.render_bwd(scene, params, tf, dI, e_s, acts_t, 'tc')
torch.cuda.synchronize()
ws = engine._workspaces[torch.cuda.current_device()]
PL = 2 if Bt * scene.n_active < 2 ** 19 else 1
PL = int(os.environ.get('BHNERF_TC_PLANES', PL))
fixed = 1 << 20
del_fs = npad * (PL * 1024 + 32)
raw = ws[fixed: fixed + Bt * del_fs].reshape(Bt, del_fs)
ntile = npad // 128
def images(buf, nl):   # buf: [nl*ntile*32768] bytes -> [nl, npad, 128] float
    x = buf.view(torch.bfloat16).reshape(nl, ntile, 16, 16, 8, 8)      # [l][tile][colgroup][rowgroup][r][c]
    x = x.permute(0, 1, 3, 4, 2, 5).reshape(nl, npad, 128)
    return x.float()
for b in range(Bt):
    hi = images(raw[b, : 4 * npad * 256], 4)
    tot = hi.double()
    if PL == 2:
        lo = images(raw[b, npad * 1024: npad * 1024 + 4 * npad * 256], 4)
        tot = tot + lo.double()
    ref = ds[b].permute(0, 2, 1).double()      # [l][i][j]
    for l in range(4):
        den = ref[l].abs().max().item()
        print('frame %d delta_%d: max|ref| %.3e  hi err %.3e  hi+lo err %.3e  colsum ref %.4e tc %.4e' % (
            b, l, den, (hi[l].double() - ref[l]).abs().max().item() / max(den, 1e-300),
            (tot[l] - ref[l]).abs().max().item() / max(den, 1e-300), ref[l].sum(0).abs().max().item(),
            tot[l].sum(0).abs().max().item()))
print('grad err tc vs simt %.3e' % ((g_t - g_s).abs().max() / g_s.abs().max()).item())
