"""Error attribution of the tcgen05 backward on a golden case: which input (e, d_images, saved activations)
carries the gradient error.  Usage: python scripts/tc_bwd_isolate.py [case]"""
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
img_s, e_s, acts_s = engine.render_fwd(scene, params, tf, 'simt', save_acts=True)
img_t, e_t, acts_t = engine.render_fwd(scene, params, tf, 'tc', save_acts=True)
_, dI_s = engine.loss_image(img_s, tgt, sig, off, 1.0, kind)
_, dI_t = engine.loss_image(img_t, tgt, sig, off, 1.0, kind)
ref = d['grads']
def err(g): return testing.rel_err(g.cpu().numpy(), ref)
print('dI rel diff tc vs simt: %.3e' % ((dI_t - dI_s).abs().max() / dI_s.abs().max()).item())
print('e  rel diff tc vs simt: %.3e' % ((e_t - e_s).abs().max() / e_s.abs().max()).item())
print('simt bwd (simt e, simt dI)        : %.3e' % err(engine.render_bwd(scene, params, tf, dI_s, e_s, acts_s, 'simt')))
print('simt bwd (simt e, TC dI)          : %.3e' % err(engine.render_bwd(scene, params, tf, dI_t, e_s, acts_s, 'simt')))
print('simt bwd (TC e, simt dI)          : %.3e' % err(engine.render_bwd(scene, params, tf, dI_s, e_t, acts_s, 'simt')))
print('tc bwd   (simt e, simt dI, tc acts): %.3e' % err(engine.render_bwd(scene, params, tf, dI_s, e_s, acts_t, 'tc')))
print('tc bwd   (tc e, simt dI, tc acts)  : %.3e' % err(engine.render_bwd(scene, params, tf, dI_s, e_t, acts_t, 'tc')))
print('tc bwd   (tc e, tc dI, tc acts)    : %.3e' % err(engine.render_bwd(scene, params, tf, dI_t, e_t, acts_t, 'tc')))
