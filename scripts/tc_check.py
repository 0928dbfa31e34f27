"""GPU bring-up check of the tcgen05 family against the committed goldens and the fp32 SIMT family.
Usage (on the B200 box): python scripts/tc_check.py [fwd|train]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bhnerf_b200 import constants, engine, synthetic, testing  # noqa: E402


def status():
    ws = engine._workspaces.get(torch.cuda.current_device())
    return None if ws is None else ws[:16].view(torch.int32).tolist()


def fwd_golden():
    for case in ('case_image_full', 'case_lc_IQU'):
        scene, d = testing.load_golden_scene(case)
        params = torch.as_tensor(d['params_flat'], device='cuda')
        tf = torch.as_tensor(d['t_frames'].astype(np.float32), device='cuda')
        out = {}
        for impl in ('simt', 'tc'):
            img, e, _ = engine.render_fwd(scene, params, tf, impl)
            torch.cuda.synchronize()
            ref = d['images'] if d['images'].ndim == 4 else d['images'][:, None]
            im = img.cpu().numpy().reshape(ref.shape)
            out[impl] = (im, e)
            print(case, impl, 'img rel err vs f64 oracle %.3e' % testing.rel_err(im, ref), 'status', status(), flush=True)
        print(case, 'tc vs simt e max abs diff %.3e (max e %.3e)' % ((out['tc'][1] - out['simt'][1]).abs().max().item(),
                                                                     out['simt'][1].abs().max().item()), flush=True)


def fwd_big(name='cfg1_tutorial3', nt=8, reps=3):
    c = synthetic.make_config(name, nt=nt)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    res = {}
    for impl in ('simt', 'tc'):
        for save in (False, True):
            img, e, acts = engine.render_fwd(scene, params, tf, impl, save_acts=save)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                img, e, acts = engine.render_fwd(scene, params, tf, impl, save_acts=save)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            n = nt * scene.n_active
            print('%s %s save=%d: %.3f ms, %.3e eval samples/s, %.1f TFLOP/s algorithmic fwd, status %s' % (
                name, impl, save, dt * 1e3, n / dt, n * 109312 / dt / 1e12, status()), flush=True)
            res[impl] = img
    print(name, 'tc vs simt images rel err %.3e' % ((res['tc'] - res['simt']).abs().max() / res['simt'].abs().max()).item(),
          flush=True)


def train_golden():
    for case in ('case_image_full', 'case_lc_QU', 'case_lc_IQU', 'case_vis'):
        for impl in ('simt', 'tc'):
            r = testing.run_golden_case(case, impl=impl)
            per = r.pop('per_layer')
            print(case, impl, {k: ('%.3e' % x if isinstance(x, float) else x) for k, x in r.items()}, 'status', status(),
                  flush=True)
            print('    ', ' '.join('%s %.1e' % kv for kv in per.items()), flush=True)


def train_big(name='cfg1_tutorial3', nt=8, reps=3):
    from bhnerf_b200 import _lib
    import ctypes
    c = synthetic.make_config(name, nt=nt)
    rt, pr = c['rt'], c['predictor']
    params = torch.as_tensor(synthetic.trained_like_flat_params(7)).cuda()
    scene = engine.PackedScene(rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'], rt['t_geos'],
                               rt['t_start_obs'], rt['t_injection'], pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'],
                               constants.GM_c3(t_units='hr'))
    tf = torch.as_tensor(c['t_frames']).cuda()
    off = np.zeros_like(c['target'])
    kind = c['cfg']['loss']
    tgt, sig = [torch.as_tensor(a).cuda() for a in (c['target'], c['sigma'])]
    offd = torch.as_tensor(off).cuda()
    res = {}
    lib = _lib.load()
    for impl in ('simt', 'tc'):
        out = engine.train_step_image(scene, params, tf, tgt, sig, offd, 1.0, kind, impl)
        torch.cuda.synchronize()
        lib.bhnerf_profile_begin()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = engine.train_step_image(scene, params, tf, tgt, sig, offd, 1.0, kind, impl)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        ms = (ctypes.c_double * 7)(); sc = (ctypes.c_int64 * 7)(); ln = (ctypes.c_int64 * 7)()
        lib.bhnerf_profile_end(ms, sc, ln)
        n = nt * scene.n_active
        print('%s %s train: %.3f ms, %.3e eval samples/s, %.1f TFLOP/s algorithmic; per-step ms fwd %.3f bwd %.3f wgrad %.3f '
              'heads %.3f misc %.3f status %s' % (name, impl, dt * 1e3, n / dt, n * 317184 / dt / 1e12, ms[0] / reps,
                                                  ms[1] / reps, ms[2] / reps, ms[3] / reps, ms[4] / reps, status()), flush=True)
        res[impl] = [o.clone() for o in out]
    ls, is_, gs = res['simt']; lt, it_, gt = res['tc']
    print(name, 'tc vs simt: loss %.3e images %.3e grads %.3e (l2 %.3e)' % (
        abs(lt.item() - ls.item()) / abs(ls.item()), ((it_ - is_).abs().max() / is_.abs().max()).item(),
        ((gt - gs).abs().max() / gs.abs().max()).item(), ((gt - gs).norm() / gs.norm()).item()), flush=True)
    # per-layer gradient errors
    o = 0
    for i, (fi, fo) in enumerate([(21, 128), (128, 128), (128, 128), (149, 128), (128, 1)]):
        for nm, sz in (('W', fi * fo), ('b', fo)):
            a, bb = gt[o:o + sz], gs[o:o + sz]
            print('   %s%d max-rel %.3e (ref max %.3e)' % (nm, i, ((a - bb).abs().max() / gs.abs().max()).item(), bb.abs().max().item()))
            o += sz


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
    if what == 'fwd':
        fwd_golden()
        fwd_big('cfg1_tutorial3', 8)
        fwd_big('cfg2_lp_flare', 16)
    else:
        train_golden()
        train_big('cfg1_tutorial3', 8)
        train_big('cfg2_lp_flare', 16)
