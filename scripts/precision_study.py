"""Emulated-precision study that fixes the tensor-core operand plan.

Runs on CPU (torch float64 matmuls over operands rounded to bf16 / split into bf16 hi+lo planes;
fp32 accumulation error is negligible next to bf16 operand rounding).  Build container only
(uses the reference's kgeo for real geodesics when present, else the committed fixture).

forward modes  : x1 = a_hi*w_hi;  x2a = (a_hi+a_lo)*w_hi;  x3 = a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
backward modes : recompute precision / dgrad weight precision / wgrad operand precision
Reports max-norm and L2 relative errors of images and of the flat gradient vs the float64 oracle.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from oracle import bhnerf_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

torch.set_num_threads(8)


def bf(x):
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def split(x):
    hi = bf(x)
    lo = bf(x - hi)
    return hi, lo


def mm(a, w, mode):
    if mode == 'exact':
        return a @ w
    ah, al = split(a); wh, wl = split(w)
    if mode == 'x1':
        return ah @ wh
    if mode == 'x2a':
        return ah @ wh + al @ wh
    if mode == 'x2w':
        return ah @ wh + ah @ wl
    if mode == 'x3':
        return ah @ wh + al @ wh + ah @ wl
    raise ValueError(mode)


def forward(feat, W, b, mode):
    """returns pre-activations z[0..3], out"""
    zs = []
    h = feat
    for i in range(4):
        x = h if i < 3 else torch.cat([h, feat], -1)
        z = mm(x, W[i], mode) + b[i]
        zs.append(z)
        h = torch.relu(z)
    out = h @ W[4] + b[4]      # 128->1 on CUDA cores in fp32
    return zs, out[..., 0]


def backward(feat, W, zs_mask, h_acts, dout, dgrad_mode, wgrad_mode):
    """dout: (N,) d loss / d out.  h_acts = [feat, h0, h1, h2, h3] used as wgrad operands."""
    grads = [None] * 5
    h3 = h_acts[4]
    grads[4] = (h3.T @ dout[:, None], dout.sum()[None])
    d = (dout[:, None] * W[4][:, 0][None, :]) * (zs_mask[3])
    for i in (3, 2, 1, 0):
        x = h_acts[i] if i < 3 else torch.cat([h_acts[3], feat], -1)
        if i == 0:
            x = feat
        if wgrad_mode == 'exact':
            gw = x.T @ d
        elif wgrad_mode == 'x1':
            gw = bf(x).T @ bf(d)
        elif wgrad_mode == 'x2d':    # act hi only, delta hi+lo
            dh, dl = split(d); gw = bf(x).T @ dh + bf(x).T @ dl
        grads[i] = (gw, d.sum(0))
        if i > 0:
            Wi = W[i][:128]         # hidden rows only (no grad wrt posenc inputs)
            if dgrad_mode == 'exact':
                dn = d @ Wi.T
            elif dgrad_mode == 'x1':
                dn = bf(d) @ bf(Wi).T
            elif dgrad_mode == 'x2w':
                wh, wl = split(Wi); dn = bf(d) @ wh.T + bf(d) @ wl.T
            elif dgrad_mode == 'x3':
                wh, wl = split(Wi); dh, dl = split(d); dn = dh @ wh.T + dl @ wh.T + dh @ wl.T
            d = dn * zs_mask[i - 1]
    return torch.cat([torch.cat([g[0].reshape(-1), g[1].reshape(-1)]) for g in grads])


def main():
    if ref_shim.available():
        A = B = 32; G = 64
        geos = ref_shim.kerr_geodesics(0.2, np.deg2rad(60.0), 16.0, A, B, G)
        Om = ref_shim.keplerian_omega(geos); g = ref_shim.doppler_factor(geos, Om)
        coords = np.array([geos['x'], geos['y'], geos['z']]).astype(np.float32)
        inp = dict(coords=coords, Omega=Om.astype(np.float32), g=g.astype(np.float32),
                   dtau=geos['dtau'].astype(np.float32), Sigma=geos['Sigma'].astype(np.float32),
                   t_geos=geos['t'].astype(np.float32))
        rmin = float(geos['r'].min())
    else:
        d = np.load('tests/golden/kerr_a0.2_i60_16x16x32.npz'); inp = {k: d[k] for k in d.files}
        A = B = 16; G = 32; rmin = float(d['r_min'])
    rmax, zw = 8.0, 4.0
    t_frames = np.linspace(0, 1, 6)
    for pname, params in (('he_uniform', O.init_params(1)), ('trained_like', O.trained_like_params(7))):
        p = O._params_t(params, torch.float64)
        W = [p[f'Dense_{i}']['kernel'] for i in range(5)]; b = [p[f'Dense_{i}']['bias'] for i in range(5)]
        dt = torch.float64
        warped = O.velocity_warp_coords(inp['coords'], inp['Omega'], t_frames, 0.0, inp['t_geos'], -1000.0,
                                        O.GM_C3_SGRA_HR, dt, torch.float32)
        valid = torch.isfinite(warped[..., 0])
        net_in = torch.where(torch.isfinite(warped), warped, torch.zeros_like(warped))
        feat = O.posenc(net_in / rmax, 3).reshape(-1, 21)
        co = torch.as_tensor(inp['coords'], dtype=dt)
        dom = O.fill_unsupervised_emission(torch.ones_like(valid, dtype=dt), co, rmin, rmax, zw) * valid
        wgt = torch.as_tensor(inp['g'].astype(np.float64) ** 2 * inp['dtau'] * inp['Sigma'])
        wfull = (dom * wgt).reshape(-1)
        act = wfull != 0
        feat_a = feat[act]; w_a = wfull[act]
        print(f'== params={pname}: {act.sum().item()} active of {act.numel()} samples')

        def render(mode):
            zs, out = forward(feat_a, W, b, mode)
            e = torch.sigmoid(out - 10.0)
            img = torch.zeros(act.numel(), dtype=dt); img[act] = e * w_a
            return zs, e, img.reshape(len(t_frames), A, B, G).sum(-1)

        zs0, e0, img0 = render('exact')
        tgt = img0 * 0.7 + 0.01
        for mode in ('x1', 'x2a', 'x2w', 'x3'):
            _, _, img = render(mode)
            print(f'  fwd {mode:4s}: images max-rel {((img - img0).abs().max() / img0.abs().max()).item():.2e}  '
                  f'L2-rel {((img - img0).norm() / img0.norm()).item():.2e}')
        # gradients of 'full' loss, sigma=1
        dI = 2 * (img0 - tgt)
        dimg = dI[..., None].expand(len(t_frames), A, B, G).reshape(-1)[act]
        masks0 = [(z > 0).to(dt) for z in zs0]
        hs0 = [feat_a] + [torch.relu(z) for z in zs0]
        g_ref = backward(feat_a, W, masks0, hs0, dimg * w_a * e0 * (1 - e0), 'exact', 'exact')

        def rep(tag, gvec):
            err = gvec - g_ref
            # per-layer L2-rel too
            print(f'  bwd {tag:38s}: grad max-rel {(err.abs().max() / g_ref.abs().max()).item():.2e}  '
                  f'L2-rel {(err.norm() / g_ref.norm()).item():.2e}')
        for rec, dg, wg, save_e, save_mask in (
                ('x3', 'x3', 'x2d', False, False),
                ('x1', 'x1', 'x1', True, True),
                ('x1', 'x1', 'x1', True, False),
                ('x1', 'x2w', 'x1', True, True),
                ('x1', 'x2w', 'x1', True, False),
                ('x1', 'x1', 'x1', False, False),
                ('x3', 'x2w', 'x1', False, False),
                ('x3', 'x1', 'x1', False, False)):
            zs, out = forward(feat_a, W, b, rec)
            e = e0 if save_e else torch.sigmoid(out - 10.0)
            masks = masks0 if save_mask else [(z > 0).to(dt) for z in zs]
            hs = [feat_a] + [torch.relu(z) for z in zs]
            gv = backward(feat_a, W, masks, hs, dimg * w_a * e * (1 - e), dg, wg)
            rep(f'recompute={rec} dgrad={dg} wgrad={wg} e={"saved" if save_e else "rec"} mask={"saved" if save_mask else "rec"}', gv)


if __name__ == '__main__':
    main()
