"""CPU oracle for the bhnerf render/train hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain torch-on-CPU (float64 by default), of the
reference algorithm the CUDA kernels in ``bhnerf_b200/csrc`` implement.  It is
imported ONLY by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py``.  The product path never routes through it.

Parity pin status
-----------------
* warp / domain fill / ray integral / posenc are pinned against the reference's
  OWN source executed under a numpy shim (``oracle/ref_shim.py``,
  ``tests/golden/make_golden.py`` -> ``tests/golden/ref_*.npz``).
* The flax ``MLP`` (``nn.Dense`` + he_uniform + relu + skip), ``nn.sigmoid``,
  ``jax.value_and_grad``, ``optax.adam``/``polynomial_schedule`` live in third-party
  packages that are NOT in /root/reference and not installable here (flax/optax/jax,
  versions unpinned by the reference: requirements.txt:13 says flax==0.3.4 but the
  code uses flax.linen).  Their published definitions are restated below and anchored
  on the reference's call sites; for those pieces parity is **unpinned** by any
  reference-held golden vector (the reference ships no tests at all).

Every function cites the reference file:line it follows (paths under /root/reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

# GM/c^3 for Sgr A* (M = 4.154e6 Msun) in hours.  bhnerf/constants.py:13,17 with
# astropy's CODATA2018 G, c and IAU2015 GM_sun = 1.3271244e20 m^3/s^2.
GM_SUN = 1.3271244e20
C_LIGHT = 299792458.0
SGRA_MASS_MSUN = 4.154e6
GM_C3_SGRA_SEC = GM_SUN * SGRA_MASS_MSUN / C_LIGHT ** 3
GM_C3_SGRA_HR = GM_C3_SGRA_SEC / 3600.0

POSENC_DEG = 3
NET_DEPTH = 4
NET_WIDTH = 128
N_FEAT = 3 + 2 * 3 * POSENC_DEG  # 21
LAYER_SHAPES = [(N_FEAT, NET_WIDTH), (NET_WIDTH, NET_WIDTH), (NET_WIDTH, NET_WIDTH),
                (NET_WIDTH + N_FEAT, NET_WIDTH), (NET_WIDTH, 1)]
N_PARAMS = sum(i * o + o for i, o in LAYER_SHAPES)  # 55169


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def init_params(seed=1, dtype=np.float32):
    """he_uniform kernels, zero biases: flax ``nn.Dense(kernel_init=he_uniform())``
    (bhnerf/network.py:49-50).  he_uniform = variance_scaling(2.0, 'fan_in', 'uniform')
    -> U(-sqrt(6/fan_in), +sqrt(6/fan_in)).  JAX's threefry stream is not reproduced
    (weights are inputs to the path); a numpy Generator with a fixed seed is used."""
    rng = np.random.default_rng(seed)
    params = OrderedDict()
    for i, (fi, fo) in enumerate(LAYER_SHAPES):
        lim = math.sqrt(6.0 / fi)
        params[f'Dense_{i}'] = {
            'kernel': rng.uniform(-lim, lim, size=(fi, fo)).astype(dtype),
            'bias': np.zeros((fo,), dtype=dtype),
        }
    return {'MLP_0': params}


def trained_like_params(seed=7, dtype=np.float32, bias_out=9.0):
    """A parameter set whose emission is O(0.1-1) in places so sigmoid(o-10) is exercised off
    its tail (he_uniform init alone gives e ~ 5e-5 everywhere)."""
    p = init_params(seed, dtype)
    rng = np.random.default_rng(seed + 1000)
    d = p['MLP_0']
    for i in range(5):
        d[f'Dense_{i}']['bias'] = rng.normal(0, 0.1, d[f'Dense_{i}']['bias'].shape).astype(dtype)
    d['Dense_4']['kernel'] = (d['Dense_4']['kernel'] * 6.0).astype(dtype)
    d['Dense_4']['bias'] = np.full((1,), bias_out, dtype=dtype)
    return p


def flatten_params(params):
    """Flat packed layout used by the C ABI: for layer i = 0..4: kernel (in,out) row-major,
    then bias (out,).  Dense_3 rows 0..127 = hidden, 128..148 = posenc skip (concat order at
    bhnerf/network.py:61: ``jnp.concatenate([x, inputs])``)."""
    d = params['MLP_0'] if 'MLP_0' in params else params
    out = []
    for i in range(5):
        out.append(np.asarray(d[f'Dense_{i}']['kernel']).reshape(-1))
        out.append(np.asarray(d[f'Dense_{i}']['bias']).reshape(-1))
    return np.concatenate(out)


def unflatten_params(flat):
    flat = np.asarray(flat)
    d = OrderedDict()
    o = 0
    for i, (fi, fo) in enumerate(LAYER_SHAPES):
        k = flat[o:o + fi * fo].reshape(fi, fo); o += fi * fo
        b = flat[o:o + fo]; o += fo
        d[f'Dense_{i}'] = {'kernel': k, 'bias': b}
    return {'MLP_0': d}


# --------------------------------------------------------------------------------------
# geometry helpers
# --------------------------------------------------------------------------------------
def rotation_matrix(axis, angle):
    """Euler-Rodrigues rotation matrix, shape (3,3,...) -- bhnerf/utils.py:97-132."""
    axis = torch.as_tensor(axis, dtype=angle.dtype)
    axis = axis / torch.sqrt(torch.dot(axis, axis))
    a = torch.cos(angle / 2.0)
    s = torch.sin(angle / 2.0)
    b, c, d = [-ax * s for ax in axis]
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    rows = [
        torch.stack([aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)]),
        torch.stack([2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)]),
        torch.stack([2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]),
    ]
    return torch.stack(rows)


def warp_time(t_frames, t_start_obs, t_geos, t_injection, GM_c3, time_dtype):
    """t_M = ((t_frames - t_start_obs)/GM_c3 + t_geos) - t_injection, in this op order
    (bhnerf/emission.py:200-201).  ``time_dtype=float32`` reproduces the reference's float32
    rounding of this stage bit-for-bit (IEEE sub/div/add/sub, no contraction possible)."""
    tf = _t(t_frames, time_dtype).reshape(-1)
    tg = _t(t_geos, time_dtype)
    ts = torch.tensor(float(t_start_obs), dtype=time_dtype)
    c = torch.tensor(float(GM_c3), dtype=time_dtype)
    ti = torch.tensor(float(t_injection), dtype=time_dtype)
    tfe = tf.reshape((-1,) + (1,) * tg.dim())
    t_g = (tfe - ts) / c + tg
    return t_g - ti


def velocity_warp_coords(coords, Omega, t_frames, t_start_obs, t_geos, t_injection,
                         GM_c3=GM_C3_SGRA_HR, dtype=torch.float64, time_dtype=None,
                         rot_axis=(0.0, 0.0, 1.0)):
    """bhnerf/emission.py:143-211.  Returns (warped (Bt,...,3) with NaN before injection)."""
    time_dtype = time_dtype or dtype
    coords_t = _t(coords, dtype)
    tg = _t(t_geos, time_dtype)
    if tg.dim() == 0:                                              # scalar t_geos broadcasts like Omega (:196-200)
        tg = tg.reshape((1,) * (coords_t.dim() - 1))
    t_M = warp_time(t_frames, t_start_obs, tg, t_injection, GM_c3, time_dtype)
    Om = _t(Omega, time_dtype)
    if Om.dim() == 0:                                              # emission.py:190-191: scalar Omega
        Om = Om.reshape((1,) * (coords_t.dim() - 1))
    theta = (t_M * Om)                                            # emission.py:204
    theta = torch.where(t_M < 0.0, torch.full_like(theta, float('nan')), theta)  # :205
    theta = theta.to(dtype)
    R = rotation_matrix(rot_axis, -theta)                          # :207, (3,3,Bt,...)
    warped = torch.sum(R * coords_t.unsqueeze(1).unsqueeze(0), dim=1)  # :209 (3,Bt,...)
    return torch.movedim(warped, 0, -1)                            # :210


def safe_sin(x):
    """``jnp.sin(x % (100*pi))`` with python-sign modulo (bhnerf/network.py:16)."""
    return torch.sin(torch.remainder(x, 100.0 * math.pi))


def posenc(x, deg=POSENC_DEG):
    """bhnerf/network.py:98-122: [x | sin(2^i x) scale-major | sin(2^i x + pi/2)]."""
    if deg == 0:
        return x
    scales = torch.tensor([2.0 ** i for i in range(deg)], dtype=x.dtype)
    xb = (x[..., None, :] * scales[:, None]).reshape(list(x.shape[:-1]) + [-1])
    four = safe_sin(torch.cat([xb, xb + 0.5 * math.pi], dim=-1))
    return torch.cat([x, four], dim=-1)


def posenc_ref32(x, deg=POSENC_DEG):
    """posenc with the sin ARGUMENTS formed in the reference's float32 arithmetic, sin in x.dtype.

    On float32 inputs (what JAX runs) the reference computes  xb = x*2^i,  xb + fl32(pi/2)  and the
    python-sign modulo by fl32(100*pi) = 314.15927 in float32 (bhnerf/network.py:16,118-121).  For every
    negative argument that modulo ADDS 314.15927f and rounds to the float32 grid at 314 (ulp 3.05e-5),
    i.e. a +5.9e-6 bias plus up to 1.5e-5 rounding -- deterministic IEEE arithmetic that is part of the
    reference's result (it moves gradients by up to 1.4e-2 relative).  This function
    reproduces those arguments bit-for-bit and then takes an exact sin, so it equals the reference up to
    libm ulp noise; ``posenc`` above is the same formula evaluated entirely in x.dtype."""
    if deg == 0:
        return x
    x32 = x.detach().to(torch.float32)
    scales = torch.tensor([2.0 ** i for i in range(deg)], dtype=torch.float32)
    xb = (x32[..., None, :] * scales[:, None]).reshape(list(x.shape[:-1]) + [-1])
    args = torch.cat([xb, xb + torch.tensor(0.5 * math.pi, dtype=torch.float32)], dim=-1)
    args = torch.remainder(args, torch.tensor(100.0 * math.pi, dtype=torch.float32))
    return torch.cat([x, torch.sin(args.to(x.dtype))], dim=-1)


def mlp_forward(p, x):
    """flax MLP, bhnerf/network.py:18-64: 4x128 relu, concat(inputs) after layer i=2, Dense(1)."""
    inputs = x
    skip_layer = NET_DEPTH // 2
    for i in range(NET_DEPTH):
        x = torch.relu(x @ p[f'Dense_{i}']['kernel'] + p[f'Dense_{i}']['bias'])
        if i % skip_layer == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    return x @ p['Dense_4']['kernel'] + p['Dense_4']['bias']


def fill_unsupervised_emission(emission, coords, rmin, rmax, z_width):
    """bhnerf/emission.py:343-374 (un-warped coords; strict inequalities)."""
    r_sq = coords[0] ** 2 + coords[1] ** 2 + coords[2] ** 2
    zero = torch.zeros_like(emission)
    emission = torch.where(r_sq < rmin ** 2, zero, emission)
    emission = torch.where(r_sq > rmax ** 2, zero, emission)
    emission = torch.where(torch.abs(coords[2]) > z_width, zero, emission)
    return emission


def _params_t(params, dtype, requires_grad=False):
    d = params['MLP_0'] if 'MLP_0' in params else params
    out = OrderedDict()
    for k, v in d.items():
        out[k] = {kk: _t(vv, dtype).clone().requires_grad_(requires_grad) for kk, vv in v.items()}
    return out


def predict_emission(p, t_frames, coords, Omega, t_start_obs, t_geos, t_injection,
                     scale, rmin, rmax, z_width, GM_c3=GM_C3_SGRA_HR,
                     dtype=torch.float64, time_dtype=None, feature_mode='ref32'):
    """NeRF_Predictor.__call__, bhnerf/network.py:191-237.  ``p`` = torch param dict.
    feature_mode 'ref32' (default): posenc_ref32 (reference float32 argument arithmetic);
    'pure': posenc evaluated entirely in ``dtype``."""
    coords_t = _t(coords, dtype)
    warped = velocity_warp_coords(coords_t, Omega, t_frames, t_start_obs, t_geos, t_injection,
                                  GM_c3, dtype, time_dtype)
    valid = torch.isfinite(warped)                                   # network.py:226
    net_in = torch.where(valid, warped, torch.zeros_like(warped))    # :227
    u = net_in / scale
    if feature_mode == 'ref32':
        u = u.to(torch.float32).to(dtype)                           # the reference's MLP input is float32
        feat = posenc_ref32(u, POSENC_DEG)
    else:
        feat = posenc(u, POSENC_DEG)
    out = mlp_forward(p, feat)                                       # :229
    emission = torch.sigmoid(out[..., 0] - 10.0)                     # :230
    emission = fill_unsupervised_emission(emission, coords_t, rmin, rmax, z_width)  # :231
    emission = torch.where(valid[..., 0], emission, torch.zeros_like(emission))     # :232
    return emission


def radiative_trasfer(emission, g, dtau, Sigma):
    """bhnerf/kgeo.py:595-622: sum over the last (geo) axis of g^2 * e * dtau * Sigma."""
    return (g ** 2 * emission * dtau * Sigma).sum(dim=-1)


def image_plane_prediction(p, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos,
                           t_injection, scale, rmin, rmax, z_width, GM_c3=GM_C3_SGRA_HR,
                           dtype=torch.float64, time_dtype=None, squeeze=True, feature_mode='ref32'):
    """bhnerf/network.py:373-420.  Returns (Bt,A,B) for scalar J, (Bt,S,A,B) otherwise
    (with jnp.squeeze dropping size-1 axes, :418, when ``squeeze``)."""
    emission = predict_emission(p, t_frames, coords, Omega, t_start_obs, t_geos, t_injection,
                                scale, rmin, rmax, z_width, GM_c3, dtype, time_dtype, feature_mode)
    if not np.isscalar(J):
        Jt = _t(J, dtype)
        emission = Jt.unsqueeze(0) * emission.unsqueeze(1)           # :416-417
        if squeeze:
            emission = torch.squeeze(emission)                       # :418
    return radiative_trasfer(emission, _t(g, dtype), _t(dtau, dtype), _t(Sigma, dtype))


def loss_image(images, target, sigma, offset, scale, dtype_str):
    """bhnerf/network.py:476-484."""
    if dtype_str == 'full':
        return scale * torch.sum(torch.abs((images - target - offset) / sigma) ** 2)
    if dtype_str == 'lc':
        lc = images.sum(dim=(-1, -2))
        return scale * torch.sum(torch.abs((lc - target - offset) / sigma) ** 2)
    raise AttributeError('image dtype ({}) not supported'.format(dtype_str))


def loss_eht(images, target, sigma, A, scale, dtype_str):
    """bhnerf/network.py:542-564.  A: (Bt,[3,]V,P) complex; images (Bt,A,B)."""
    cdt = torch.complex128 if images.dtype == torch.float64 else torch.complex64
    vec = images.reshape(*images.shape[:-2], -1, 1).to(cdt)          # :542
    while vec.dim() < A.dim():
        vec = vec.unsqueeze(-3)                                      # :543
    vis = torch.matmul(A.to(cdt), vec).squeeze(-1)                   # :544
    if dtype_str == 'vis':
        chisq = torch.sum((torch.abs(vis - target) / sigma) ** 2)    # :548
    elif dtype_str == 'amp':
        chisq = torch.sum(torch.abs((torch.abs(vis) - target) / sigma) ** 2)  # :553
    elif dtype_str == 'cphase':
        clphase = torch.angle(torch.prod(vis, dim=-2))               # :558
        chisq = torch.sum((1.0 - torch.cos(target - clphase)) / (sigma ** 2))  # :559
    else:
        raise AttributeError('eht dtype ({}) not supported'.format(dtype_str))
    return scale * chisq, vis


def value_and_grad(params, loss_kind, dtype_str, target, sigma, third, t_frames, rt, predictor,
                   scale=1.0, GM_c3=GM_C3_SGRA_HR, dtype=torch.float64, time_dtype=torch.float32,
                   with_grad=True, feature_mode='ref32'):
    """jax.value_and_grad(loss_fn_*, argnums=0) -- bhnerf/network.py:617, :677.

    rt: dict with coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection
        (the positional ABI of raytracing_args, bhnerf/network.py:874-892).
    predictor: dict(scale, rmin, rmax, z_width).  third = offset ('image') or A ('eht').
    Returns dict(loss, images, grads(flat f64 numpy), [vis])."""
    p = _params_t(params, dtype, requires_grad=with_grad)
    images = image_plane_prediction(
        p, t_frames, rt['coords'], rt['Omega'], rt['J'], rt['g'], rt['dtau'], rt['Sigma'],
        rt['t_start_obs'], rt['t_geos'], rt['t_injection'], predictor['scale'], predictor['rmin'],
        predictor['rmax'], predictor['z_width'], GM_c3, dtype, time_dtype, squeeze=False,
        feature_mode=feature_mode)
    out = {}
    if loss_kind == 'image':
        loss = loss_image(images, _t(target, dtype), _t(sigma, dtype), _t(third, dtype), scale,
                          dtype_str)
    else:
        cdt = torch.complex128 if dtype == torch.float64 else torch.complex64
        tgt = torch.as_tensor(np.asarray(target))
        tgt = tgt.to(cdt) if tgt.is_complex() else tgt.to(dtype)
        loss, vis = loss_eht(images, tgt, _t(sigma, dtype), torch.as_tensor(np.asarray(third)), scale,
                             dtype_str)
        out['vis'] = vis.detach().numpy()
    out['loss'] = float(loss.detach())
    out['images'] = images.detach().numpy()
    if with_grad:
        loss.backward()
        flat = []
        for i in range(5):
            flat.append(p[f'Dense_{i}']['kernel'].grad.reshape(-1))
            flat.append(p[f'Dense_{i}']['bias'].grad.reshape(-1))
        out['grads'] = torch.cat(flat).numpy()
    return out


# --------------------------------------------------------------------------------------
# optimiser: optax.adam(learning_rate=polynomial_schedule(lr_init, lr_final, 1, num_iters))
# bhnerf/network.py:171-182, applied at :621 / :681 by TrainState.apply_gradients.
# --------------------------------------------------------------------------------------
def polynomial_schedule(count, init_value, end_value, power, transition_steps):
    """optax.polynomial_schedule (transition_begin=0): count clipped to [0, transition_steps];
    lr = (init-end) * (1 - count/transition_steps)^power + end."""
    c = min(max(count, 0), transition_steps)
    frac = 1.0 - c / transition_steps
    return (init_value - end_value) * (frac ** power) + end_value


def adam_step(flat_params, flat_grads, mu, nu, count, lr_init=1e-4, lr_final=1e-6, num_iters=5000,
              b1=0.9, b2=0.999, eps=1e-8):
    """One optax.adam update (scale_by_adam then scale_by_learning_rate(schedule(count))).
    ``count`` is the number of updates already applied (0 for the first step): optax evaluates
    the schedule at the pre-increment count and bias-corrects with count+1."""
    g = np.asarray(flat_grads, dtype=np.float64)
    mu = b1 * mu + (1 - b1) * g
    nu = b2 * nu + (1 - b2) * g * g
    t = count + 1
    mu_hat = mu / (1 - b1 ** t)
    nu_hat = nu / (1 - b2 ** t)
    lr = polynomial_schedule(count, lr_init, lr_final, 1, num_iters)
    new_params = np.asarray(flat_params, dtype=np.float64) - lr * mu_hat / (np.sqrt(nu_hat) + eps)
    return new_params, mu, nu


# --------------------------------------------------------------------------------------
# numpy synthesiser (bhnerf/emission.py:213-303): warp -> trilinear 64^3 lookup -> RT
# --------------------------------------------------------------------------------------
def world_to_image_coords(coords, fov, npix):
    """bhnerf/utils.py world_to_image_coords: image = (coords + fov/2) / fov * (npix - 1)."""
    out = []
    for i in range(coords.shape[-1]):
        out.append((coords[..., i] + fov[i] / 2.0) / fov[i] * (npix[i] - 1))
    return np.stack(out, axis=-1)


def image_plane_dynamics(emission_0, grid_fov, rt, t_frames, GM_c3=GM_C3_SGRA_HR):
    """emission_0: (nx,ny,nz) numpy grid spanning [-fov/2, fov/2]^3 (xarray coords in the
    reference); trilinear ``map_coordinates(order=1, cval=0)``; NaN coords (pre-injection)
    propagate NaN in the reference -- callers there never hit them (t_injection=-r_o)."""
    import scipy.ndimage
    warped = velocity_warp_coords(rt['coords'], rt['Omega'], t_frames, rt['t_start_obs'], rt['t_geos'],
                                  rt['t_injection'], GM_c3, torch.float64).numpy()
    npix = emission_0.shape
    fov = [grid_fov] * 3
    ic = np.moveaxis(world_to_image_coords(warped, fov, npix), -1, 0)
    e = scipy.ndimage.map_coordinates(emission_0, ic, order=1, cval=0.0)
    J = rt['J']
    if not np.isscalar(J):
        e = np.asarray(J)[None] * e[:, None]
    g, dtau, Sigma = [np.asarray(rt[k], dtype=np.float64) for k in ('g', 'dtau', 'Sigma')]
    return (g ** 2 * e * dtau * Sigma).sum(axis=-1)


# --------------------------------------------------------------------------------------
# GRID_Predictor (bhnerf/network.py:254-357): voxel grid + jax.scipy.ndimage.map_coordinates
# --------------------------------------------------------------------------------------
def jax_map_coordinates_order1(grid, ic):
    """jax.scipy.ndimage.map_coordinates(grid, ic, order=1, mode='constant', cval=0) restated in torch
    (jax/_src/scipy/ndimage.py, published algorithm; jax is not installed): per axis the two neighbours
    floor(c), floor(c)+1 with weights 1-frac, frac; a neighbour outside [0, n) contributes cval (0)
    -- corner by corner, unlike scipy's 'constant' mode which returns cval for any point outside the extent.
    grid: torch (nx,ny,nz); ic: torch (3, ...).  Differentiable w.r.t. grid."""
    n = grid.shape
    lo = torch.floor(ic)
    fr = ic - lo
    lo = lo.long()
    out = torch.zeros(ic.shape[1:], dtype=grid.dtype)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                idx = [lo[0] + dx, lo[1] + dy, lo[2] + dz]
                ok = torch.ones_like(out, dtype=torch.bool)
                w = torch.ones_like(out)
                for a, d in enumerate((dx, dy, dz)):
                    ok &= (idx[a] >= 0) & (idx[a] < n[a])
                    w = w * (fr[a] if d else 1 - fr[a])
                val = grid[idx[0].clamp(0, n[0] - 1), idx[1].clamp(0, n[1] - 1), idx[2].clamp(0, n[2] - 1)]
                out = out + torch.where(ok, val, torch.zeros_like(val)) * w
    return out


def grid_predictor_images(grid, t_frames, rt, predictor, GM_c3=GM_C3_SGRA_HR, dtype=torch.float64):
    """GRID_Predictor.__call__ + image_plane_prediction (network.py:306-357, 373-420).  grid: torch tensor
    (res,res,res) (requires_grad allowed).  Returns images (nt, [S,] A, B) as a torch tensor."""
    coords = _t(rt['coords'], dtype)
    warped = velocity_warp_coords(rt['coords'], rt['Omega'], t_frames, rt['t_start_obs'], rt['t_geos'],
                                  rt['t_injection'], GM_c3, dtype)                       # (nt, ..., 3), NaN before injection
    valid = torch.isfinite(warped)
    net_in = torch.where(valid, warped, torch.zeros_like(warped)).movedim(-1, 0)
    res = grid.shape[0]
    net_in = (net_in + predictor['scale']) / (2 * predictor['scale']) * (res - 1.0)
    e = torch.sigmoid(jax_map_coordinates_order1(grid, net_in) - 10.0)
    e = fill_unsupervised_emission(e, coords, predictor['rmin'], predictor['rmax'], predictor['z_width'])
    e = torch.where(valid[..., 0], e, torch.zeros_like(e))
    J = rt['J']
    if not np.isscalar(J):
        e = _t(J, dtype)[None] * e[:, None]
    g, dtau, Sigma = [_t(rt[k], dtype) for k in ('g', 'dtau', 'Sigma')]
    return (g ** 2 * e * dtau * Sigma).sum(-1)


# --------------------------------------------------------------------------------------
# Polarization factors J = (I, Q, U): alma.image_plane_model's chain (bhnerf/alma.py:47-60) =
# azimuthal_velocity_vector (kgeo.py:199-223) -> doppler_factor (:225-248) -> magnetic_field_fluid_frame (:274-313,
# with fluid_frame_tetrad :315-350) -> parallel_transport (:438-519, V_frac = 0), float64.  For a purely azimuthal
# 4-velocity (u^r = u^theta = 0, u.u = -1) the tetrad of :338-347 collapses to
#   e_t = -u,  e_r = (0, 1, 0, 0)/sqrt(g_rr),  e_th = (0, 0, 1, 0)/sqrt(g_thth),  e_ph = (u_ph, 0, 0, -u_t)/(sqrt(Delta) sin th)
# (u_t, u_ph covariant), which is what is written out below; pinned on the reference's own functions in
# tests/test_oracle.py (oracle/ref_shim.reference_polarization_factors).
# --------------------------------------------------------------------------------------
def polarization_factors(geos, Omega, b_consts, Q_frac, rmin, rmax, z_width, spectral_index=1):
    """geos: dict of float64 arrays (..., G) with r, theta, affine, lam, eta, alpha, beta (per-ray values broadcast) and
    scalars spin, inc (E = M = 1).  Returns J (3, ..., G) with NaN -> 0 (bhnerf/alma.py:60)."""
    r, th, aff = [np.asarray(geos[k], dtype=np.float64) for k in ('r', 'theta', 'affine')]
    lam, eta, alpha, beta = [np.asarray(geos[k], dtype=np.float64) for k in ('lam', 'eta', 'alpha', 'beta')]
    a, inc = float(geos['spin']), float(geos['inc'])
    Om = np.asarray(Omega, dtype=np.float64)
    arad, avert, ator = b_consts['arad'], b_consts['avert'], b_consts['ator']
    with np.errstate(all='ignore'):
        sth, cth = np.sin(th), np.cos(th)
        Delta = r ** 2 + a ** 2 - 2 * r
        Sigma = r ** 2 + a ** 2 * cth ** 2
        Xi = (r ** 2 + a ** 2) ** 2 - a ** 2 * Delta * sth ** 2
        Rpot = (r ** 2 + a ** 2 - a * lam) ** 2 - Delta * (eta + (lam - a) ** 2)
        Rpot = np.where(np.abs(Rpot) > 1e-10, Rpot, 0.0)
        Thpot = eta + a ** 2 * cth ** 2 - lam ** 2 / np.tan(th) ** 2
        # wave vector k_mu (covariant), kgeo.py:91-117; the signs follow the turning points along the ray
        pm_r = np.sign(np.gradient(r, axis=-1) / np.gradient(aff, axis=-1))
        pm_th = np.sign(np.gradient(th, axis=-1) / np.gradient(aff, axis=-1))
        k_t, k_r = -np.ones_like(r), np.sqrt(np.clip(Rpot, 0, None)) * pm_r / Delta
        k_th, k_ph = np.sqrt(np.clip(Thpot, 0, None)) * pm_th, lam
        # metric and the azimuthal 4-velocity
        g_tt, g_rr, g_thth = -(1 - 2 * r / Sigma), Sigma / Delta, Sigma
        g_phph, g_tph = Xi * sth ** 2 / Sigma, -2 * a * r * sth ** 2 / Sigma
        ut = 1 / np.sqrt(-(g_tt + 2 * Om * g_tph + g_phph * Om ** 2))
        uph = ut * Om
        u_t, u_ph = g_tt * ut + g_tph * uph, g_phph * uph + g_tph * ut          # covariant
        s = u_t * ut + u_ph * uph                                               # = -1 where the orbit is timelike
        N_r, N_th, N_ph = np.sqrt(-g_rr * s), np.sqrt(g_thth), np.sqrt(-s * Delta * sth ** 2)
        gdop = 1.0 / -(k_t * ut + k_ph * uph)                                   # doppler_factor, NaN -> 0 (kgeo.py:246)
        gdop = np.where(np.isnan(gdop), 0.0, gdop)
        # k in the fluid frame (spatial part): k'_j = e_j^mu k_mu
        kp = np.stack([-s * k_r / N_r, k_th / N_th, (u_ph * k_t - u_t * k_ph) / N_ph], axis=-1)
        # fluid-frame magnetic field (kgeo.py:290-313)
        Br, Bth, Bph = arad * sth + avert * cth, -avert * sth, ator * np.ones_like(r)
        b0 = Bph * u_ph
        b1, b2, b3 = Br / u_t, Bth / u_t, (Bph + b0 * u_ph) / u_t
        bl = [g_tt * b0 + g_tph * b3, g_rr * b1, g_thth * b2, g_phph * b3 + g_tph * b0]
        bp = np.stack([-s * bl[1] / N_r, bl[2] / N_th, (u_ph * bl[0] - u_t * bl[3]) / N_ph], axis=-1)
        z = r * cth
        domain = (np.abs(z) < z_width) & (r > rmin) & (r < rmax)
        b_mean = np.sqrt((bp[domain] ** 2).sum(-1)).mean()                      # alma.py:55-57
        bp = bp / b_mean
        # parallel_transport (kgeo.py:476-518)
        k_mag = np.sqrt((kp ** 2).sum(-1))
        f_loc = np.cross(kp, bp, axis=-1) / k_mag[..., None]
        f_t, f_r = u_ph / N_ph * f_loc[..., 2], -s / N_r * f_loc[..., 0]
        f_th, f_ph = f_loc[..., 1] / N_th, -u_t / N_ph * f_loc[..., 2]
        b_mag = np.sqrt((bp ** 2).sum(-1))
        sin_b = np.sqrt((f_loc ** 2).sum(-1)) / k_mag
        I = gdop ** spectral_index * b_mag ** (spectral_index + 1) * sin_b ** (spectral_index + 1)
        Q = Q_frac * I
        gi_tt, gi_rr, gi_thth = -Xi / (Delta * Sigma), Delta / Sigma, 1 / Sigma
        gi_phph, gi_tph = (Delta - a ** 2 * sth ** 2) / (Delta * Sigma * sth ** 2), -2 * a * r / (Delta * Sigma)
        ku = [gi_tt * k_t + gi_tph * k_ph, gi_rr * k_r, gi_thth * k_th, gi_phph * k_ph + gi_tph * k_t]
        A = (ku[0] * f_r - ku[1] * f_t) + a * sth ** 2 * (ku[1] * f_ph - ku[3] * f_r)
        B = ((r ** 2 + a ** 2) * (ku[3] * f_th - ku[2] * f_ph) - a * (ku[0] * f_th - ku[2] * f_t)) * sth
        kappa = (r - 1j * a * cth) * (A - 1j * B)
        mu = -(alpha + a * np.sin(inc))
        chi2 = np.angle(((beta + 1j * mu) * np.conj(kappa)) / ((beta - 1j * mu) * kappa))
        J = np.stack([I, np.cos(chi2) * Q, np.sin(chi2) * Q])
    return np.nan_to_num(J, nan=0.0)
