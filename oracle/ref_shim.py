"""Loader that executes the REFERENCE'S OWN source (read-only, /root/reference) under numpy.

TEST INFRASTRUCTURE ONLY, and usable only in the build container: /root/reference does not
exist on the GPU box.  Used by ``tests/golden/make_golden.py`` to generate the committed golden
vectors (``make_golden*.py``) and by the live-reference checks of ``tests/test_oracle.py`` (skipped when the reference is
absent).

jax / flax / astropy / xarray / kgeo's plotting deps are not installed, so:
  * ``bhnerf/{utils,constants,emission,kgeo}.py`` are loaded by path with ``jax.numpy`` aliased
    to numpy and inert stubs for astropy/xarray/matplotlib (their hot-path helpers are written
    against ``_np = jnp if use_jax else np``);
  * ``posenc``/``safe_sin`` are exec'd from ``bhnerf/network.py`` lines 16 and 98-122;
  * kgeo's analytic Kerr ray tracer is loaded by file path (appendix A.1).
Nothing is copied from the reference into this repository.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get('BHNERF_REFERENCE', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF, 'bhnerf', 'emission.py'))


_loaded = {}


def _install_stubs():
    if 'stubs' in _loaded:
        return
    if not hasattr(np, 'infty'):
        np.infty = np.inf
    if not hasattr(np, 'NaN'):
        np.NaN = np.nan
    if not hasattr(np, 'Inf'):
        np.Inf = np.inf
    j = types.ModuleType('jax'); j.numpy = np
    u = types.ModuleType('astropy.units'); u.Quantity = type('Quantity', (), {}); u.lightyear = 1.0
    u.hr = 'hr'
    c = types.ModuleType('astropy.constants'); c.G = c.c = c.M_sun = 1.0; c.__all__ = ['G', 'c', 'M_sun']
    a = types.ModuleType('astropy'); a.units = u; a.constants = c
    mods = {'jax': j, 'jax.numpy': np, 'astropy': a, 'astropy.units': u, 'astropy.constants': c}
    for n in ['matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d', 'h5py']:
        mods[n] = types.ModuleType(n)
    mods['xarray'] = _mini_xarray()
    mods['mpl_toolkits.mplot3d'].Axes3D = object
    for k, v in mods.items():
        sys.modules.setdefault(k, v)
    _loaded['stubs'] = True


class _DA(np.ndarray):
    """The sliver of xarray.DataArray that bhnerf/kgeo.py's tetrad / parallel-transport algebra uses (xarray is not
    installed): an ndarray that remembers where its 4-vector axis 'mu' sits ('first' after xr.concat(dim='mu'), 'last'
    after .transpose(..., 'mu')), so that .sel(mu=i) works; every other dimension broadcasts positionally -- the callers
    below pre-broadcast the per-ray coordinates (alpha, beta, lam) to the full (beta, alpha, geo) shape, which is what
    xarray's name-based alignment would do."""
    mu = None

    def __new__(cls, a, mu=None):
        obj = np.asarray(a, dtype=np.float64 if not np.iscomplexobj(a) else None).view(cls)
        obj.mu = mu
        return obj

    def __array_finalize__(self, obj):
        self.mu = getattr(obj, 'mu', None)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        mus = [getattr(x, 'mu', None) for x in inputs]
        mu = 'first' if 'first' in mus else ('last' if 'last' in mus else None)
        args = [np.asarray(x) if isinstance(x, _DA) else x for x in inputs]
        if 'first' in mus and 'last' in mus:               # xarray aligns by name: bring every 'mu' axis to the front
            args = [np.moveaxis(a, -1, 0) if m == 'last' else a for a, m in zip(args, mus)]
        if 'out' in kw:
            kw['out'] = tuple(np.asarray(o) if isinstance(o, _DA) else o for o in kw['out'])
        res = getattr(ufunc, method)(*args, **kw)
        return _DA(res, mu) if isinstance(res, np.ndarray) else res

    def sel(self, mu):
        a = np.asarray(self)
        return _DA(a[mu] if self.mu == 'first' else a[..., mu])

    def transpose(self, *dims):
        if dims == (Ellipsis, 'mu') and self.mu == 'first':
            return _DA(np.moveaxis(np.asarray(self), 0, -1), 'last')
        if not dims or (dims == (Ellipsis, 'mu') and self.mu == 'last'):
            return self if dims else _DA(np.asarray(self).T, self.mu)
        raise NotImplementedError(dims)

    def clip(self, min=None, max=None):
        return _DA(np.clip(np.asarray(self), min, max), self.mu)

    def fillna(self, value):
        a = np.asarray(self)
        return _DA(np.where(np.isnan(a), value, a), self.mu)

    def sum(self, dim=None, axis=None, skipna=None, **kw):
        a = np.asarray(self)
        if dim == 'mu':
            return _DA(a.sum(axis=0 if self.mu == 'first' else -1))
        return _DA(a.sum(axis=axis, **kw), self.mu if axis is None else None) if axis is not None else a.sum(**kw)


def _mini_xarray():
    m = types.ModuleType('xarray')
    m.DataArray = lambda v=0.0, **kw: _DA(v)

    def concat(items, dim, coords=None):
        assert dim == 'mu'
        arrs = [np.asarray(x, dtype=np.float64) for x in items]
        shape = np.broadcast_shapes(*[a.shape for a in arrs])
        return _DA(np.stack([np.broadcast_to(a, shape) for a in arrs], axis=0), 'first')
    m.concat = concat
    m.full_like = lambda a, fill_value: _DA(np.full_like(np.asarray(a), fill_value))
    m.Dataset = lambda d=None, **kw: types.SimpleNamespace(**{k: _DA(v) for k, v in (d or {}).items()})
    return m


def _load(pkgname, name, path):
    spec = importlib.util.spec_from_file_location(pkgname + '.' + name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[pkgname + '.' + name] = m
    spec.loader.exec_module(m)
    return m


def load_kgeo():
    """kgeo's analytic tracer (kgeo/kgeo/kerr_raytracing_ana.py), bypassing kgeo/__init__.py."""
    if 'kgeo' in _loaded:
        return _loaded['kgeo']
    _install_stubs()
    pkg = types.ModuleType('kgeo'); pkg.__path__ = [os.path.join(REF, 'kgeo', 'kgeo')]
    sys.modules['kgeo'] = pkg
    base = os.path.join(REF, 'kgeo', 'kgeo')
    _load('kgeo', 'kerr_raytracing_utils', os.path.join(base, 'kerr_raytracing_utils.py'))
    _load('kgeo', 'scipy_ellip_binding', os.path.join(base, 'scipy_ellip_binding.py'))
    ana = _load('kgeo', 'kerr_raytracing_ana', os.path.join(base, 'kerr_raytracing_ana.py'))
    _loaded['kgeo'] = ana
    return ana


def load_bhnerf():
    """Returns a namespace with the reference's utils, emission, kgeo(bhnerf) modules + posenc."""
    if 'bhnerf' in _loaded:
        return _loaded['bhnerf']
    _install_stubs()
    if 'kgeo' not in sys.modules:
        sys.modules['kgeo'] = types.ModuleType('kgeo')
    pkg = types.ModuleType('bhnerf'); pkg.__path__ = [os.path.join(REF, 'bhnerf')]
    sys.modules['bhnerf'] = pkg
    base = os.path.join(REF, 'bhnerf')
    ns = types.SimpleNamespace()
    ns.utils = _load('bhnerf', 'utils', os.path.join(base, 'utils.py')); pkg.utils = ns.utils
    ns.constants = _load('bhnerf', 'constants', os.path.join(base, 'constants.py')); pkg.constants = ns.constants
    ns.kgeo = _load('bhnerf', 'kgeo', os.path.join(base, 'kgeo.py')); pkg.kgeo = ns.kgeo
    ns.emission = _load('bhnerf', 'emission', os.path.join(base, 'emission.py')); pkg.emission = ns.emission
    # posenc / safe_sin: network.py cannot be imported (flax/optax at module level)
    lines = open(os.path.join(base, 'network.py')).read().split('\n')
    src = lines[15] + '\n' + '\n'.join(lines[97:122])
    env = {'jnp': np, 'np': np}
    exec(compile(src, 'bhnerf/network.py[16,98-122]', 'exec'), env)
    ns.posenc = env['posenc']; ns.safe_sin = env['safe_sin']
    # the same source under JAX-like float32 promotion (float32 arrays, weak python scalars)
    f32ns = types.SimpleNamespace(array=lambda v: np.array(v, dtype=np.float32), pi=np.float32(np.pi), sin=np.sin,
                                  reshape=np.reshape, concatenate=np.concatenate)
    env32 = {'jnp': f32ns, 'np': np}
    exec(compile(src, 'bhnerf/network.py[16,98-122]', 'exec'), env32)
    ns.posenc_f32 = env32['posenc']
    _loaded['bhnerf'] = ns
    return ns


def kerr_geodesics(spin, inclination, fov_M, num_alpha, num_beta, ngeo, distance=1000.0):
    """Real Kerr geodesics through the reference's tracer, then the algebra of
    ``Geodesics.get_dataset`` (kgeo/kgeo/kerr_raytracing_utils.py:220-279) and of
    ``image_plane_geos`` (bhnerf/kgeo.py:6-63) restated in numpy (xarray is absent).
    Returns float64 arrays with dims (beta, alpha, geo) exactly as the reference's Dataset."""
    ana = load_kgeo()
    alpha_1d = np.linspace(-fov_M / 2, fov_M / 2, num_alpha)
    beta_1d = np.linspace(-fov_M / 2, fov_M / 2, num_beta)
    alpha, beta = np.meshgrid(alpha_1d, beta_1d, indexing='ij')
    g = ana.raytrace_ana(float(spin), [0, float(distance), float(inclination), 0],
                         [alpha.ravel(), beta.ravel()], ngeo, plotdata=False, verbose=False)

    def rs(x):
        return np.asarray(x).reshape(-1, num_alpha, num_beta).T      # (beta, alpha, geo)
    r, theta, phi, t, mino, affine = rs(g.r_s), rs(g.th_s), rs(g.ph_s), rs(g.t_s), rs(g.tausteps), rs(g.sig_s)
    a = float(spin); M = 1.0; E = 1.0
    x = r * np.cos(phi) * np.sin(theta)
    y = r * np.sin(phi) * np.sin(theta)
    z = r * np.cos(theta)
    Delta = r ** 2 + a ** 2 - 2 * M * r
    Sigma = r ** 2 + a ** 2 * np.cos(theta) ** 2
    Xi = (r ** 2 + a ** 2) ** 2 - a ** 2 * Delta * np.sin(theta) ** 2
    alpha_c = alpha_1d[None, :, None]     # coord 'alpha' on dim 'alpha' (axis 1)
    beta_c = beta_1d[:, None, None]       # coord 'beta' on dim 'beta' (axis 0)
    lam = -alpha_c * np.sin(inclination)
    eta = beta_c ** 2 + (alpha_c ** 2 - a ** 2) * np.cos(inclination) ** 2          # kerr_raytracing_utils.py:266
    R = (r ** 2 + a ** 2 - a * lam) ** 2 - Delta * (eta + (lam - a) ** 2)            # :267
    R = np.where(np.abs(R) > 1e-10, R, 0.0)                                           # :270
    with np.errstate(divide='ignore', invalid='ignore'):
        Theta = eta + a ** 2 * np.cos(theta) ** 2 - lam ** 2 / np.tan(theta) ** 2    # :271
    dtau = np.concatenate([np.zeros_like(mino[..., :1]), np.diff(mino, axis=-1)], axis=-1)
    bc = lambda v: np.broadcast_to(v, r.shape).copy()
    out = dict(x=x, y=y, z=z, r=r, theta=theta, phi=phi, t=t, mino=mino, affine=affine, dtau=dtau, Sigma=Sigma,
               Delta=Delta, Xi=Xi, omega=2 * a * M * r / Xi, lam=bc(lam), eta=bc(eta), R=R, Theta=Theta, alpha=bc(alpha_c),
               beta=bc(beta_c), spin=a, M=M, E=E, r_o=float(distance), inc=float(inclination))
    return out


def reference_polarization_factors(geos, Omega, b_consts, Q_frac, rmin, rmax, z_width):
    """Stokes factors J = (I, Q, U) of alma.image_plane_model (bhnerf/alma.py:47-60) computed by the REFERENCE'S OWN
    functions -- kgeo.azimuthal_velocity_vector, doppler_factor, magnetic_field_fluid_frame, parallel_transport
    (bhnerf/kgeo.py:199-248, 274-313, 438-519) -- executed on the numpy geodesics of kerr_geodesics() with the mini-xarray
    above.  Returns (J [3, beta, alpha, geo] float64 with NaN -> 0, g)."""
    ns = load_bhnerf()
    G = types.SimpleNamespace(**{k: (_DA(v) if isinstance(v, np.ndarray) else v) for k, v in geos.items()})
    Om = _DA(Omega)
    umu = ns.kgeo.azimuthal_velocity_vector(G, Om)
    with np.errstate(all='ignore'):
        g = ns.kgeo.doppler_factor(G, umu)
        b = ns.kgeo.magnetic_field_fluid_frame(G, umu, **b_consts)
        domain = np.bitwise_and(np.bitwise_and(np.abs(np.asarray(G.z)) < z_width, np.asarray(G.r) > rmin), np.asarray(G.r) < rmax)
        b_mean = np.sqrt(np.sum(np.asarray(b)[domain] ** 2, axis=-1)).mean()
        b = np.asarray(b) / b_mean
        J = np.nan_to_num(np.asarray(ns.kgeo.parallel_transport(G, umu, np.asarray(g), b, Q_frac=Q_frac, V_frac=0)), nan=0.0)
    return J, np.asarray(g)


def keplerian_omega(geos):
    """Tutorial3 cell 2: sign(spin+eps) * sqrt(M) / (r^1.5 + spin*sqrt(M))."""
    return np.sign(geos['spin'] + np.finfo(float).eps) * np.sqrt(geos['M']) / (
        geos['r'] ** 1.5 + geos['spin'] * np.sqrt(geos['M']))


def doppler_factor(geos, Omega, fillna=0.0):
    """bhnerf/kgeo.py:199-248 restated: u^t from the metric, k_t=-E, k_ph=E*lam, the u^r=u^th=0
    terms vanish  =>  g = E / (E u^t - E lam u^ph)."""
    r, th, a, M = geos['r'], geos['theta'], geos['spin'], geos['M']
    Sigma, Xi = geos['Sigma'], geos['Xi']
    g_tt = -(1 - 2 * M * r / Sigma)
    g_phph = Xi * np.sin(th) ** 2 / Sigma
    g_tph = -2 * M * a * r * np.sin(th) ** 2 / Sigma
    with np.errstate(invalid='ignore', divide='ignore'):
        ut = 1 / np.sqrt(-(g_tt + 2 * Omega * g_tph + g_phph * Omega ** 2))
        uph = ut * Omega
        g = geos['E'] / -(-geos['E'] * ut + geos['E'] * geos['lam'] * uph)
    if fillna is not None and fillna is not False:
        g = np.where(np.isnan(g), fillna, g)
    return g
