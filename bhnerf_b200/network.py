"""Host-side mirror of the reference's ``bhnerf/network.py`` for the render/train hot path.

Same names, argument order and error behaviour as the reference (file:line cited per function);
arrays are numpy / torch instead of jax, and every compute step is a C-ABI call into
libbhnerf_b200.so through :mod:`bhnerf_b200.engine`.  Only the default architecture the kernels
are specialised for is accepted (net_depth=4, net_width=128, posenc_deg=3, relu, out_channel=1,
do_skip=True: bhnerf/network.py:147-157) -- anything else raises, it does not fall back."""
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import constants, engine, utils
from ._lib import N_PARAMS

LAYER_SHAPES = [(21, 128), (128, 128), (128, 128), (149, 128), (128, 1)]


# ------------------------------------------------------------------------------------------------
# parameter pytree <-> flat buffer (flax names params['MLP_0']['Dense_i']['kernel'|'bias'])
# ------------------------------------------------------------------------------------------------
def flatten_params(params):
    d = params['MLP_0'] if 'MLP_0' in params else params
    out = []
    for i, (fi, fo) in enumerate(LAYER_SHAPES):
        k = np.asarray(d['Dense_%d' % i]['kernel'], dtype=np.float32)
        b = np.asarray(d['Dense_%d' % i]['bias'], dtype=np.float32)
        if k.shape != (fi, fo) or b.shape != (fo,):
            raise ValueError('Dense_%d has shape %s/%s, expected %s/%s' % (i, k.shape, b.shape, (fi, fo), (fo,)))
        out += [k.reshape(-1), b.reshape(-1)]
    return np.concatenate(out)


def unflatten_params(flat):
    flat = flat.detach().cpu().numpy() if isinstance(flat, torch.Tensor) else np.asarray(flat)
    d, o = OrderedDict(), 0
    for i, (fi, fo) in enumerate(LAYER_SHAPES):
        k = flat[o:o + fi * fo].reshape(fi, fo).copy(); o += fi * fo
        b = flat[o:o + fo].copy(); o += fo
        d['Dense_%d' % i] = {'kernel': k, 'bias': b}
    return {'MLP_0': d}


class TrainState:
    """Stand-in for flax ``TrainState`` (bhnerf/network.py:182): flat device buffers for params and the
    Adam moments, ``step`` counter, and the schedule hyper-parameters of ``init_state``."""

    def __init__(self, predictor, flat_params, num_iters, lr_init, lr_final, device):
        self.predictor = predictor
        self.apply_fn = predictor.apply
        self.flat = torch.as_tensor(np.asarray(flat_params, dtype=np.float32), device=device).clone()
        self.mu = torch.zeros_like(self.flat)
        self.nu = torch.zeros_like(self.flat)
        self.step = 0
        self.num_iters, self.lr_init, self.lr_final = int(num_iters), float(lr_init), float(lr_final)

    @property
    def params(self):
        return self.predictor._unflatten(self.flat)

    def apply_gradients(self, grads, grad_scale=1.0, guard=None):
        """optax.adam + polynomial_schedule(lr_init, lr_final, 1, num_iters) (network.py:173-174,:621).  `guard`: tcgen05
        workspace whose health flags veto the update (engine.step_guard)."""
        engine.adam_step(self.flat, grads, self.mu, self.nu, self.step, self.lr_init, self.lr_final,
                         self.num_iters, grad_scale=grad_scale, guard=guard)
        self.step += 1
        return self

    def state_dict(self):
        return {'step': self.step, 'params': self.params, 'mu': self.mu.cpu().numpy(), 'nu': self.nu.cpu().numpy()}


class NeRF_Predictor:
    """bhnerf/network.py:124-252."""

    def __init__(self, scale=1.0, rmin=0.0, rmax=np.inf, z_width=np.inf, posenc_deg=3, posenc_var=2e-5,
                 net_depth=4, net_width=128, activation='relu', out_channel=1, do_skip=True):
        if (posenc_deg, net_depth, net_width, out_channel, do_skip) != (3, 4, 128, 1, True) or \
                activation not in ('relu', None) and getattr(activation, '__name__', '') != 'relu':
            raise NotImplementedError('bhnerf_b200 kernels are specialised for the reference defaults '
                                      '(posenc_deg=3, net_depth=4, net_width=128, relu, out_channel=1, do_skip=True)')
        self.scale, self.rmin, self.rmax, self.z_width = float(scale), float(rmin), float(rmax), float(z_width)
        self.posenc_deg, self.posenc_var, self.net_depth, self.net_width = posenc_deg, posenc_var, net_depth, net_width
        self.out_channel, self.do_skip = out_channel, do_skip

    def init_params(self, raytracing_args=None, seed=1):
        """he_uniform kernels / zero biases (network.py:49-50, :159-169).  numpy Generator, not threefry."""
        rng = np.random.default_rng(seed)
        d = OrderedDict()
        for i, (fi, fo) in enumerate(LAYER_SHAPES):
            lim = math.sqrt(6.0 / fi)
            d['Dense_%d' % i] = {'kernel': rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32),
                                 'bias': np.zeros((fo,), dtype=np.float32)}
        return {'MLP_0': d}

    def init_state(self, params, num_iters=5000, lr_init=1e-4, lr_final=1e-6, lr_inject=None, checkpoint_dir='',
                   device=None):
        """network.py:171-189.  ``lr_inject`` is accepted and changes nothing, exactly as in the reference: it adds an
        ``optax.masked`` Adam for leaves named 't_injection' (:176-180), and no such leaf exists -- the learnable injection
        time is commented out (:235) -- so every parameter still gets the lr_init -> lr_final schedule."""
        device = torch.device(device if device is not None else 'cuda')
        state = TrainState(self, flatten_params(params), num_iters, lr_init, lr_final, device)
        if checkpoint_dir:
            from .optimization import restore_checkpoint
            restore_checkpoint(checkpoint_dir, state)
        return state

    def domain(self):
        return dict(scale=self.scale, rmin=self.rmin, rmax=self.rmax, z_width=self.z_width)

    # ---- what the step functions need from a predictor: flat parameters <-> pytree, render forward / pull-back ----
    _unflatten = staticmethod(lambda flat: unflatten_params(flat))
    _flatten = staticmethod(lambda params: flatten_params(params))

    def _render_fwd(self, scene, flat, tf, impl=None, save_acts=False):
        return engine.render_fwd(scene, flat, tf, impl, save_acts=save_acts)

    def _render_bwd(self, scene, flat, tf, d_images, e, acts, impl=None):
        return engine.render_bwd(scene, flat, tf, d_images, e, acts, impl)

    def apply(self, variables, t_frames, t_units, coords, Omega, t_start_obs, t_geos, t_injection, impl=None):
        """NeRF_Predictor.__call__ (network.py:191-237): emission on the given points, shape (Bt, *coords.shape[1:])
        (no Bt axis for a scalar t_frames).  Runs the forward kernel with unit ray weights."""
        params = variables['params'] if 'params' in variables else variables
        flat = params if isinstance(params, torch.Tensor) else flatten_params(params)
        coords = np.asarray(coords, dtype=np.float32)
        pts_shape = coords.shape[1:]
        c2 = coords.reshape(3, -1, 1) if coords.ndim == 2 else coords.reshape(3, -1, coords.shape[-1])
        P, G = c2.shape[1], c2.shape[2]
        ones = np.ones((P, G), dtype=np.float32)
        Om = np.broadcast_to(np.asarray(Omega, dtype=np.float32), pts_shape).reshape(P, G)
        tg = np.broadcast_to(np.asarray(t_geos, dtype=np.float32), pts_shape).reshape(P, G)
        t_units = t_units if t_units is not None else None
        GM_c3 = constants.GM_c3(t_units=t_units) if t_units is not None else 1.0
        # domain fill is applied exactly as in the reference; samples it zeroes are simply not evaluated
        scene = engine.PackedScene(c2, Om, 1.0, ones, ones, ones, tg, utils.time_value(t_start_obs, t_units or 'hr'),
                                   float(t_injection), self.scale, self.rmin, self.rmax, self.z_width, GM_c3)
        tf = np.atleast_1d(utils.time_value(t_frames, t_units or 'hr')).astype(np.float32)
        _, e, _ = engine.render_fwd(scene, flat, tf, impl)
        # scatter the compacted emission back to the dense point set (dead samples are exactly 0)
        dense = torch.zeros((len(tf), P * G), dtype=torch.float32, device=scene.device)
        if scene.n_active:
            dense[:, scene.dense_index] = e[:, :scene.n_active]
        dense = dense.cpu().numpy()
        dense = dense.reshape((len(tf),) + tuple(pts_shape))
        return dense[0] if np.ndim(t_frames) == 0 else dense

    def save_params(self, directory, filename='NeRF_Predictor_params.yml'):
        """network.py:239-247."""
        import yaml
        os.makedirs(directory, exist_ok=True)
        keys = ['scale', 'rmin', 'rmax', 'z_width', 'posenc_deg', 'posenc_var', 'net_depth', 'net_width',
                'out_channel', 'do_skip']
        with open(os.path.join(directory, filename), 'w') as f:
            yaml.dump({k: getattr(self, k) for k in keys}, f)

    @classmethod
    def from_yml(cls, directory, filename='NeRF_Predictor_params.yml'):
        """network.py:249-252."""
        import yaml
        with open(os.path.join(directory, filename)) as f:
            return cls(**yaml.safe_load(f))


class GRID_Predictor:
    """bhnerf/network.py:254-370: a learnable grid_res^3 voxel grid instead of the MLP.  Same warp, domain fill and
    injection mask; lookup = jax.scipy.ndimage.map_coordinates(order=1, cval=0) at (coords+scale)/(2 scale)*(res-1),
    emission = sigmoid(v - 10).  Kernels: csrc/grid.cu (bhnerf_grid_render_fwd / _bwd, mode 1)."""

    def __init__(self, scale=1.0, rmin=0.0, rmax=np.inf, z_width=np.inf, grid_res=64):
        self.scale, self.rmin, self.rmax, self.z_width = float(scale), float(rmin), float(rmax), float(z_width)
        self.grid_res = int(grid_res)

    def init_params(self, raytracing_args=None, seed=1):
        """network.py:305: grid initialised to -10 (emission sigmoid(-20) ~ 2e-9)."""
        return {'grid': np.full((self.grid_res,) * 3, -10.0, dtype=np.float32)}

    def init_state(self, params, num_iters=5000, lr_init=1e-4, lr_final=1e-6, lr_inject=None, checkpoint_dir='',
                   device=None):
        """network.py:287-303 (optax.adam + polynomial_schedule on the grid).  ``lr_inject``: accepted, a no-op as in the
        reference (its masked Adam matches leaves named 't_injection'; the grid params have none, :292-296)."""
        device = torch.device(device if device is not None else 'cuda')
        state = TrainState(self, self._flatten(params), num_iters, lr_init, lr_final, device)
        if checkpoint_dir:
            from .optimization import restore_checkpoint
            restore_checkpoint(checkpoint_dir, state)
        return state

    def domain(self):
        return dict(scale=self.scale, rmin=self.rmin, rmax=self.rmax, z_width=self.z_width)

    def _unflatten(self, flat):
        g = flat.detach().cpu().numpy() if isinstance(flat, torch.Tensor) else np.asarray(flat)
        return {'grid': g.reshape((self.grid_res,) * 3).copy()}

    def _flatten(self, params):
        g = params['grid'] if isinstance(params, dict) else params
        g = g.detach().cpu().numpy() if isinstance(g, torch.Tensor) else np.asarray(g, dtype=np.float32)
        assert g.size == self.grid_res ** 3, 'grid must be grid_res^3'
        return np.ascontiguousarray(g, dtype=np.float32).reshape(-1)

    def _grid(self, flat, device):
        t = flat if isinstance(flat, torch.Tensor) else torch.as_tensor(self._flatten(flat), device=device)
        return t.to(device=device, dtype=torch.float32).reshape((self.grid_res,) * 3)

    def _render_fwd(self, scene, flat, tf, impl=None, save_acts=False):
        images, e = engine.grid_render_fwd(scene, self._grid(flat, scene.device), 2.0 * self.scale, tf,
                                           engine.GRID_MODE_PREDICTOR, want_e=save_acts)
        return images, e, None

    def _render_bwd(self, scene, flat, tf, d_images, e=None, acts=None, impl=None):
        return engine.grid_render_bwd(scene, self._grid(flat, scene.device), 2.0 * self.scale, tf, d_images).reshape(-1)

    def apply(self, variables, t_frames, t_units, coords, Omega, t_start_obs, t_geos, t_injection, impl=None):
        """GRID_Predictor.__call__ (network.py:306-357): emission on the given points, (Bt, *coords.shape[1:])."""
        params = variables['params'] if 'params' in variables else variables
        coords = np.asarray(coords, dtype=np.float32)
        pts_shape = coords.shape[1:]
        c2 = coords.reshape(3, -1, 1) if coords.ndim == 2 else coords.reshape(3, -1, coords.shape[-1])
        P, G = c2.shape[1], c2.shape[2]
        ones = np.ones((P, G), dtype=np.float32)
        Om = np.broadcast_to(np.asarray(Omega, dtype=np.float32), pts_shape).reshape(P, G)
        tg = np.broadcast_to(np.asarray(t_geos, dtype=np.float32), pts_shape).reshape(P, G)
        GM_c3 = constants.GM_c3(t_units=t_units) if t_units is not None else 1.0
        scene = engine.PackedScene(c2, Om, 1.0, ones, ones, ones, tg, utils.time_value(t_start_obs, t_units or 'hr'),
                                   float(t_injection), self.scale, self.rmin, self.rmax, self.z_width, GM_c3)
        tf = np.atleast_1d(utils.time_value(t_frames, t_units or 'hr')).astype(np.float32)
        _, e, _ = self._render_fwd(scene, params if isinstance(params, torch.Tensor) else self._flatten(params), tf,
                                   save_acts=True)
        dense = torch.zeros((len(tf), P * G), dtype=torch.float32, device=scene.device)
        if scene.n_active:
            dense[:, scene.dense_index] = e[:, :scene.n_active]
        dense = dense.cpu().numpy().reshape((len(tf),) + tuple(pts_shape))
        return dense[0] if np.ndim(t_frames) == 0 else dense

    def save_params(self, directory, filename='GRID_Predictor_params.yml'):
        """network.py:359-366 (the reference's key list names NeRF fields; only the ones this class has are written)."""
        import yaml
        os.makedirs(directory, exist_ok=True)
        with open(os.path.join(directory, filename), 'w') as f:
            yaml.dump({k: getattr(self, k) for k in ('scale', 'rmin', 'rmax', 'z_width')}, f)

    @classmethod
    def from_yml(cls, directory, filename='GRID_Predictor_params.yml'):
        """network.py:368-370."""
        import yaml
        with open(os.path.join(directory, filename)) as f:
            return cls(**yaml.safe_load(f))


# ------------------------------------------------------------------------------------------------
# scene cache: the 9 frame-independent raytracing args are prepacked once per (arrays, predictor)
# ------------------------------------------------------------------------------------------------
_scene_cache = OrderedDict()


SCENE_CACHE_SIZE = int(os.environ.get('BHNERF_SCENE_CACHE', '64'))   # >= the number of sub-pixel raytracing_args sets


def _ray_block(a, ray_shard, lead):
    """Rays [r*P/R, (r+1)*P/R) of an array whose axes after the first `lead` are (A, B, G) (ray-major flattening of A x B)."""
    rank, world = ray_shard
    a = np.asarray(a)
    shp = a.shape
    P = int(np.prod(shp[lead:-1]))
    if P % world:
        raise ValueError('%d rays are not divisible by %d ranks' % (P, world))
    per = P // world
    flat = a.reshape(shp[:lead] + (P, shp[-1]))
    return np.ascontiguousarray(flat[..., rank * per:(rank + 1) * per, :]).reshape(shp[:lead] + (per, 1, shp[-1]))


def _scene_for(predictor, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units, device=None,
               ray_shard=None):
    """Prepacked scene of the raytracing args, cached on the identity of the arrays (they are kept alive by the cache; an
    in-place mutation of a cached array is NOT seen -- pass a fresh array).  ray_shard = (rank, world): the scene of this
    rank's contiguous block of rays only (image_shape (P/world, 1))."""
    tsv = float(utils.time_value(t_start_obs, t_units))
    key = (id(coords), id(Omega), id(J) if not np.isscalar(J) else ('scalar', float(J)), id(g), id(dtau), id(Sigma),
           id(t_geos), tsv, float(t_injection), predictor.scale, predictor.rmin, predictor.rmax, predictor.z_width,
           str(t_units), str(device), ray_shard)
    hit = _scene_cache.get(key)
    if hit is not None:
        _scene_cache.move_to_end(key)
        return hit[0]
    if ray_shard is None:
        parts = (coords, Omega, J, g, dtau, Sigma, t_geos)
    else:
        parts = (_ray_block(coords, ray_shard, 1), _ray_block(Omega, ray_shard, 0),
                 J if np.isscalar(J) else _ray_block(J, ray_shard, 1), _ray_block(g, ray_shard, 0),
                 _ray_block(dtau, ray_shard, 0), _ray_block(Sigma, ray_shard, 0), _ray_block(t_geos, ray_shard, 0))
    c_, Om_, J_, g_, dt_, Sg_, tg_ = parts
    scene = engine.PackedScene(c_, Om_, J_, g_, dt_, Sg_, tg_, tsv, float(t_injection), predictor.scale,
                               predictor.rmin, predictor.rmax, predictor.z_width, constants.GM_c3(t_units=t_units),
                               device=device)
    if ray_shard is not None:
        scene.full_image_shape = tuple(np.shape(Omega)[:-1])
    # keep the source arrays alive so the id()-based key stays valid
    _scene_cache[key] = (scene, (coords, Omega, J, g, dtau, Sigma, t_geos))
    while len(_scene_cache) > SCENE_CACHE_SIZE:
        _scene_cache.popitem(last=False)
    return scene


def _predictor_of(predictor_fn):
    p = getattr(predictor_fn, '__self__', predictor_fn)
    if not isinstance(p, (NeRF_Predictor, GRID_Predictor)):
        raise TypeError('predictor_fn must be NeRF_Predictor.apply or GRID_Predictor.apply')
    return p


def _flat(params, device, pred=None):
    if isinstance(params, torch.Tensor):
        return params
    return torch.as_tensor((pred._flatten if pred is not None else flatten_params)(params), device=device)


def _shape_images(images, scene, J):
    """(Bt,S,P) -> (Bt,A,B) for scalar J, (Bt,S,A,B) otherwise, then jnp.squeeze semantics of
    network.py:418 (size-1 axes dropped when J is an array)."""
    Bt = images.shape[0]
    if not scene.polarized:
        return images.reshape((Bt,) + scene.image_shape)
    out = images.reshape((Bt, scene.S) + scene.image_shape)
    return out.squeeze() if (Bt == 1 or scene.S == 1) else out


def image_plane_prediction(params, predictor_fn, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos,
                           t_injection, t_units, impl=None):
    """bhnerf/network.py:373-420.  Returns a device tensor of images."""
    pred = _predictor_of(predictor_fn)
    scene = _scene_for(pred, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units)
    tf = np.atleast_1d(utils.time_value(t_frames, t_units)).astype(np.float32) if not isinstance(t_frames, torch.Tensor) else t_frames
    images, _, _ = pred._render_fwd(scene, _flat(params, scene.device, pred), tf, impl)
    return _shape_images(images, scene, J)


def _image_targets(scene, target, sigma, offset, dtype, Bt):
    if dtype not in ('full', 'lc'):
        raise AttributeError('image dtype ({}) not supported'.format(dtype))
    return [engine._dev_f32(a, scene.device).reshape(Bt, -1) for a in (target, sigma, offset)]


def loss_fn_image(params, predictor_fn, target, sigma, offset, t_frames, coords, Omega, J, g, dtau, Sigma,
                  t_start_obs, t_geos, t_injection, scale, t_units, dtype, impl=None):
    """bhnerf/network.py:422-484.  Returns (scale*loss, [images])."""
    pred = _predictor_of(predictor_fn)
    scene = _scene_for(pred, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    tgt, sig, off = _image_targets(scene, target, sigma, offset, dtype, tf.numel())
    images, _, _ = pred._render_fwd(scene, _flat(params, scene.device, pred), tf, impl)
    loss, _ = engine.loss_image(images, tgt, sig, off, float(scale), dtype)
    return loss, [_shape_images(images, scene, J)]


class SeparableDFT(object):
    """Opt-in stand-in for the explicit matrix ``A`` of loss_fn_eht (network.py:542-544) when it is the plain Fourier kernel
    of the pixel grid, ``A[..., k, (i,j)] = pulse_k exp(-2 pi i (u_k x_i + v_k y_j))`` -- what ehtim's chisqdata_* build.
    Holds the baselines instead of the matrix: ``uv`` with the shape of ``A`` minus the pixel axis plus a trailing 2
    ((nt, [npol,] [3,] nvis, 2), cycles per unit of x / y), the pixel grid ``x_i = x0 + i dx`` (alpha axis), ``y_j = y0 + j dy``
    and an optional complex ``pulse`` factor per visibility.  The eht steps then run the separable tensor-core head
    (bhnerf_vis_dft_fwd / _bwd) and never touch anything of size nvis x npix."""

    def __init__(self, uv, image_shape, x0, dx, y0, dy, pulse=None):
        self.uv = np.ascontiguousarray(np.asarray(uv, dtype=np.float32))
        assert self.uv.shape[-1] == 2
        self.image_shape = tuple(int(v) for v in image_shape)
        self.grid = (float(x0), float(dx), float(y0), float(dy))
        self.pulse = None if pulse is None else np.ascontiguousarray(np.asarray(pulse, dtype=np.complex64))
        assert self.pulse is None or self.pulse.shape == self.uv.shape[:-1]

    @classmethod
    def from_fov(cls, uv, image_shape, fov, pulse=None):
        """Pixel centres (arange(n) - n/2) * fov/n on both axes (the convention of the synthetic configs)."""
        NA, NB = image_shape
        return cls(uv, image_shape, -0.5 * fov, fov / NA, -0.5 * fov, fov / NB, pulse)

    @property
    def shape(self):                          # the shape the explicit A would have
        return self.uv.shape[:-1] + (self.image_shape[0] * self.image_shape[1],)

    def dim(self):
        return len(self.shape)

    def __getitem__(self, key):               # frame indexing only (TemporalBatchedArgs, chunking)
        if isinstance(key, tuple):
            assert all(k is Ellipsis for k in key[1:]), 'a SeparableDFT is indexed along its frame axis only'
            key = key[0]
        return SeparableDFT(self.uv[key], self.image_shape, *self.grid,
                            pulse=None if self.pulse is None else self.pulse[key])

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        assert shape[-1] == self.shape[-1]
        return SeparableDFT(self.uv.reshape(tuple(shape[:-1]) + (2,)), self.image_shape, *self.grid,
                            pulse=None if self.pulse is None else self.pulse.reshape(tuple(shape[:-1])))

    def contiguous(self):
        return self


def _vis_head(A, images, target, sigma, scale, dtype, want_grad=True):
    """(loss[1], vis, d_images | None) of the eht head for an explicit matrix (bhnerf_vis_head) or a SeparableDFT."""
    if not isinstance(A, SeparableDFT):
        return engine.vis_head(A, images, target, sigma, scale, dtype, want_grad=want_grad)
    n = A.shape[0]
    NA, NB = A.image_shape
    vis = engine.vis_dft_fwd(A.uv, images.reshape(n, NA, NB), A.grid, pulse=A.pulse)
    loss, dvis = engine.loss_vis(vis, target, sigma, scale, dtype)
    dI = engine.vis_dft_bwd(A.uv, dvis, A.grid, NA, NB, pulse=A.pulse).reshape(n, 1, NA * NB) if want_grad else None
    return loss, vis, dI


def _eht_prepare(scene, target, sigma, A, dtype, Bt, keep_host=False):
    """Shape checks of loss_fn_eht (network.py:542-559) and the flattening the C ABI takes.  The reference multiplies
    ``A (nt, [npol,] nvis, npix)`` with the image vectors ``(nt, [npol,] npix, 1)`` -- one DFT matrix per frame AND
    polarization (optimization.py:235-251 stacks them on axis 1) -- so the pol axis folds into the frame axis:
    returns A as (nt*S, rows, npix) with rows = nvis ('vis','amp') or 3*ncphase ('cphase')."""
    if dtype not in ('vis', 'amp', 'cphase'):
        raise AttributeError('eht dtype ({}) not supported'.format(dtype))
    if isinstance(A, SeparableDFT):
        if tuple(A.image_shape) != tuple(scene.image_shape):
            raise AttributeError('SeparableDFT image_shape {} does not match the rays {}'.format(A.image_shape, scene.image_shape))
    elif keep_host and not (isinstance(A, torch.Tensor) and A.is_cuda):
        # stays on the host (zero-copy view): gradient_step_eht uploads it chunk by chunk UNDER the render of the chunk
        A = A if isinstance(A, torch.Tensor) else torch.as_tensor(np.asarray(A))
        if A.dtype != torch.complex64:
            A = A.to(torch.complex64)
    else:
        A = engine._c64(A, scene.device)
    tshape = tuple(np.shape(target)) if not isinstance(target, torch.Tensor) else tuple(target.shape)
    S = scene.S
    pol = scene.polarized and not (S == 1 and A.dim() == (4 if dtype == 'cphase' else 3))
    lead = (Bt, S) if pol else (Bt,)
    want_ndim = len(lead) + (3 if dtype == 'cphase' else 2)
    if not scene.polarized and S != 1:
        raise AttributeError('images have {} Stokes channels but A has no polarization axis'.format(S))
    if A.dim() != want_ndim or tuple(A.shape[:len(lead)]) != lead or A.shape[-1] != scene.P or \
            (dtype == 'cphase' and A.shape[-3] != 3):
        raise AttributeError('A should have shape {} = {}, got {}'.format(
            '(nt, [npol,] 3, ncphase, npix)' if dtype == 'cphase' else '(nt, [npol,] nvis, npix)',
            lead + ((3, 'V', scene.P) if dtype == 'cphase' else ('V', scene.P)), tuple(A.shape)))
    vis_ndim = len(lead) + (2 if dtype == 'cphase' else 1)
    if dtype == 'cphase':
        if len(tshape) + 1 != vis_ndim:
            raise AttributeError('visibilities (ndim={}) should have +1 dimensions as target (ndim={}) for dtype={}'.format(
                vis_ndim, len(tshape), dtype))
        return A.reshape(Bt * (S if pol else 1), 3 * A.shape[-2], scene.P)      # triangle axis folded into the row axis
    if len(tshape) != vis_ndim:
        raise AttributeError('visibilities (ndim={}) should have same dimensions as target (ndim={}) for dtype={}'.format(
            vis_ndim, len(tshape), dtype))
    return A.reshape(Bt * (S if pol else 1), A.shape[-2], scene.P)


def _eht_sigma(sigma, target, device):
    """sigma broadcast to the target's shape (the reference divides by it with numpy broadcasting)."""
    sig = engine._dev_f32(sigma, device)
    return sig.expand(target.shape).contiguous() if sig.shape != target.shape else sig


def _eht_rows(x, n):
    """target / sigma of an eht loss as (n, -1) rows (n = frames x polarizations)."""
    return x.reshape(n, -1)


def loss_fn_eht(params, predictor_fn, target, sigma, A, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs,
                t_geos, t_injection, scale, t_units, dtype, impl=None):
    """bhnerf/network.py:486-564 ('vis', 'amp', 'cphase')."""
    pred = _predictor_of(predictor_fn)
    scene = _scene_for(pred, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    A = _eht_prepare(scene, target, sigma, A, dtype, tf.numel())
    images, _, _ = pred._render_fwd(scene, _flat(params, scene.device, pred), tf, impl)
    n = A.shape[0]                                           # frames x polarizations
    tgt = engine._c64(target, scene.device) if dtype == 'vis' else engine._dev_f32(target, scene.device)
    loss, _, _ = _vis_head(A, images.reshape(n, 1, scene.P), _eht_rows(tgt, n),
                           _eht_rows(_eht_sigma(sigma, tgt, scene.device), n), float(scale), dtype, want_grad=False)
    return loss, [_shape_images(images, scene, J)]


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _pmean_and_apply(state, grads, update=True, guard=None, reduce='mean'):
    """jax.lax.pmean(grads,'batch') + state.apply_gradients (network.py:620-621).  On the GPU the exchange is the C ABI's
    bhnerf_allreduce_mean (NCCL on the kernel stream); `reduce='sum'` is the ray-sharded step, whose per-rank gradients are
    partial sums of ONE device's gradient.  (CPU tensors -- the gloo tests of the host logic -- go through torch.distributed.)"""
    dist = _dist()
    scale = 1.0
    if dist is not None:
        if grads.is_cuda:
            (engine.allreduce_mean if reduce == 'mean' else engine.allreduce_sum)(grads)
        else:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM)
            scale = 1.0 / dist.get_world_size() if reduce == 'mean' else 1.0
    if update:
        if guard is not None:
            state.apply_gradients(grads, grad_scale=scale, guard=guard)
        else:
            state.apply_gradients(grads, grad_scale=scale)
    return state


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of the whole train step.  The reference's training loops run tiny steps (batchsize 6 of a 64x64
# image plane: ~0.3 ms of GPU work) tens of thousands of times; host dispatch of the ~10 launches of a step costs more
# than the step.  Every C-ABI call only enqueues on the stream and the Adam step counter lives in device memory
# (bhnerf_adam_step_dev), so the sequence [train_step_image -> adam] is captured once per (state, scene, batch shape)
# and replayed; the per-step inputs travel in ONE pinned-host -> device copy.  BHNERF_CUDA_GRAPHS=0 disables it.
# ------------------------------------------------------------------------------------------------
_USE_GRAPHS = os.environ.get('BHNERF_CUDA_GRAPHS', '1') != '0'
_graph_cache = OrderedDict()


class _GraphedImageStep:
    N_STAGE = 4          # pinned staging buffers in rotation (a buffer is rewritten only after its copy completed)

    def __init__(self, state, scene, Bt, kind, scale, impl):
        dev = scene.device
        self.Bt, self.kind = Bt, kind
        self.n_t = Bt * scene.S * scene.P if kind == 'full' else Bt * scene.S
        total = Bt + 3 * self.n_t
        self.stage = [torch.empty(total, dtype=torch.float32).pin_memory() for _ in range(self.N_STAGE)]
        self.stage_np = [t.numpy() for t in self.stage]
        self.stage_ev = [None] * self.N_STAGE
        self.turn = 0
        self.dev = torch.zeros(total, dtype=torch.float32, device=dev)
        n = self.n_t
        self.tf, self.tgt, self.sig, self.off = (self.dev[:Bt], self.dev[Bt:Bt + n], self.dev[Bt + n:Bt + 2 * n],
                                                 self.dev[Bt + 2 * n:])
        self.sig.fill_(1.0)
        self.out = (torch.empty(1, dtype=torch.float32, device=dev),
                    torch.empty((Bt, scene.S, scene.P), dtype=torch.float32, device=dev),
                    torch.empty(engine.N_PARAMS, dtype=torch.float32, device=dev))
        self.count = torch.full((1,), int(state.step), dtype=torch.int32, device=dev)
        self.synced_step = int(state.step)
        self.state, self.scene = state, scene          # keep the captured buffers alive
        lib = engine._lib.load()

        if _dist() is not None:
            engine.comm(dev)                           # collective: created before the capture, on every rank

        def body():
            engine.train_step_image(scene, state.flat, self.tf, self.tgt, self.sig, self.off, float(scale), kind, impl,
                                    out=self.out)
            engine.allreduce_mean(self.out[2])         # jax.lax.pmean (no-op with one rank); NCCL is graph-capturable
            engine.check(lib.bhnerf_adam_step_dev(engine._ptr(state.flat), engine._ptr(self.out[2]), engine._ptr(state.mu),
                                                  engine._ptr(state.nu), state.flat.numel(), engine._ptr(self.count),
                                                  state.lr_init, state.lr_final, state.num_iters, 0.9, 0.999, 1e-8, 1.0,
                                                  engine._ptr(engine.step_guard(dev, impl)), engine._stream()))
        # warm-up outside the capture (lazy module load, function attributes, workspace growth), on saved copies so that
        # it leaves the optimiser state untouched
        keep = [t.clone() for t in (state.flat, state.mu, state.nu, self.count)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        for t, k in zip((state.flat, state.mu, state.nu, self.count), keep):
            t.copy_(k)
        self.ws = engine._workspaces.get(dev.index if dev.index is not None else torch.cuda.current_device())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()

    def _fill(self, dst_np, dst_dev, src):
        """numpy -> pinned staging (returns True), device tensor -> device slot directly (returns False)."""
        if isinstance(src, torch.Tensor):
            dst_dev.copy_(src.reshape(-1), non_blocking=True)
            return False
        np.copyto(dst_np, np.asarray(src, dtype=np.float32).reshape(-1))
        return True

    def run(self, state, tf, tgt, sig, off):
        k = self.turn = (self.turn + 1) % self.N_STAGE
        if self.stage_ev[k] is not None:
            self.stage_ev[k].synchronize()
        h, Bt, n = self.stage_np[k], self.Bt, self.n_t
        host = [self._fill(h[:Bt], self.tf, tf), self._fill(h[Bt:Bt + n], self.tgt, tgt),
                self._fill(h[Bt + n:Bt + 2 * n], self.sig, sig), self._fill(h[Bt + 2 * n:], self.off, off)]
        if all(host):
            self.dev.copy_(self.stage[k], non_blocking=True)
        else:
            for ok, (a, b) in zip(host, ((0, Bt), (Bt, Bt + n), (Bt + n, Bt + 2 * n), (Bt + 2 * n, Bt + 3 * n))):
                if ok:
                    self.dev[a:b].copy_(self.stage[k][a:b], non_blocking=True)
        if any(host):
            ev = self.stage_ev[k] = self.stage_ev[k] or torch.cuda.Event()
            ev.record()
        if int(state.step) != self.synced_step:           # restored checkpoint / external update of the counter
            self.count.fill_(int(state.step))
        self.graph.replay()
        state.step += 1
        self.synced_step = int(state.step)
        return self.out[0].clone(), self.out[1].clone()


def _graphed_image_step(state, scene, Bt, kind, scale, impl):
    key = (id(state), state.flat.data_ptr(), id(scene), Bt, kind, float(scale), engine.resolve_impl(impl))
    hit = _graph_cache.get(key)
    if hit is None:
        hit = _graph_cache[key] = _GraphedImageStep(state, scene, Bt, kind, scale, impl)
        while len(_graph_cache) > 4:
            _graph_cache.popitem(last=False)
    else:
        _graph_cache.move_to_end(key)
    return hit


def _gather_rays(images_r, scene_r, J):
    """(Bt,S,P/R) of every rank -> full images shaped like image_plane_prediction's (host-side plumbing, off the data path)."""
    import torch.distributed as dist
    outs = [torch.empty_like(images_r) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, images_r.contiguous())
    full = torch.cat(outs, dim=2)
    Bt, S = full.shape[:2]
    shp = scene_r.full_image_shape
    if not scene_r.polarized:
        return full.reshape((Bt,) + shp)
    out = full.reshape((Bt, S) + shp)
    return out.squeeze() if (Bt == 1 or S == 1) else out


def _ray_sharded_image_step(state, t_units, dtype, target, sigma, offset, t_frames, rt_args, scale, impl, ray_shard,
                            update):
    """Ray sharding for batches with fewer frames than ranks -- the reference cannot run those at all
    (optimization.py:39, :360-362).  Every rank renders ALL frames of the batch for its contiguous block of P/world rays;
    'full' needs no exchange before the loss, 'lc' all-reduces the (Bt,S) partial lightcurves BEFORE the loss
    non-linearity; the per-rank gradients are partial sums, so they are all-reduced with SUM: the step equals ONE device's
    step on the whole batch."""
    if dtype not in ('full', 'lc'):
        raise AttributeError('image dtype ({}) not supported'.format(dtype))
    pred = state.predictor
    rank, world = ray_shard
    scene = _scene_for(pred, *rt_args, t_units, device=state.flat.device, ray_shard=ray_shard)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    Bt, S, Pr = tf.numel(), scene.S, scene.P
    images, e, acts = pred._render_fwd(scene, state.flat, tf, impl, save_acts=update)
    if dtype == 'full':
        def own(a):
            a = engine._dev_f32(a, scene.device).reshape(Bt, S, -1)
            return a[:, :, rank * Pr:(rank + 1) * Pr].contiguous()
        loss, dI = engine.loss_image(images, own(target), own(sigma), own(offset), float(scale), 'full')
        engine.allreduce_sum(loss)
    else:
        lc = engine.allreduce_sum(engine.lightcurve(images))
        loss, dI = engine.loss_lightcurve(lc, target, sigma, offset, float(scale), Pr)
    if update:
        grads = pred._render_bwd(scene, state.flat, tf, dI, e, acts, impl)
        guard = engine.step_guard(scene.device, impl) if isinstance(pred, NeRF_Predictor) else None
        state = _pmean_and_apply(state, grads, guard=guard, reduce='sum')
    return loss, state, _gather_rays(images, scene, rt_args[2])


def gradient_step_image(state, t_units, dtype, target, sigma, offset, t_frames, coords, Omega, J, g, dtau, Sigma,
                        t_start_obs, t_geos, t_injection, scale, impl=None, ray_shard=None):
    """bhnerf/network.py:566-622: value_and_grad(loss_fn_image) -> pmean -> apply_gradients.
    Returns (loss, state, images).  ray_shard = (rank, world): see _ray_sharded_image_step."""
    pred = state.predictor
    if ray_shard is not None:
        return _ray_sharded_image_step(state, t_units, dtype, target, sigma, offset, t_frames,
                                       (coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection), scale, impl,
                                       ray_shard, update=True)
    scene = _scene_for(pred, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units,
                       device=state.flat.device)
    if _USE_GRAPHS and isinstance(pred, NeRF_Predictor) and ray_shard is None and dtype in ('full', 'lc'):
        tfh = t_frames if isinstance(t_frames, torch.Tensor) else np.atleast_1d(utils.time_value(t_frames, t_units))
        Bt = int(tfh.numel() if isinstance(tfh, torch.Tensor) else tfh.size)
        step = _graphed_image_step(state, scene, Bt, dtype, scale, impl)
        want = step.n_t
        for a in (target, sigma, offset):
            if int(a.numel() if isinstance(a, torch.Tensor) else np.size(a)) != want:
                raise AssertionError('target shape mismatch')
        loss, images = step.run(state, tfh, target, sigma, offset)
        return loss, state, _shape_images(images, scene, J)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    tgt, sig, off = _image_targets(scene, target, sigma, offset, dtype, tf.numel())
    if isinstance(pred, GRID_Predictor):       # voxel grid: forward -> loss head -> pull-back to the grid (csrc/grid.cu)
        images, _, _ = pred._render_fwd(scene, state.flat, tf)
        loss, dI = engine.loss_image(images, tgt, sig, off, float(scale), dtype)
        grads = pred._render_bwd(scene, state.flat, tf, dI)
    else:
        loss, images, grads = engine.train_step_image(scene, state.flat, tf, tgt, sig, off, float(scale), dtype, impl)
    guard = engine.step_guard(scene.device, impl) if isinstance(pred, NeRF_Predictor) else None
    state = _pmean_and_apply(state, grads, guard=guard)
    return loss, state, _shape_images(images, scene, J)


def test_image(state, t_units, dtype, target, sigma, offset, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs,
               t_geos, t_injection, scale, impl=None, ray_shard=None):
    """bhnerf/network.py:684-739: forward + loss only."""
    if ray_shard is not None:
        return _ray_sharded_image_step(state, t_units, dtype, target, sigma, offset, t_frames,
                                       (coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection), scale, impl,
                                       ray_shard, update=False)
    loss, [images] = loss_fn_image(state.flat, state.predictor.apply, target, sigma, offset, t_frames, coords, Omega,
                                   J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, scale, t_units, dtype, impl)
    return loss, state, images


def _ray_sharded_eht_step(state, t_units, dtype, target, sigma, A, t_frames, rt_args, scale, impl, ray_shard, update):
    """Ray sharding of the eht step: every rank multiplies ITS pixel columns of the per-frame DFT matrices with its block
    of the images; the (Bt,V) partial visibilities are all-reduced (SUM) before the chi^2, the gradients after it."""
    pred = state.predictor
    rank, world = ray_shard
    scene = _scene_for(pred, *rt_args, t_units, device=state.flat.device, ray_shard=ray_shard)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    Bt, Pr = tf.numel(), scene.P
    A = engine._c64(A, scene.device)
    A_r = A[..., rank * Pr:(rank + 1) * Pr].contiguous()                 # this rank's pixel columns
    A_r = _eht_prepare(scene, target, sigma, A_r, dtype, Bt)
    n = A_r.shape[0]
    images, e, acts = pred._render_fwd(scene, state.flat, tf, impl, save_acts=update)
    vis = engine.allreduce_sum(engine.vis_fwd(A_r, images.reshape(n, 1, Pr)))
    tgt = engine._c64(target, scene.device) if dtype == 'vis' else engine._dev_f32(target, scene.device)
    loss, dvis = engine.loss_vis(vis, _eht_rows(tgt, n), _eht_rows(_eht_sigma(sigma, tgt, scene.device), n), float(scale), dtype)
    if update:
        dI = engine.vis_bwd(A_r, dvis, Pr).reshape(images.shape)
        grads = pred._render_bwd(scene, state.flat, tf, dI, e, acts, impl)
        guard = engine.step_guard(scene.device, impl) if isinstance(pred, NeRF_Predictor) else None
        state = _pmean_and_apply(state, grads, guard=guard, reduce='sum')
    return loss, state, _gather_rays(images, scene, rt_args[2])


def gradient_step_eht(state, t_units, dtype, target, sigma, A, t_frames, coords, Omega, J, g, dtau, Sigma,
                      t_start_obs, t_geos, t_injection, scale, impl=None, ray_shard=None):
    """bhnerf/network.py:624-682."""
    pred = state.predictor
    if ray_shard is not None:
        return _ray_sharded_eht_step(state, t_units, dtype, target, sigma, A, t_frames,
                                     (coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection), scale, impl,
                                     ray_shard, update=True)
    scene = _scene_for(pred, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units,
                       device=state.flat.device)
    tf = engine._dev_f32(utils.time_value(t_frames, t_units) if not isinstance(t_frames, torch.Tensor) else t_frames,
                         scene.device).reshape(-1)
    A = _eht_prepare(scene, target, sigma, A, dtype, tf.numel(), keep_host=True)
    A_on_host = isinstance(A, torch.Tensor) and not A.is_cuda
    # the chi^2 is separable per frame: process frame chunks whose saved activations fit the workspace cap
    Bt = tf.numel()
    Bc = Bt if isinstance(pred, GRID_Predictor) else engine.frames_per_chunk(scene, Bt, impl)
    if A_on_host and Bt >= 16:
        # a host-resident A is the long pole (8 V P bytes per frame over PCIe): quarter-batch chunks let the upload of chunk
        # k+1 run under the head and the backward of chunk k
        Bc = min(Bc, max(8, -(-Bt // 4)))
    tgt = engine._c64(target, scene.device) if dtype == 'vis' else engine._dev_f32(target, scene.device)
    sig = _eht_sigma(sigma, tgt, scene.device)
    npol = A.shape[0] // Bt                                   # DFT matrices per frame (1, or one per Stokes channel)
    tgt, sig = _eht_rows(tgt, Bt * npol), _eht_rows(sig, Bt * npol)
    loss, grads, imgs = None, None, []
    starts = list(range(0, Bt, Bc))
    rows_of = lambda b0: slice(b0 * npol, min(b0 + Bc, Bt) * npol)
    pending = {}                                              # chunk start -> (device A, copy event)
    for k, b0 in enumerate(starts):
        sl = slice(b0, min(b0 + Bc, Bt))
        rows = rows_of(b0)
        images, e, acts = pred._render_fwd(scene, state.flat, tf[sl], impl, save_acts=not isinstance(pred, GRID_Predictor))
        if A_on_host:
            # the chunk's DFT matrices (25 MB per frame at the cfg3 shape) travel on a copy stream while the render of the
            # chunk -- enqueued above -- runs, and the NEXT chunk's follow right behind, under this chunk's head + backward
            for b1 in starts[k:k + 2]:
                if b1 not in pending:
                    pending[b1] = engine.upload_async(A[rows_of(b1)], scene.device)
            A_c = engine.wait_upload(*pending.pop(b0))
        else:
            A_c = A[rows].contiguous()
        l, _, dI = _vis_head(A_c, images.reshape(A_c.shape[0], 1, scene.P), tgt[rows], sig[rows], float(scale), dtype)
        dI = dI.reshape(images.shape)
        g = pred._render_bwd(scene, state.flat, tf[sl], dI, e, acts, impl)
        loss = l if loss is None else engine.add_inplace(loss, l)       # per-chunk partials accumulate on the device
        grads = g if grads is None else engine.add_inplace(grads, g)
        imgs.append(images)
    images = imgs[0] if len(imgs) == 1 else torch.cat(imgs, dim=0)
    guard = engine.step_guard(scene.device, impl) if isinstance(pred, NeRF_Predictor) else None
    state = _pmean_and_apply(state, grads, guard=guard)
    return loss, state, _shape_images(images, scene, J)


def test_eht(state, t_units, dtype, target, sigma, A, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs,
             t_geos, t_injection, scale, impl=None, ray_shard=None):
    """bhnerf/network.py:741-795."""
    if ray_shard is not None:
        return _ray_sharded_eht_step(state, t_units, dtype, target, sigma, A, t_frames,
                                     (coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection), scale, impl,
                                     ray_shard, update=False)
    loss, [images] = loss_fn_eht(state.flat, state.predictor.apply, target, sigma, A, t_frames, coords, Omega, J, g,
                                 dtau, Sigma, t_start_obs, t_geos, t_injection, scale, t_units, dtype, impl)
    return loss, state, images


def raytracing_args(geos, Omega, t_injection, t_start_obs, J=1.0):
    """bhnerf/network.py:850-894: ordered dict whose VALUE ORDER is the positional ABI of the step
    functions.  ``geos`` is any mapping/namespace with x,y,z,t,dtau,Sigma and either ``g`` or the
    fields kgeo.doppler_factor needs (r, theta, spin, M, E, lam, Xi)."""
    from . import kgeo
    get = (lambda k: geos[k]) if hasattr(geos, '__getitem__') else (lambda k: getattr(geos, k))
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    coords = f32([get('x'), get('y'), get('z')])
    try:
        gfac = get('g')
    except (KeyError, AttributeError):
        gfac = kgeo.doppler_factor(geos, kgeo.azimuthal_velocity_vector(geos, Omega))
    return OrderedDict({
        'coords': coords, 'Omega': f32(Omega), 'J': J if np.isscalar(J) else f32(J), 'g': f32(gfac),
        'dtau': f32(get('dtau')), 'Sigma': f32(get('Sigma')), 't_start_obs': t_start_obs, 't_geos': f32(get('t')),
        't_injection': t_injection})


def sample_3d_grid(apply_fn, params, t_frame=0, t_start_obs=0, Omega=0, fov=None, coords=None, resolution=64,
                   chunk=-1):
    """bhnerf/network.py:797-840: emission on a res^3 grid with identity warp (t_geos=t_injection=0)."""
    if (coords is None) and (fov is not None):
        grid_1d = np.linspace(-fov / 2, fov / 2, resolution)
        coords = np.array(np.meshgrid(grid_1d, grid_1d, grid_1d, indexing='ij'))
    elif coords is None:
        raise AttributeError('Either coords or fov+resolution must be provided')
    t_units = getattr(t_frame, 'unit', None)
    return apply_fn({'params': params}, t_frame, t_units, coords, Omega, t_start_obs, 0.0, 0.0)


def image_plane_checkpoint(raytracing_args, checkpoint_dir, t, rmin=0.0, rmax=np.inf, batchsize=20):
    """bhnerf/network.py:896-906: load the predictor + latest checkpoint of a run and render the whole movie (forward
    only, chunked).  Returns a numpy array (nt, [S,] A, B)."""
    from . import optimization
    predictor = NeRF_Predictor.from_yml(checkpoint_dir)
    predictor.rmax = min(rmax, predictor.rmax)
    predictor.rmin = max(rmin, predictor.rmin)
    params = predictor.init_params(raytracing_args)
    state = predictor.init_state(params, checkpoint_dir=checkpoint_dir)
    J = np.atleast_1d(raytracing_args)[0]['J']
    num_stokes = 1 if np.isscalar(J) else np.shape(J)[0]
    target = np.zeros((len(t), num_stokes)) if not np.isscalar(J) else np.zeros((len(t), 1))
    train_step = optimization.TrainStep.image(t, target, dtype='lc')
    _, image_plane = optimization.total_movie_loss(batchsize, state, train_step, raytracing_args, return_frames=True)
    return image_plane
