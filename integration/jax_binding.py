"""The binding a bhnerf maintainer adds so that bhnerf/network.py keeps its JAX API while the hot path runs in
libbhnerf_b200.so (jax.ffi custom calls over the C ABI of include/bhnerf_b200.h).

NOT importable in this repository's image: jax / jaxlib are not installed.  What it needs from this
repository -- bhnerf_b200.jax_scene (prepack, typed scene attributes, workspace sizes) and the C entry points -- exists and
is exercised here through ctypes; tests/test_abi_host.py checks on the CPU that this file parses, that every name it takes
from the repository resolves, and that its ffi_call attributes are the ones integration/xla_ffi_shim.cc binds.

Usage inside bhnerf (replaces the body of network.image_plane_prediction, bhnerf/network.py:373-420):

    from integration import jax_binding as b200
    images = b200.image_plane_prediction(params, predictor, t_frames, coords, Omega, J, g, dtau, Sigma,
                                         t_start_obs, t_geos, t_injection, t_units)

loss_fn_image / loss_fn_eht / gradient_step_* stay as they are: jax.value_and_grad differentiates through the
custom_vjp below, jax.lax.pmean and optax.adam are untouched (north star: host code stays Python/JAX).
"""
import ctypes
import functools

import jax
import jax.numpy as jnp
import numpy as np

_lib = ctypes.CDLL('libbhnerf_xla.so')        # integration/xla_ffi_shim.cc linked against libbhnerf_b200.so
jax.ffi.register_ffi_target('bhnerf_render_fwd', jax.ffi.pycapsule(_lib.BhnerfRenderFwd), platform='CUDA')
jax.ffi.register_ffi_target('bhnerf_render_bwd', jax.ffi.pycapsule(_lib.BhnerfRenderBwd), platform='CUDA')

N_PARAMS = 55169


def flatten(params):
    """flax tree params['MLP_0']['Dense_i']['kernel'|'bias'] -> flat [55169] (kernel (in,out) row-major, then bias)."""
    d = params['MLP_0'] if 'MLP_0' in params else params
    return jnp.concatenate([x.reshape(-1) for i in range(5) for x in (d['Dense_%d' % i]['kernel'], d['Dense_%d' % i]['bias'])])


def unflatten_like(flat, params):
    leaves, o = {}, 0
    d = params['MLP_0'] if 'MLP_0' in params else params
    for i in range(5):
        k, b = d['Dense_%d' % i]['kernel'], d['Dense_%d' % i]['bias']
        leaves['Dense_%d' % i] = {'kernel': flat[o:o + k.size].reshape(k.shape), 'bias': flat[o + k.size:o + k.size + b.size]}
        o += k.size + b.size
    return {'MLP_0': leaves} if 'MLP_0' in params else leaves


@functools.partial(jax.custom_vjp, nondiff_argnums=(0,))
def _render(scene, flat_params, t_frames):
    return _render_fwd(scene, flat_params, t_frames)[0]


def _render_fwd(scene, flat_params, t_frames):
    """scene: bhnerf_b200.jax_scene.JaxScene (packed device buffer, typed scalar attributes, workspace sizes)."""
    Bt = t_frames.shape[0]
    out_types = (jax.ShapeDtypeStruct((Bt, scene.S, scene.P), jnp.float32),
                 jax.ShapeDtypeStruct((Bt, scene.n_pad), jnp.float32),
                 jax.ShapeDtypeStruct((scene.acts_bytes(Bt),), jnp.uint8),
                 jax.ShapeDtypeStruct((scene.fwd_workspace_bytes,), jnp.uint8))
    images, e, acts, _ = jax.ffi.ffi_call('bhnerf_render_fwd', out_types)(scene.as_jax(), flat_params, t_frames,
                                                                        **scene.attrs)
    return images, (flat_params, t_frames, e, acts)


def _render_bwd(scene, res, d_images):
    flat_params, t_frames, e, acts = res
    Bt = t_frames.shape[0]
    out_types = (jax.ShapeDtypeStruct((N_PARAMS,), jnp.float32),
                 jax.ShapeDtypeStruct((scene.bwd_workspace_bytes(Bt),), jnp.uint8))
    d_params, _ = jax.ffi.ffi_call('bhnerf_render_bwd', out_types)(scene.as_jax(), flat_params, t_frames, d_images, e, acts,
                                                                  **scene.attrs)
    return d_params, None            # gradient w.r.t. params only (argnums=0, bhnerf/network.py:617)


_render.defvjp(lambda scene, p, t: _render_fwd(scene, p, t), _render_bwd)


def image_plane_prediction(params, predictor, t_frames, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos,
                           t_injection, t_units, scene_cache={}):
    """Same signature/return as bhnerf.network.image_plane_prediction (predictor instead of predictor_fn)."""
    from bhnerf_b200.jax_scene import prepack            # thin ctypes call of bhnerf_prepack, cached per argument set
    scene = prepack(predictor, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units, scene_cache)
    images = _render(scene, flatten(params), jnp.asarray(t_frames, jnp.float32).reshape(-1))
    images = images.reshape((images.shape[0], scene.S) + scene.image_shape)
    return jnp.squeeze(images) if not np.isscalar(J) else images[:, 0]     # network.py:415-419
