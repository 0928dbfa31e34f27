"""Host-side scene object of the jax.ffi binding (integration/jax_binding.py): the prepack, the typed scalar attributes
the XLA-FFI handlers of integration/xla_ffi_shim.cc take, and the size queries of the C ABI.

This module does NOT import jax: it is the part of the binding that is exercised in this repository (the ctypes prepack
and the C-ABI size queries run on the GPU box; tests/test_abi_host.py checks on the CPU that binding, shim and this module
agree on names).  The packed scene lives in a torch CUDA tensor; under JAX it is handed over zero-copy with
``jax.dlpack.from_dlpack`` (``JaxScene.as_jax``)."""
import numpy as np

from . import _lib, constants, engine, utils

# name -> ctype of every scalar attribute a handler receives: exactly the fields of `bhnerf_scene_t` besides the pointer
# (include/bhnerf_b200.h).  integration/xla_ffi_shim.cc binds them one by one with .Attr<T>("name").
ATTRS = (('n_active', 'int32'), ('n_pad', 'int32'), ('P', 'int32'), ('G', 'int32'), ('S', 'int32'),
         ('t_start_obs', 'float32'), ('GM_c3', 'float32'), ('t_injection', 'float32'), ('scale', 'float32'))


class JaxScene(object):
    """Prepacked scene + what jax.ffi.ffi_call needs around it."""

    def __init__(self, packed_scene):
        self.scene = packed_scene                      # engine.PackedScene (owns the device buffer)
        self.P, self.G, self.S = packed_scene.P, packed_scene.G, packed_scene.S
        self.n_active, self.n_pad = packed_scene.n_active, packed_scene.n_pad
        self.image_shape = packed_scene.image_shape
        self.polarized = packed_scene.polarized

    @property
    def attrs(self):
        """Typed scalar attributes for ffi_call(..., **scene.attrs) (numpy scalars: jax.ffi encodes them as typed attrs)."""
        st = self.scene.struct
        return {name: (np.int32 if ty == 'int32' else np.float32)(getattr(st, name)) for name, ty in ATTRS}

    @property
    def packed(self):
        """The packed device buffer (torch uint8 CUDA tensor)."""
        return self.scene.packed

    def as_jax(self):
        """The packed buffer as a jax array (zero copy); needs jax."""
        import jax.dlpack
        return jax.dlpack.from_dlpack(self.scene.packed)

    def acts_bytes(self, Bt, impl=_lib.IMPL_TC):
        return int(_lib.load().bhnerf_acts_bytes(self.scene.ref, int(Bt), impl))

    @property
    def fwd_workspace_bytes(self):
        return int(_lib.load().bhnerf_fwd_workspace_bytes(_lib.IMPL_TC))

    def bwd_workspace_bytes(self, Bt, impl=_lib.IMPL_TC):
        """Workspace of bhnerf_render_bwd with saved residuals for ALL Bt frames in one chunk."""
        lib = _lib.load()
        one = lib.bhnerf_bwd_workspace_bytes(self.scene.ref, int(Bt), impl)
        fixed = lib.bhnerf_bwd_fixed_workspace_bytes(self.scene.ref, int(Bt), impl)
        return int(one + (one - fixed - 1024) * (int(Bt) - 1))


def prepack(predictor, coords, Omega, J, g, dtau, Sigma, t_start_obs, t_geos, t_injection, t_units, scene_cache=None):
    """bhnerf_prepack of one raytracing_args set (bhnerf/network.py:850-894) for the predictor's recovery domain
    (NeRF_Predictor fields scale, rmin, rmax, z_width, network.py:147-157), cached on the identity of the arrays."""
    key = (id(coords), id(Omega), id(J) if not np.isscalar(J) else float(J), id(g), id(dtau), id(Sigma), id(t_geos),
           float(utils.time_value(t_start_obs, t_units)), float(t_injection), float(predictor.scale), float(predictor.rmin),
           float(predictor.rmax), float(predictor.z_width), str(t_units))
    if scene_cache is not None and key in scene_cache:
        return scene_cache[key][0]
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    ps = engine.PackedScene(f32(coords), f32(Omega), J if np.isscalar(J) else f32(J), f32(g), f32(dtau), f32(Sigma),
                            f32(t_geos), utils.time_value(t_start_obs, t_units), float(t_injection), float(predictor.scale),
                            float(predictor.rmin), float(predictor.rmax), float(predictor.z_width),
                            constants.GM_c3(t_units=t_units))
    scene = JaxScene(ps)
    if scene_cache is not None:
        scene_cache[key] = (scene, (coords, Omega, J, g, dtau, Sigma, t_geos))     # keep the id()-keyed arrays alive
    return scene
