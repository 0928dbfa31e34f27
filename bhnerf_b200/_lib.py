"""ctypes binding of the C ABI in include/bhnerf_b200.h (built by __graft_entry__.build()).

There is no CPU fallback and no alternate backend: if the shared library is missing, or a symbol
declared in the header is not exported, loading fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BHNERF_B200_LIB', os.path.join(_HERE, 'lib', 'libbhnerf_b200.so'))   # override: instrumented builds

IMPL_SIMT, IMPL_TC = 0, 1
LOSS_FULL, LOSS_LC, LOSS_VIS, LOSS_AMP, LOSS_CPHASE = 0, 1, 2, 3, 4
LOSS_KINDS = {'full': LOSS_FULL, 'lc': LOSS_LC, 'vis': LOSS_VIS, 'amp': LOSS_AMP, 'cphase': LOSS_CPHASE}
N_PARAMS = 55169


class Scene(C.Structure):
    """struct bhnerf_scene (include/bhnerf_b200.h)."""
    _fields_ = [('packed', C.c_void_p), ('n_active', C.c_int32), ('n_pad', C.c_int32), ('P', C.c_int32),
                ('G', C.c_int32), ('S', C.c_int32), ('t_start_obs', C.c_float), ('GM_c3', C.c_float),
                ('t_injection', C.c_float), ('scale', C.c_float)]


_vp, _i32, _f32, _sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t
_SP = C.POINTER(Scene)

# name -> (restype, argtypes): one entry per function declared in include/bhnerf_b200.h
SIGNATURES = {
    'bhnerf_last_error': (C.c_char_p, []),
    'bhnerf_version': (C.c_int, []),
    'bhnerf_device_check': (C.c_int, [C.POINTER(C.c_int)] * 3),
    'bhnerf_packed_bytes': (_sz, [_i32, _i32, _i32]),
    'bhnerf_prepack': (C.c_int, [_vp] * 7 + [_i32, _i32, _i32, _f32, _f32, _f32, _vp, _sz, _SP, _vp]),
    'bhnerf_acts_bytes': (_sz, [_SP, _i32, _i32]),
    'bhnerf_fwd_workspace_bytes': (_sz, [_i32]),
    'bhnerf_render_fwd': (C.c_int, [_SP, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _i32, _vp]),
    'bhnerf_bwd_workspace_bytes': (_sz, [_SP, _i32, _i32]),
    'bhnerf_bwd_fixed_workspace_bytes': (_sz, [_SP, _i32, _i32]),
    'bhnerf_render_bwd': (C.c_int, [_SP, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _i32, _vp]),
    'bhnerf_loss_image': (C.c_int, [_vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'bhnerf_vis_fwd': (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'bhnerf_loss_vis': (C.c_int, [_vp, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'bhnerf_vis_dft_fwd': (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    'bhnerf_vis_dft_bwd': (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _vp, _vp, _vp]),
    'bhnerf_vis_head': (C.c_int, [_vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    'bhnerf_vis_bwd': (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'bhnerf_train_workspace_bytes': (_sz, [_SP, _i32, _i32]),
    'bhnerf_train_step_image': (C.c_int, [_SP, _vp, _vp, _i32, _vp, _vp, _vp, _f32, _i32, _vp, _vp, _vp, _vp,
                                          _sz, _i32, _vp]),
    'bhnerf_velocity_warp_coords': (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, _i32, _f32, _f32, _f32, _vp, _vp]),
    'bhnerf_fill_unsupervised_emission': (C.c_int, [_vp, _vp, _i32, C.c_int64, _f32, _f32, _f32, _f32, _vp]),
    'bhnerf_radiative_transfer': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'bhnerf_launch_count': (C.c_int64, []),
    'bhnerf_profile_begin': (C.c_int, []),
    'bhnerf_profile_end': (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'bhnerf_grid_render_fwd': (C.c_int, [_SP, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _vp, _i32, _vp, _vp, _vp]),
    'bhnerf_interpolate_coords': (C.c_int, [_vp, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _vp, C.c_int64, _vp, _vp]),
    'bhnerf_grid_render_bwd': (C.c_int, [_SP, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _i32, _vp, _vp, _vp]),
    'bhnerf_geodesic_inputs': (C.c_int, [_vp] * 7 + [C.c_int64, _i32, C.c_double, C.c_double, C.c_double, C.c_double]
                               + [_vp] * 7),
    'bhnerf_polarization_workspace_bytes': (_sz, [C.c_int64, _i32]),
    'bhnerf_polarization_factors': (C.c_int, [_vp] * 8 + [C.c_int64, _i32] + [C.c_double] * 10 + [_i32, _vp, _vp, _sz, _vp]),
    'bhnerf_adam_step_dev': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _f32, _f32, _i32, _f32, _f32, _f32, _f32, _vp, _vp]),
    'bhnerf_workspace_status': (C.c_int, [_vp, C.POINTER(C.c_int32), _vp]),
    'bhnerf_comm_unique_id': (C.c_int, [_vp]),
    'bhnerf_comm_init': (C.c_int, [_i32, _i32, _vp, C.POINTER(C.c_void_p)]),
    'bhnerf_comm_destroy': (C.c_int, [_vp]),
    'bhnerf_allreduce_mean': (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    'bhnerf_allreduce_sum': (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    'bhnerf_lightcurve': (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    'bhnerf_loss_lightcurve': (C.c_int, [_vp, _vp, _vp, _vp, _f32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'bhnerf_add_inplace': (C.c_int, [_vp, _vp, _i32, _vp]),
    'bhnerf_adam_step': (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _i32, _f32, _f32, _f32, _f32,
                                   _vp, _vp]),
}

_lib = None


class BhnerfError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every declared symbol.  Works without a GPU (no compute call)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BhnerfError('bhnerf_b200: %s not found -- run `python __graft_entry__.py` (nvcc, sm_100a). '
                          'There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise BhnerfError('bhnerf_b200: symbol %s missing from %s' % (name, LIB_PATH)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise BhnerfError('bhnerf_b200 C ABI error %d: %s' % (rc, load().bhnerf_last_error().decode()))


def header_functions():
    """Names of the functions declared in include/bhnerf_b200.h (used by the symbol-export test)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), 'include', 'bhnerf_b200.h')
    txt = open(hdr).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(bhnerf_[a-z_0-9]+)\s*\(', txt)))
