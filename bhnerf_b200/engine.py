"""Host-side plumbing above the C ABI: device buffers (torch tensors), streams and workspaces.

PyTorch is used for device memory and streams only; every compute step below is a call into
libbhnerf_b200.so (hand-written CUDA for sm_100a).  Nothing here falls back to torch math."""
import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import IMPL_SIMT, IMPL_TC, LOSS_KINDS, N_PARAMS, Scene, check


def resolve_impl(impl=None):
    impl = impl if impl is not None else os.environ.get('BHNERF_IMPL', 'auto')
    if isinstance(impl, int):
        return impl
    impl = str(impl).lower()
    if impl in ('tc', 'tcgen05', '1'):
        return IMPL_TC
    if impl in ('simt', 'fp32', '0'):
        return IMPL_SIMT
    if impl == 'auto':
        return DEFAULT_IMPL
    raise ValueError('unknown impl %r' % (impl,))


DEFAULT_IMPL = IMPL_TC     # tcgen05 family (parity-green on B200); 'simt' selects the fp32 reference kernels
# cap on the activation workspace of one fused step / backward; more frames than fit are processed in chunks
# default 40 GB of the B200's 180 GB: a full cfg2 step (100 frames x 0.33 GB of saved activations) runs as one chunk
DEFAULT_MAX_WORKSPACE = int(float(os.environ.get('BHNERF_MAX_WORKSPACE_GB', '40')) * 2 ** 30)


def set_max_workspace(nbytes):
    """Cap (bytes) of the activation workspace used when a call does not pass `max_workspace` itself.  The default, 40 GB of
    the B200's 180 GB (or BHNERF_MAX_WORKSPACE_GB), lets a full cfg2 step run as one chunk; a smaller cap only changes how
    many frames a chunk holds, never the result (tests/test_gpu_parity.py: frame chunking)."""
    global DEFAULT_MAX_WORKSPACE
    DEFAULT_MAX_WORKSPACE = int(nbytes)
    return DEFAULT_MAX_WORKSPACE


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32)), device=device)


_workspaces = {}          # device index -> current scratch tensor
_ws_watch = {}            # device index -> weakrefs of every scratch tensor that may still be written (a captured CUDA graph
                          # keeps writing health flags into the workspace it captured, even after a grow-only reallocation)
_pending_flags = {}       # device index -> sticky flags read from a workspace when it was retired


def _dev_key(device):
    return device.index if device.index is not None else torch.cuda.current_device()


def workspace(nbytes, device):
    """Grow-only per-device scratch (allocated off the hot path after the first step)."""
    key = _dev_key(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:                      # carry the retired buffer's sticky flags over (one sync, off the hot path)
            _pending_flags[key] = [a | b for a, b in zip(_pending_flags.get(key, [0] * 8), _read_flags(ws, device))]
        _workspaces[key] = ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        ws[:256].zero_()                        # the caching allocator may hand back memory with stale status words
        _ws_watch.setdefault(key, []).append(weakref.ref(ws))
    return ws


def current_workspace(device):
    """The scratch tensor the next C-ABI step on `device` will use (None before the first step)."""
    return _workspaces.get(_dev_key(torch.device(device)))


def _read_flags(ws, device):
    flags = (C.c_int32 * 8)()
    with torch.cuda.device(device):
        check(_lib.load().bhnerf_workspace_status(_ptr(ws), flags, _stream()))
    return list(flags)


STATUS_FLAGS = ('forward pipeline aborted', 'dgrad chain aborted', 'wgrad aborted',
                'forward activation exceeded the fp16 operand range (|h| > 65504): images invalid',
                'backward produced a non-finite parameter gradient')


def workspace_status(device=None, impl=None, raise_on_error=True):
    """Health flags of the tcgen05 steps on `device` since the previous call (C ABI: bhnerf_workspace_status; the flags are
    sticky and cleared by the read).  Every workspace that can still be written is polled -- the current one and any a
    captured CUDA graph holds.  Synchronises the stream: call it off the hot path (Optimizer.run polls it when it logs).
    Raises BhnerfError if a flag is set (or returns the flags with raise_on_error=False)."""
    if resolve_impl(impl) != IMPL_TC:
        return [0] * 8
    device = torch.device(device if device is not None else 'cuda')
    key = _dev_key(device)
    flags = _pending_flags.pop(key, [0] * 8)
    live = []
    for r in _ws_watch.get(key, []):
        ws = r()
        if ws is not None:
            live.append(r)
            flags = [a | b for a, b in zip(flags, _read_flags(ws, device))]
    _ws_watch[key] = live
    bad = [STATUS_FLAGS[i] for i in range(len(STATUS_FLAGS)) if flags[i]]
    if bad and raise_on_error:
        raise _lib.BhnerfError('bhnerf_b200 tcgen05 step failed: ' + '; '.join(bad))
    return flags


class PackedScene:
    """Prepacked, frame-independent scene: the non-optimised arguments of network.raytracing_args
    (bhnerf/network.py:850-894) + the NeRF_Predictor domain constants (bhnerf/network.py:147-157)."""

    def __init__(self, coords, Omega, J, g, dtau, Sigma, t_geos, t_start_obs, t_injection, scale, rmin, rmax,
                 z_width, GM_c3, device=None):
        lib = _lib.load()
        device = torch.device(device if device is not None else 'cuda')
        coords = _dev_f32(coords, device)
        assert coords.shape[0] == 3, 'coords must be (3, ...)'
        self.image_shape = tuple(coords.shape[1:-1])
        G = coords.shape[-1]
        P = int(np.prod(self.image_shape)) if self.image_shape else 1
        coords = coords.reshape(3, P, G)

        def field(a):
            a = _dev_f32(a, device)
            if a.dim() == 0:
                a = a.expand(P, G)
            return a.reshape(P, G).contiguous()
        Omega, g, dtau, Sigma, t_geos = [field(a) for a in (Omega, g, dtau, Sigma, t_geos)]
        if J is None or np.isscalar(J) or (hasattr(J, 'ndim') and J.ndim == 0):
            assert J is None or float(J) == 1.0, 'scalar J must be 1.0 (network.py:874 default)'
            Jd, S = None, 1
            self.polarized = False
        else:
            Jd = _dev_f32(J, device)
            S = Jd.shape[0]
            Jd = Jd.reshape(S, P, G).contiguous()
            self.polarized = True
        nbytes = lib.bhnerf_packed_bytes(P, G, S)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        sc = Scene()
        with torch.cuda.device(device):
            check(lib.bhnerf_prepack(_ptr(coords), _ptr(Omega), _ptr(g), _ptr(dtau), _ptr(Sigma), _ptr(t_geos),
                                     _ptr(Jd), P, G, S, float(rmin), float(rmax), float(z_width), _ptr(packed),
                                     nbytes, C.byref(sc), _stream()))
        # shrink to what is used: row_ptr + (7+S) arrays of n_pad
        used = ((P + 1 + 127) // 128 * 128) * 4 + (7 + S) * sc.n_pad * 4
        self.packed = packed[:used].clone()
        del packed
        sc.packed = self.packed.data_ptr()
        sc.t_start_obs = float(t_start_obs); sc.GM_c3 = float(GM_c3)
        sc.t_injection = float(t_injection); sc.scale = float(scale)
        self.struct = sc
        self.device = device
        self.P, self.G, self.S = P, G, S
        self.n_active, self.n_pad = sc.n_active, sc.n_pad
        self.rmin, self.rmax, self.z_width = float(rmin), float(rmax), float(z_width)

    @property
    def ref(self):
        return C.byref(self.struct)

    def _int_array(self, which):
        off = ((self.P + 1 + 127) // 128 * 128) * 4 + (5 + self.S + which) * self.n_pad * 4
        return self.packed[off: off + self.n_pad * 4].view(torch.int32)

    @property
    def row_ptr(self):
        return self.packed[: (self.P + 1) * 4].view(torch.int32)

    @property
    def ray_index(self):
        """[n_pad] ray of each compacted sample (-1 = padding)."""
        return self._int_array(0)

    @property
    def dense_index(self):
        """[n_active] flat index ray*G + k of each compacted sample in the dense (P,G) arrays."""
        n = self.n_active
        return self._int_array(0)[:n].long() * self.G + self._int_array(1)[:n].long()


def render_fwd(scene, params, t_frames, impl=None, save_acts=False):
    """images [Bt,S,P], e [Bt,n_pad], acts (or None).  C ABI: bhnerf_render_fwd."""
    lib = _lib.load(); impl = resolve_impl(impl)
    dev = scene.device
    params = _dev_f32(params, dev); t_frames = _dev_f32(t_frames, dev).reshape(-1)
    assert params.numel() == N_PARAMS
    Bt = t_frames.numel()
    images = torch.empty((Bt, scene.S, scene.P), dtype=torch.float32, device=dev)
    e = torch.empty((Bt, scene.n_pad), dtype=torch.float32, device=dev)
    acts = None
    if save_acts:
        acts = torch.empty(lib.bhnerf_acts_bytes(scene.ref, Bt, impl), dtype=torch.uint8, device=dev)
    wsb = lib.bhnerf_fwd_workspace_bytes(impl)
    ws = workspace(wsb, dev) if wsb else None
    with torch.cuda.device(dev):
        check(lib.bhnerf_render_fwd(scene.ref, _ptr(params), _ptr(t_frames), Bt, _ptr(images), _ptr(e), _ptr(acts),
                                    _ptr(ws), wsb, impl, _stream()))
    return images, e, acts


def render_bwd(scene, params, t_frames, d_images, e=None, acts=None, impl=None, max_workspace=None):
    """d_params [55169].  C ABI: bhnerf_render_bwd (recomputes what is not passed in)."""
    lib = _lib.load(); impl = resolve_impl(impl)
    dev = scene.device
    params = _dev_f32(params, dev); t_frames = _dev_f32(t_frames, dev).reshape(-1)
    d_images = _dev_f32(d_images, dev)
    Bt = t_frames.numel()
    assert d_images.numel() == Bt * scene.S * scene.P
    one = lib.bhnerf_bwd_workspace_bytes(scene.ref, Bt, impl)
    fixed = lib.bhnerf_bwd_fixed_workspace_bytes(scene.ref, Bt, impl)
    full = one + (one - fixed - 1024) * (Bt - 1)
    max_workspace = DEFAULT_MAX_WORKSPACE if max_workspace is None else max_workspace
    nbytes = max(one, min(full, int(max_workspace)))
    ws = workspace(nbytes, dev)
    grads = torch.empty(N_PARAMS, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_render_bwd(scene.ref, _ptr(params), _ptr(t_frames), Bt, _ptr(d_images), _ptr(e), _ptr(acts),
                                    _ptr(grads), _ptr(ws), nbytes, impl, _stream()))
    return grads


def loss_image(images, target, sigma, offset, scale, kind):
    """(loss[1], d_images).  C ABI: bhnerf_loss_image (loss_fn_image, bhnerf/network.py:476-484)."""
    lib = _lib.load()
    dev = images.device
    Bt, S, P = images.shape
    target, sigma, offset = [_dev_f32(a, dev) for a in (target, sigma, offset)]
    k = LOSS_KINDS[kind] if isinstance(kind, str) else kind
    want = Bt * S * P if k == 0 else Bt * S
    assert target.numel() == want and sigma.numel() == want and offset.numel() == want, 'target shape mismatch'
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dI = torch.empty_like(images)
    with torch.cuda.device(dev):
        check(lib.bhnerf_loss_image(_ptr(images), _ptr(target), _ptr(sigma), _ptr(offset), float(scale), k, Bt, S, P,
                                    _ptr(loss), _ptr(dI), _stream()))
    return loss, dI


def _c64(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.complex64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.complex64)), device=dev)


_copy_streams = {}


def upload_async(x_host, dev):
    """Host tensor -> device on a per-device COPY stream, so that the transfer overlaps whatever the current stream is
    doing.  Returns (device tensor, event): the consumer calls ``wait_upload`` right before its first use.  A pinned source
    makes the call return at once; a pageable one blocks the host for the staging copy but still overlaps the GPU work that
    is already enqueued."""
    dev = torch.device(dev)
    key = _dev_key(dev)
    cs = _copy_streams.get(key)
    if cs is None:
        cs = _copy_streams[key] = torch.cuda.Stream(device=dev)
    x_host = x_host.contiguous()
    with torch.cuda.stream(cs):
        y = x_host.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
    return y, ev


def wait_upload(y, ev):
    """Make the current stream wait for an ``upload_async`` transfer; returns the device tensor."""
    cur = torch.cuda.current_stream(y.device)
    cur.wait_event(ev)
    y.record_stream(cur)
    return y


def vis_fwd(A, images):
    """vis [Bt,V] complex64 = A[b] @ vec(I[b])  (bhnerf/network.py:542-544).  images [Bt,1,P] or [Bt,P]."""
    lib = _lib.load(); dev = images.device
    Bt, V, P = A.shape
    vis = torch.empty((Bt, V), dtype=torch.complex64, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_vis_fwd(_ptr(A), _ptr(images), Bt, V, P, _ptr(vis), _stream()))
    return vis


def add_inplace(dst, src):
    """dst += src on the device.  C ABI: bhnerf_add_inplace."""
    lib = _lib.load()
    assert dst.numel() == src.numel() and dst.dtype == torch.float32 and src.dtype == torch.float32
    with torch.cuda.device(dst.device):
        check(lib.bhnerf_add_inplace(_ptr(dst), _ptr(src), dst.numel(), _stream()))
    return dst


def loss_vis(vis, target, sigma, scale, kind):
    """(loss[1], d_vis).  'vis'/'amp': vis [Bt,V]; 'cphase': vis [Bt,3V] (baseline-triangle axis folded into V),
    target/sigma [Bt,V].  C ABI: bhnerf_loss_vis (bhnerf/network.py:546-559)."""
    lib = _lib.load(); dev = vis.device
    Bt, V = vis.shape
    k = LOSS_KINDS[kind] if isinstance(kind, str) else kind
    if k == LOSS_KINDS['cphase']:
        assert V % 3 == 0
        V //= 3
    target = _c64(target, dev) if k == LOSS_KINDS['vis'] else _dev_f32(target, dev)
    sigma = _dev_f32(sigma, dev)
    if sigma.numel() == 1:
        sigma = sigma.reshape(1).expand(Bt * V).contiguous()
    assert target.numel() == Bt * V and sigma.numel() == Bt * V
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dvis = torch.empty_like(vis)
    with torch.cuda.device(dev):
        check(lib.bhnerf_loss_vis(_ptr(vis), _ptr(target), _ptr(sigma), float(scale), k, Bt, V, _ptr(loss), _ptr(dvis),
                                  _stream()))
    return loss, dvis


def vis_head(A, images, target, sigma, scale, kind, want_grad=True, group_bytes=0):
    """The eht head of a step in one C-ABI call (bhnerf_vis_head): per L2-sized group of frames vis = A I -> chi^2 ->
    d_images = A^H d_vis.  A [n, rows, P] complex64 (n = frames x polarizations), images [n, 1, P] / [n, P]; target / sigma
    [n, V] (V = rows, or rows/3 for 'cphase').  Returns (loss[1], vis [n, rows], d_images [n, 1, P] or None)."""
    lib = _lib.load(); dev = images.device
    n, rows, P = A.shape
    k = LOSS_KINDS[kind] if isinstance(kind, str) else kind
    V = rows // 3 if k == LOSS_KINDS['cphase'] else rows
    target = _c64(target, dev) if k == LOSS_KINDS['vis'] else _dev_f32(target, dev)
    sigma = _dev_f32(sigma, dev)
    if sigma.numel() == 1:
        sigma = sigma.reshape(1).expand(n * V).contiguous()
    assert target.numel() == n * V and sigma.numel() == n * V and images.numel() == n * P
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    vis = torch.empty((n, rows), dtype=torch.complex64, device=dev)
    dvis = torch.empty_like(vis)
    dI = torch.empty((n, 1, P), dtype=torch.float32, device=dev) if want_grad else None
    with torch.cuda.device(dev):
        check(lib.bhnerf_vis_head(_ptr(A), _ptr(images), _ptr(target), _ptr(sigma), float(scale), k, n, rows, P, _ptr(loss),
                                  _ptr(vis), _ptr(dvis), _ptr(dI), int(group_bytes), _stream()))
    return loss, vis, dI


def vis_dft_fwd(uv, images, grid, pulse=None):
    """Separable visibility head, forward (C ABI: bhnerf_vis_dft_fwd; tcgen05): vis [Bt,V] complex64 from uv [Bt,V,2] and
    images [Bt,NA,NB]; grid = (x0, dx, y0, dy) of the pixel centres; pulse [Bt,V] complex64 or None."""
    lib = _lib.load(); dev = images.device
    uv = _dev_f32(uv, dev)
    Bt, V = uv.shape[:2]
    NA, NB = images.shape[-2:]
    assert images.numel() == Bt * NA * NB
    pulse = None if pulse is None else _c64(pulse, dev)
    vis = torch.empty((Bt, V), dtype=torch.complex64, device=dev)
    st = torch.zeros(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_vis_dft_fwd(_ptr(uv), _ptr(pulse), _ptr(images), Bt, V, NA, NB, *[float(v) for v in grid], _ptr(vis),
                                     _ptr(st), _stream()))
    return vis


def vis_dft_bwd(uv, dvis, grid, NA, NB, pulse=None):
    """Separable visibility head, pull-back to the images (C ABI: bhnerf_vis_dft_bwd): d_images [Bt,NA,NB]."""
    lib = _lib.load(); dev = dvis.device
    uv = _dev_f32(uv, dev)
    Bt, V = uv.shape[:2]
    pulse = None if pulse is None else _c64(pulse, dev)
    dI = torch.empty((Bt, NA, NB), dtype=torch.float32, device=dev)
    st = torch.zeros(2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_vis_dft_bwd(_ptr(uv), _ptr(pulse), _ptr(dvis), Bt, V, NA, NB, *[float(v) for v in grid], _ptr(dI),
                                     _ptr(st), _stream()))
    return dI


def vis_bwd(A, dvis, P):
    lib = _lib.load(); dev = dvis.device
    Bt, V = dvis.shape
    dI = torch.empty((Bt, 1, P), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.bhnerf_vis_bwd(_ptr(A), _ptr(dvis), Bt, V, P, _ptr(dI), _stream()))
    return dI


GRID_MODE_DYNAMICS, GRID_MODE_PREDICTOR = 0, 1


def _fov3(fov):
    f = np.broadcast_to(np.asarray(fov, dtype=np.float64), (3,))
    return [float(f[0]), float(f[1]), float(f[2])]


def grid_render_fwd(scene, grid, fov, t_frames, mode, want_e=False):
    """images [Bt,S,P] (and per-sample emission [Bt,n_pad]) of the voxel-grid renderer.  C ABI: bhnerf_grid_render_fwd."""
    lib = _lib.load(); dev = scene.device
    grid = _dev_f32(grid, dev); t_frames = _dev_f32(t_frames, dev).reshape(-1)
    assert grid.dim() == 3
    Bt = t_frames.numel()
    images = torch.empty((Bt, scene.S, scene.P), dtype=torch.float32, device=dev)
    e = torch.zeros((Bt, scene.n_pad), dtype=torch.float32, device=dev) if want_e else None
    fx, fy, fz = _fov3(fov)
    with torch.cuda.device(dev):
        check(lib.bhnerf_grid_render_fwd(scene.ref, _ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2], fx, fy, fz,
                                         int(mode), _ptr(t_frames), Bt, _ptr(images), _ptr(e), _stream()))
    return images, e


def grid_render_bwd(scene, grid, fov, t_frames, d_images):
    """d_grid (same shape as grid) of GRID_Predictor's render.  C ABI: bhnerf_grid_render_bwd."""
    lib = _lib.load(); dev = scene.device
    grid = _dev_f32(grid, dev); t_frames = _dev_f32(t_frames, dev).reshape(-1)
    d_images = _dev_f32(d_images, dev)
    Bt = t_frames.numel()
    assert d_images.numel() == Bt * scene.S * scene.P
    d_grid = torch.empty_like(grid)
    fx, fy, fz = _fov3(fov)
    with torch.cuda.device(dev):
        check(lib.bhnerf_grid_render_bwd(scene.ref, _ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2], fx, fy, fz,
                                         _ptr(t_frames), Bt, _ptr(d_images), _ptr(d_grid), _stream()))
    return d_grid


def train_step_image(scene, params, t_frames, target, sigma, offset, scale, kind, impl=None, max_workspace=None,
                     out=None):
    """Fused fwd -> ray integral -> loss -> bwd.  Returns (loss[1], images [Bt,S,P], grads [55169]).
    C ABI: bhnerf_train_step_image.  All inputs must already be device tensors for the hot loop."""
    lib = _lib.load(); impl = resolve_impl(impl)
    dev = scene.device
    params = _dev_f32(params, dev); t_frames = _dev_f32(t_frames, dev).reshape(-1)
    target, sigma, offset = [_dev_f32(a, dev) for a in (target, sigma, offset)]
    Bt = t_frames.numel()
    k = LOSS_KINDS[kind] if isinstance(kind, str) else kind
    want = Bt * scene.S * scene.P if k == 0 else Bt * scene.S
    assert target.numel() == want and sigma.numel() == want and offset.numel() == want, 'target shape mismatch'
    full = lib.bhnerf_train_workspace_bytes(scene.ref, Bt, impl)
    one = lib.bhnerf_train_workspace_bytes(scene.ref, 1, impl)
    max_workspace = DEFAULT_MAX_WORKSPACE if max_workspace is None else max_workspace
    nbytes = max(one, min(full, int(max_workspace)))
    ws = workspace(nbytes, dev)
    if out is None:
        out = (torch.empty(1, dtype=torch.float32, device=dev),
               torch.empty((Bt, scene.S, scene.P), dtype=torch.float32, device=dev),
               torch.empty(N_PARAMS, dtype=torch.float32, device=dev))
    loss, images, grads = out
    with torch.cuda.device(dev):
        check(lib.bhnerf_train_step_image(scene.ref, _ptr(params), _ptr(t_frames), Bt, _ptr(target), _ptr(sigma),
                                          _ptr(offset), float(scale), k, _ptr(loss), _ptr(images), _ptr(grads),
                                          _ptr(ws), nbytes, impl, _stream()))
    return loss, images, grads


def frames_per_chunk(scene, Bt, impl=None, max_workspace=None):
    """How many frames of saved activations fit the workspace cap (>= 1)."""
    lib = _lib.load(); impl = resolve_impl(impl)
    per = max(lib.bhnerf_acts_bytes(scene.ref, 1, impl), 1)
    cap = DEFAULT_MAX_WORKSPACE if max_workspace is None else int(max_workspace)
    bc = int(max(1, min(Bt, cap // per)))
    nchunks = (Bt + bc - 1) // bc
    return (Bt + nchunks - 1) // nchunks          # equal chunks instead of a short tail


def adam_step(params, grads, mu, nu, count, lr_init=1e-4, lr_final=1e-6, num_iters=5000, b1=0.9, b2=0.999,
              eps=1e-8, grad_scale=1.0, guard=None):
    """In-place optax.adam + linear schedule on the flat buffers.  C ABI: bhnerf_adam_step.  `guard`: the tcgen05
    workspace whose health flags veto the update (None = unguarded)."""
    lib = _lib.load()
    with torch.cuda.device(params.device):
        check(lib.bhnerf_adam_step(_ptr(params), _ptr(grads), _ptr(mu), _ptr(nu), params.numel(), int(count),
                                   float(lr_init), float(lr_final), int(num_iters), float(b1), float(b2), float(eps),
                                   float(grad_scale), _ptr(guard), _stream()))
    return params


def step_guard(device, impl=None):
    """Workspace to pass as `guard` to the Adam update that follows a render step of kernel family `impl` (tcgen05 only:
    the fp32 SIMT family keeps weights, not flags, at the head of its scratch)."""
    return current_workspace(device) if resolve_impl(impl) == IMPL_TC else None


# ------------------------------------------------------------------------------------------------
# gradient exchange below the C ABI (bhnerf_allreduce_mean / _sum: NCCL on the kernel stream)
# ------------------------------------------------------------------------------------------------
_comms = {}


def _dist_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def comm(device=None):
    """This rank's NCCL communicator inside libbhnerf_b200 (None with a single rank).  Created collectively on first
    use: rank 0 draws the unique id (bhnerf_comm_unique_id), torch.distributed -- the host-side plumbing -- broadcasts its
    128 bytes, every rank calls bhnerf_comm_init.  Afterwards the data path never touches torch.distributed."""
    dist, rank, world = _dist_world()
    if dist is None:
        return None
    key = torch.cuda.current_device() if device is None else _dev_key(torch.device(device))
    hit = _comms.get(key)
    if hit is not None:
        return hit
    lib = _lib.load()
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        check(lib.bhnerf_comm_unique_id(idbuf))
    box = [idbuf.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    idbuf = C.create_string_buffer(box[0], 128)
    handle = C.c_void_p()
    with torch.cuda.device(key):
        check(lib.bhnerf_comm_init(rank, world, idbuf, C.byref(handle)))
    _comms[key] = handle
    return handle


def allreduce_mean(t):
    """In-place mean over ranks of a float32 device tensor (jax.lax.pmean, network.py:620).  C ABI: bhnerf_allreduce_mean."""
    c = comm(t.device)
    if c is None:
        return t
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    with torch.cuda.device(t.device):
        check(_lib.load().bhnerf_allreduce_mean(_ptr(t), t.numel(), c, _stream()))
    return t


def allreduce_sum(t):
    """In-place sum over ranks (float32, or complex64 seen as interleaved float32).  C ABI: bhnerf_allreduce_sum."""
    c = comm(t.device)
    if c is None:
        return t
    assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.complex64)
    n = t.numel() * (2 if t.dtype == torch.complex64 else 1)
    with torch.cuda.device(t.device):
        check(_lib.load().bhnerf_allreduce_sum(_ptr(t), n, c, _stream()))
    return t


def lightcurve(images):
    """lc [Bt,S] = sum over rays of images [Bt,S,P].  C ABI: bhnerf_lightcurve."""
    Bt, S, P = images.shape
    lc = torch.empty((Bt, S), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        check(_lib.load().bhnerf_lightcurve(_ptr(images), Bt, S, P, _ptr(lc), _stream()))
    return lc


def loss_lightcurve(lc, target, sigma, offset, scale, P):
    """(loss[1], d_images [Bt,S,P]) of the 'lc' head from (all-reduced) lightcurves.  C ABI: bhnerf_loss_lightcurve."""
    dev = lc.device
    Bt, S = lc.shape
    target, sigma, offset = [_dev_f32(a, dev).reshape(Bt, S) for a in (target, sigma, offset)]
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dI = torch.empty((Bt, S, P), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().bhnerf_loss_lightcurve(_ptr(lc), _ptr(target), _ptr(sigma), _ptr(offset), float(scale), Bt, S, P,
                                                 _ptr(loss), _ptr(dI), _stream()))
    return loss, dI
