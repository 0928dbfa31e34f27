"""Mirror of the reference's ``bhnerf/optimization.py`` for the train-step orchestration.

Data parallelism: the reference uses one process driving all GPUs with ``jax.pmap`` over observation
frames and ``jax.lax.pmean`` of the gradients (optimization.py:209-216, network.py:620).  Here it is one
process per GPU (``torch.distributed``, NCCL): ``shard`` gives each rank its contiguous slice of the
batch's frames, every rank holds replicated geodesics + params, and the only exchange is the
all-reduce(mean) of the 55 169-float gradient before Adam."""
import os
import pickle

import numpy as np
import torch

from . import engine, network, utils


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _same_on_all_ranks(values):
    """The reference samples batch indices in its single driver process; with one process per GPU rank 0's
    draw is broadcast so every rank works on the same batch."""
    import torch.distributed as dist
    if _world()[1] == 1:
        return values
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.as_tensor(np.asarray(values, dtype=np.int64), device=dev).clone()
    dist.broadcast(t, src=0)
    return t.cpu().numpy().reshape(np.shape(values))


def device_count():
    """jax.device_count() analogue: number of ranks (one GPU each)."""
    return _world()[1]


def ray_sharded(n_frames):
    """True when a batch of n_frames cannot be split over the ranks by frames (fewer frames than ranks, or not a
    multiple): the step then shards RAYS instead (network._ray_sharded_image_step) -- every rank keeps the whole batch.
    The reference has no such mode: its shard() simply fails (optimization.py:360-362), which is why its own scripts'
    batchsize 6 cannot use 8 devices."""
    return _world()[1] > 1 and n_frames % _world()[1] != 0


def shard(xs):
    """optimization.py:360-362 reshapes to (ndev, -1, ...) and pmap gives device d the d-th slice; here the
    calling rank gets its slice directly.  A batch that does not divide evenly is left whole on every rank
    (ray sharding, see ray_sharded)."""
    rank, world = _world()

    def one(x):
        n = x.shape[0]
        if n % world:
            return x
        per = n // world
        return x[rank * per:(rank + 1) * per]
    if isinstance(xs, (list, tuple)):
        return type(xs)(one(x) for x in xs)
    return one(xs)


def total_movie_loss(batchsize, state, train_step, raytracing_args, return_frames=False):
    """optimization.py:14-66: chunk all frames, forward only, sum the loss / nt, optionally return frames.
    With several ranks the per-rank losses are summed (the reference's ``loss.sum()`` over devices) and
    the frames of every rank are gathered in order."""
    nt = train_step.args[0].num_frames
    ndev = device_count()
    if nt % ndev:
        raise AttributeError('batch size should be an integer multiplication of the device number')
    nt_tilde = nt - nt % batchsize
    indices = np.array_split(np.arange(0, nt_tilde), nt_tilde / batchsize) if nt_tilde else []
    nt_tilde1 = int(ndev * np.ceil(nt / ndev))
    indices.append(np.arange(nt_tilde, nt_tilde1) % nt)
    frames, total_loss = [], 0.0
    for inds in indices:
        if inds.size == 0:
            break
        loss, state, images = train_step(state, raytracing_args, inds, update_state=False)
        loss = _allreduce_scalar(loss)
        total_loss += float(loss)
        if return_frames:
            images = _allgather_frames(images)
            polarized = not np.isscalar(np.atleast_1d(raytracing_args)[0]['J'])
            frames.append(images.reshape((-1,) + tuple(images.shape[-3:] if polarized else images.shape[-2:])))
    output = total_loss / nt
    if return_frames:
        output = (output, np.concatenate([f.cpu().numpy() for f in frames])[:nt])
    return output


def _allreduce_scalar(loss):
    import torch.distributed as dist
    t = loss.detach().reshape(-1).sum().reshape(1).clone()
    if _world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.item()


def _allgather_frames(images):
    import torch.distributed as dist
    rank, world = _world()
    if world == 1:
        return images
    outs = [torch.empty_like(images) for _ in range(world)]
    dist.all_gather(outs, images.contiguous())
    return torch.cat(outs, dim=0)


class TemporalBatchedArgs(object):
    """optimization.py:270-302."""

    def __init__(self, t_frames, args=[]):
        self.t_frames = t_frames
        if not isinstance(args, list):
            args = [args]
        self.num_frames = len(t_frames)
        assert all([self.num_frames == arg.shape[0] for arg in args])
        self._t_values = np.asarray(utils.time_value(t_frames, 'hr'), dtype=np.float32)
        args = list(args) + [self._t_values]
        self.args = args
        self.default_t_units = 'hr'

    def sample(self, batchsize, replace=False):
        # same draw as the reference's np.random.choice(range(n), ...) for a given seed, without building the range
        return _same_on_all_ranks(np.random.choice(self.num_frames, batchsize, replace=replace))

    def __getitem__(self, key):
        return [shard(arg[key, ...]) for arg in self.args]

    @property
    def t_units(self):
        unit = getattr(self.t_frames, 'unit', None)
        return str(unit) if unit is not None else self.default_t_units

    @property
    def t_start_obs(self):
        return self.t_frames[0]


class TrainStep(object):
    """optimization.py:145-268.  ``grad_pmap``/``test_pmap`` hold the step functions themselves (one rank =
    one device; there is no pmap to build)."""

    def __init__(self, dtype, args, grad_pmap, test_pmap, scale):
        self.dtype = np.atleast_1d(dtype)
        self.args = np.atleast_1d(args)
        self.grad_pmap = np.atleast_1d(grad_pmap)
        self.test_pmap = np.atleast_1d(test_pmap)
        self.scale = np.atleast_1d(scale)
        if np.any([arg.t_units not in ('hr', 'h') for arg in self.args]):
            raise AttributeError('only hr units supported')
        assert self.dtype.size == self.args.size == self.test_pmap.size == \
            self.grad_pmap.size == self.scale.size, 'input list sizes are not equal'
        self.num_losses = self.dtype.size

    def __call__(self, state, raytracing_args, indices, update_state=True):
        total_loss = 0.0
        total_images = 0.0
        raytracing_args = np.atleast_1d(raytracing_args)
        if update_state:
            call_fn = self.grad_pmap
            raytracing_args = [raytracing_args[int(_same_on_all_ranks(np.random.choice(len(raytracing_args))))]]
        else:
            call_fn = self.test_pmap
        kw = {'ray_shard': _world()} if ray_sharded(len(np.atleast_1d(indices))) else {}
        for rt_arg in raytracing_args:
            for i in range(self.num_losses):
                loss, state, images = call_fn[i](state, self.t_units, self.dtype[i], *self.args[i][indices],
                                                 *rt_arg.values(), self.scale[i], **kw)
                total_loss = total_loss + loss / len(raytracing_args)
                total_images = total_images + images / len(raytracing_args)
        return total_loss, state, total_images

    def __add__(self, other):
        return TrainStep(np.append(self.dtype, other.dtype), np.append(self.args, other.args),
                         np.append(self.grad_pmap, other.grad_pmap), np.append(self.test_pmap, other.test_pmap),
                         np.append(self.scale, other.scale))

    @classmethod
    def image(cls, t_frames, target, sigma=1.0, offset=0.0, scale=1.0, dtype='full'):
        """optimization.py:189-216."""
        target = np.asarray(target, dtype=np.float32)
        sigma = (sigma * np.ones_like(target)).astype(np.float32)
        offset = (offset * np.ones_like(target)).astype(np.float32)
        args = TemporalBatchedArgs(t_frames, [target, sigma, offset])
        return cls(dtype, args, network.gradient_step_image, network.test_image, scale)

    @classmethod
    def eht(cls, t_frames, obs, image_fov, image_size, chisqdata, pol='I', scale=1.0):
        """optimization.py:218-268, same signature and steps: split the ehtim observation into per-frame scans, build the
        square prior image, call ``chisqdata(obs_frame, prior, mask=[], pol=p)`` (ehtim.imaging.imager_utils.chisqdata_vis /
        _amp / _cphase) per frame and polarization, stack the polarizations on axis 1 and squeeze, convert ehtim's
        closure phases from degrees to radians (:253-255).  ehtim is a third-party package that is not vendored by the
        reference: ``obs`` only needs ``split_obs`` and ``ehtim.image.make_square`` must be importable.  When (target,
        sigma, A) are already at hand use :meth:`eht_arrays`."""
        from ehtim.image import make_square
        dtype = chisqdata.__name__.split('_')[-1]
        pol_types = ['I', 'Q', 'U']
        span = t_frames[-1] - t_frames[0]
        span_s = span.to('s').value if hasattr(span, 'to') else float(span) * 3600.0      # unit-less frames are hours
        obs_frames = obs.split_obs(t_gather=span_s / (len(t_frames) + 1))
        prior = make_square(obs, image_size, image_fov)
        target, sigma, A = [], [], []
        for p in np.atleast_1d(pol):
            if p not in pol_types:
                raise AttributeError('pol ({}) not in supported pol_types: {},{},{}'.format(p, *pol_types))
            target_p, sigma_p, A_p = [np.array(out) for out in zip(*[chisqdata(o, prior, mask=[], pol=p) for o in obs_frames])]
            target.append(target_p); sigma.append(sigma_p); A.append(A_p)
        target, sigma, A = (np.squeeze(np.stack(target, axis=1)), np.squeeze(np.stack(sigma, axis=1)),
                            np.squeeze(np.stack(A, axis=1)))
        return cls.eht_arrays(t_frames, target, sigma, A, dtype=dtype, scale=scale, degrees=(dtype == 'cphase'))

    @classmethod
    def eht_arrays(cls, t_frames, target, sigma, A, dtype='vis', scale=1.0, degrees=False):
        """The reference's ``TrainStep.eht`` after its ehtim calls: (target, sigma, A) as ehtim's chisqdata_<dtype> returns
        them, stacked per frame.  A: (nt, nvis, npix) complex64, or (nt, npol, nvis, npix) for several polarizations
        (optimization.py:235-251; target / sigma then carry the same pol axis), or (nt, [npol,] 3, ncphase, npix) for closure
        phases.  ``degrees=True`` applies the reference's np.deg2rad to closure-phase targets and sigmas (:253-255: ehtim
        reports them in degrees); pass radians otherwise."""
        target = np.asarray(target)
        sigma = np.asarray(sigma, dtype=np.float32)
        if degrees:
            if dtype != 'cphase':
                raise AttributeError('degrees=True only applies to closure phases (dtype=cphase)')
            target, sigma = np.deg2rad(target), np.deg2rad(sigma).astype(np.float32)
        if not isinstance(A, network.SeparableDFT):        # opt-in: baselines instead of the matrix (tensor-core DFT head)
            A = np.asarray(A, dtype=np.complex64)
        args = TemporalBatchedArgs(t_frames, [target, sigma, A])
        return cls(dtype, args, network.gradient_step_eht, network.test_eht, scale)

    @property
    def t_units(self):
        return self.args[0].t_units


# ---- flax msgpack state format -----------------------------------------------------------------------------------
# flax.training.checkpoints.save_checkpoint writes flax.serialization.to_bytes(state) = msgpack of the state dict, with
# every ndarray as ExtType(1, msgpack((shape, dtype.name, raw bytes))) and numpy scalars as ExtType(3, ...)
# (flax/serialization.py, published format; flax is not installed here, so this restatement is UNPINNED against flax
# itself -- it is checked against hand-built byte strings of that format in tests/test_abi_host.py).  The state dict of
# TrainState(params, tx=optax.adam(schedule)) is {'step', 'params', 'opt_state': {'0': {'count','mu','nu'}, '1': {'count'}}}
# (optax.chain(scale_by_adam, scale_by_schedule), network.py:173-182).
def _np_to_ext(a):
    import msgpack
    a = np.asarray(a)
    return msgpack.ExtType(1, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes('C')), use_bin_type=True))


def _flax_pack(obj):
    import msgpack
    if isinstance(obj, dict):
        return {str(k): _flax_pack(v) for k, v in obj.items()}
    if isinstance(obj, (np.ndarray, np.generic)):
        return _np_to_ext(obj)
    if isinstance(obj, torch.Tensor):
        return _np_to_ext(obj.detach().cpu().numpy())
    return obj


def _flax_ext_hook(code, data):
    import msgpack
    if code == 1:
        shape, dtype, buf = msgpack.unpackb(data, raw=True)
        return np.frombuffer(buf, dtype=np.dtype(dtype.decode() if isinstance(dtype, bytes) else dtype)).reshape(shape).copy()
    if code == 3:      # numpy scalar: flax packs it as the 0-d array np.asarray(x) -- the same (shape, dtype, bytes) triple
        shape, dtype, buf = msgpack.unpackb(data, raw=True)
        arr = np.frombuffer(buf, dtype=np.dtype(dtype.decode() if isinstance(dtype, bytes) else dtype)).reshape(shape)
        return arr[()]
    if code == 2:
        re_, im_ = msgpack.unpackb(data)
        return complex(re_, im_)
    return msgpack.ExtType(code, data)


def state_to_flax_bytes(state):
    """Bytes of flax.serialization.to_bytes(TrainState) for this state (flat buffers -> flax pytree names)."""
    import msgpack
    pred = state.predictor
    tree = lambda flat: pred._unflatten(flat)
    d = {'step': int(state.step), 'params': tree(state.flat),
         'opt_state': {'0': {'count': np.asarray(state.step, dtype=np.int32), 'mu': tree(state.mu), 'nu': tree(state.nu)},
                       '1': {'count': np.asarray(state.step, dtype=np.int32)}}}
    return msgpack.packb(_flax_pack(d), use_bin_type=True)


def state_from_flax_bytes(state, blob):
    """Inverse of state_to_flax_bytes; also reads checkpoints written by the reference (flax) itself."""
    import msgpack
    d = msgpack.unpackb(blob, ext_hook=_flax_ext_hook, raw=False, strict_map_key=False)
    pred, dev = state.predictor, state.flat.device
    state.flat.copy_(torch.as_tensor(pred._flatten(d['params']), device=dev))
    opt = d.get('opt_state', {})
    adam = opt.get('0', opt) if isinstance(opt, dict) else {}
    if 'mu' in adam and 'nu' in adam:
        state.mu.copy_(torch.as_tensor(pred._flatten(adam['mu']), device=dev))
        state.nu.copy_(torch.as_tensor(pred._flatten(adam['nu']), device=dev))
    state.step = int(np.asarray(d.get('step', adam.get('count', 0))))
    return state


def save_checkpoint(checkpoint_dir, state, step, keep=5, fmt=None):
    """flax.training.checkpoints.save_checkpoint (optimization.py:118-121): file ``checkpoint_<step>`` holding the
    msgpack state bytes flax writes (``fmt='flax'``, default; ``BHNERF_CHECKPOINT_FORMAT=pickle`` or ``fmt='pickle'`` keeps
    the earlier pickled dict), same ``keep`` policy."""
    fmt = fmt or os.environ.get('BHNERF_CHECKPOINT_FORMAT', 'flax')
    os.makedirs(checkpoint_dir, exist_ok=True)
    final = os.path.join(checkpoint_dir, 'checkpoint_%d' % step)
    tmp = os.path.join(checkpoint_dir, '.tmp_checkpoint_%d_%d' % (step, os.getpid()))
    with open(tmp, 'wb') as f:                     # written aside and renamed: an interrupted save never leaves a
        if fmt == 'pickle':                        # truncated file that restore_checkpoint would pick as the latest
            pickle.dump(state.state_dict(), f)
        else:
            f.write(state_to_flax_bytes(state))
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, final)
    for s in _checkpoint_steps(checkpoint_dir)[:-keep]:
        os.remove(os.path.join(checkpoint_dir, 'checkpoint_%d' % s))


def _checkpoint_steps(checkpoint_dir):
    """Sorted steps of the ``checkpoint_<step>`` files of a directory (anything else, e.g. ``checkpoint_tmp``, is ignored)."""
    out = []
    for n in os.listdir(checkpoint_dir):
        parts = n.split('_')
        if len(parts) == 2 and parts[0] == 'checkpoint' and parts[1].isdigit():
            out.append(int(parts[1]))
    return sorted(out)


def restore_checkpoint(checkpoint_dir, state):
    """flax.training.checkpoints.restore_checkpoint (network.py:185): latest ``checkpoint_<step>`` of the directory,
    flax msgpack or the earlier pickle payload (detected by the pickle protocol marker)."""
    if not os.path.isdir(checkpoint_dir):
        return state
    ck = _checkpoint_steps(checkpoint_dir)
    if not ck:
        return state
    with open(os.path.join(checkpoint_dir, 'checkpoint_%d' % ck[-1]), 'rb') as f:
        blob = f.read()
    if blob[:1] == b'\x80':                      # pickle protocol >= 2
        d = pickle.loads(blob)
        dev = state.flat.device
        state.flat.copy_(torch.as_tensor(state.predictor._flatten(d["params"]), device=dev))
        state.mu.copy_(torch.as_tensor(d['mu'], device=dev)); state.nu.copy_(torch.as_tensor(d['nu'], device=dev))
        state.step = int(d['step'])
        return state
    return state_from_flax_bytes(state, blob)


class Optimizer(object):
    """optimization.py:68-143."""

    def __init__(self, hparams, predictor, raytracing_args, save_period=-1, checkpoint_dir='', keep=5):
        self.step = 0
        self.init_step = 0
        self.num_iters = hparams['num_iters']
        self.checkpoint_dir = checkpoint_dir
        self.save_period = self.num_iters if save_period < 0 else save_period
        self.loss = np.inf
        self.keep = keep
        self.seed = hparams.get('seed', 1)
        self.status_period = int(hparams.get('status_period', 100))
        params = predictor.init_params(raytracing_args, seed=self.seed)
        self.state = predictor.init_state(params=params, num_iters=self.num_iters,
                                          lr_init=hparams.get('lr_init', 1e-4), lr_final=hparams.get('lr_final', 1e-6),
                                          lr_inject=hparams.get('lr_inject', None), checkpoint_dir=self.checkpoint_dir)
        if checkpoint_dir != '':
            predictor.save_params(checkpoint_dir)

    def log(self):
        fired = False
        for log_fn in self.log_fns:
            fired = bool(log_fn(self)) or fired
        # health flags at log cadence (the log functions read the loss, i.e. synchronise, anyway) and every
        # `status_period` steps regardless; the flags are sticky, so an overflow on ANY step in between is seen
        if fired or (self.step % self.status_period == 0):
            self.check_health()

    def check_health(self):
        """Raise on every rank if any rank's tcgen05 steps flagged an overflow / abort since the last check.  The Adam
        kernels are guarded by the same flags, so the flagged gradient itself was never applied."""
        flags = engine.workspace_status(device=self.state.flat.device, raise_on_error=False)
        bad = torch.tensor([1.0 if any(flags[:5]) else 0.0], device=self.state.flat.device)
        if _world()[1] > 1:
            import torch.distributed as dist
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)        # collective: no rank runs ahead into the next all-reduce
        if bad.item() > 0:
            mine = [engine.STATUS_FLAGS[i] for i in range(5) if flags[i]]
            raise engine._lib.BhnerfError('bhnerf_b200 tcgen05 step failed at or before step %d: %s' % (
                self.step, '; '.join(mine) if mine else 'flagged on another rank'))

    def save_checkpoint(self):
        if (self.checkpoint_dir != '') and ((self.step % self.save_period == 0) or (self.step == self.final_step)):
            self.check_health()             # never checkpoint a state an overflowed tcgen05 step produced (raises)
            if _world()[0] == 0:
                save_checkpoint(self.checkpoint_dir, self.state, int(self.step), keep=self.keep)

    def run(self, batchsize, train_step, raytracing_args, log_fns=[]):
        self.init_step = self.state.step + 1
        self.final_step = self.init_step + self.num_iters
        self.log_fns = np.atleast_1d(log_fns)
        self.train_step = train_step
        self.raytracing_args = raytracing_args
        try:
            for self.step in range(self.init_step, self.final_step):
                batch_indices = train_step.args[0].sample(batchsize)
                self.loss, self.state, images = train_step(self.state, raytracing_args, indices=batch_indices)
                self.log()
                self.save_checkpoint()
        except KeyboardInterrupt:
            return
        self.check_health()                 # flags raised since the last poll (one sync, off the hot loop)

    @property
    def params(self):
        return self.state.params


class LogFn(object):
    """optimization.py:349-357."""

    def __init__(self, log_fn, log_period=1):
        self.log_period = log_period
        self.log_fn = log_fn

    def __call__(self, optimizer):
        if self.log_period > 0:
            if (optimizer.step == 1) or ((optimizer.step % self.log_period) == 0):
                self.log_fn(optimizer)
                return True
        return False
