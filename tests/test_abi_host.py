"""CPU tests: the C-ABI library builds, loads and exports every symbol the header declares; host logic
(parameter packing, frame sharding over 2 gloo ranks, schedule, API error behaviour)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol(built_lib):
    from bhnerf_b200 import _lib
    declared = _lib.header_functions()
    assert len(declared) >= 20
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.bhnerf_version() == 2
    # nm view: every declared function is an exported text symbol
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert (' T ' + name) in out, name


def test_library_contains_sm100a_code(built_lib):
    from bhnerf_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from bhnerf_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.BhnerfError, match='no CPU fallback'):
        _lib.load()


def test_param_packing_matches_oracle_layout():
    from bhnerf_b200 import network
    from oracle import bhnerf_oracle as O
    p = O.trained_like_params(5)
    np.testing.assert_array_equal(network.flatten_params(p), O.flatten_params(p))
    q = network.unflatten_params(network.flatten_params(p))
    np.testing.assert_array_equal(q['MLP_0']['Dense_3']['kernel'], p['MLP_0']['Dense_3']['kernel'])
    bad = {'MLP_0': dict(p['MLP_0'])}
    bad['MLP_0']['Dense_3'] = {'kernel': np.zeros((128, 128), np.float32), 'bias': np.zeros(128, np.float32)}
    with pytest.raises(ValueError):
        network.flatten_params(bad)


def test_predictor_rejects_unsupported_architectures():
    from bhnerf_b200 import network
    network.NeRF_Predictor(8.0, 2.0, 8.0, 4.0)
    with pytest.raises(NotImplementedError):
        network.NeRF_Predictor(net_width=256)
    with pytest.raises(NotImplementedError):
        network.NeRF_Predictor(posenc_deg=4)


def test_constants_match_oracle():
    from bhnerf_b200 import constants
    from oracle import bhnerf_oracle as O
    assert abs(constants.GM_c3(t_units='hr') - O.GM_C3_SGRA_HR) < 1e-18
    assert abs(constants.GM_c3(t_units='hr') - 5.68347e-3) < 1e-8
    assert abs(constants.isco_pro(0.0) - 6.0) < 1e-12


def test_temporal_batched_args_and_trainstep_single_rank():
    from bhnerf_b200 import optimization as opt
    t = np.linspace(0, 1, 12)
    tgt = np.arange(12 * 4, dtype=np.float32).reshape(12, 4)
    ts = opt.TrainStep.image(t, tgt, sigma=0.5, dtype='lc')
    a = ts.args[0]
    assert a.num_frames == 12 and a.t_units == 'hr'
    target, sigma, offset, tf = a[np.array([3, 7])]
    np.testing.assert_array_equal(target, tgt[[3, 7]])
    assert (sigma == 0.5).all() and (offset == 0).all()
    np.testing.assert_allclose(tf, t[[3, 7]].astype(np.float32))
    idx = a.sample(6)
    assert len(set(idx.tolist())) == 6 and idx.max() < 12
    both = ts + opt.TrainStep.image(t, tgt, dtype='full')
    assert both.num_losses == 2 and list(both.dtype) == ['lc', 'full']


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from bhnerf_b200 import optimization as opt, network
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=2)
rank = dist.get_rank()
t = np.linspace(0, 1, 8); tgt = np.arange(8 * 3, dtype=np.float32).reshape(8, 3)
a = opt.TemporalBatchedArgs(t, [tgt])
np.random.seed(100 + rank)                       # different draws per rank ...
idx = a.sample(4)                                # ... but rank 0's draw is used everywhere
gathered = [None, None]; dist.all_gather_object(gathered, idx.tolist())
assert gathered[0] == gathered[1], gathered
target, tf = a[idx]
assert target.shape == (2, 3)                    # 4 frames / 2 ranks (optimization.py:360-362)
np.testing.assert_array_equal(target, tgt[idx][rank * 2:(rank + 1) * 2])
# pmean: all-reduce SUM then 1/ndev (network.py:620)
g = torch.full((5,), float(rank + 1))
class S:                                          # minimal state stub (no GPU on this box)
    def apply_gradients(self, grads, grad_scale=1.0): self.g = grads * grad_scale; return self
s = network._pmean_and_apply(S(), g)
assert torch.allclose(s.g, torch.full((5,), 1.5)), s.g
# a batch the ranks cannot split by frames stays whole on every rank: the step then shards RAYS (SURVEY s8e(2));
# the reference's shard() cannot run it at all (optimization.py:360-362)
assert opt.ray_sharded(3) and not opt.ray_sharded(4) and opt.ray_sharded(1)
assert opt.shard(np.zeros((3, 2))).shape == (3, 2)
whole, tfw = a[idx[:3]]
assert whole.shape == (3, 3) and tfw.shape == (3,)
# this rank's contiguous block of rays of (A,B,G) / (3,A,B,G) / (S,A,B,G) arrays, ray-major
om = np.arange(4 * 6 * 5, dtype=np.float32).reshape(4, 6, 5)
blk = network._ray_block(om, (rank, 2), 0)
assert blk.shape == (12, 1, 5)
np.testing.assert_array_equal(blk.reshape(12, 5), om.reshape(24, 5)[rank * 12:(rank + 1) * 12])
co = np.stack([om, om + 1000, om + 2000])
np.testing.assert_array_equal(network._ray_block(co, (rank, 2), 1).reshape(3, 12, 5), co.reshape(3, 24, 5)[:, rank * 12:(rank + 1) * 12])
try:
    network._ray_block(np.zeros((3, 3, 2)), (rank, 2), 0)
    raise SystemExit('9 rays on 2 ranks should be rejected')
except ValueError:
    pass
# the ray-sharded gradient exchange is a SUM (per-rank gradients are partial sums of one device's gradient)
s2 = network._pmean_and_apply(S(), torch.full((5,), float(rank + 1)), reduce='sum')
assert torch.allclose(s2.g, torch.full((5,), 3.0)), s2.g
assert opt.device_count() == 2
loss = opt._allreduce_scalar(torch.tensor([float(rank + 1)]))
assert loss == 3.0
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_frame_sharding_and_pmean_two_gloo_ranks(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_flax_msgpack_checkpoint_format_roundtrip(tmp_path):
    """optimization.save_checkpoint / restore_checkpoint write and read the msgpack state bytes flax writes
    (flax.serialization: ndarray = ExtType(1, msgpack((shape, dtype.name, bytes)))).  flax is not installed, so the byte
    layout is pinned on hand-built strings of that published format, then the state round-trips through a directory."""
    import msgpack
    import torch
    from bhnerf_b200 import network, optimization
    # known answer: a 2x2 float32 array as flax encodes it
    a = np.arange(4, dtype=np.float32).reshape(2, 2)
    want = msgpack.ExtType(1, msgpack.packb(([2, 2], 'float32', a.tobytes()), use_bin_type=True))
    assert optimization._np_to_ext(a) == want
    blob = msgpack.packb({'step': 7, 'params': {'w': want}}, use_bin_type=True)
    back = msgpack.unpackb(blob, ext_hook=optimization._flax_ext_hook, raw=False)
    assert back['step'] == 7 and np.array_equal(back['params']['w'], a) and back['params']['w'].dtype == np.float32
    # full state
    pred = network.NeRF_Predictor(8.0, 2.0, 8.0, 4.0)
    state = pred.init_state(pred.init_params(seed=3), num_iters=100, lr_init=1e-3, lr_final=1e-5, device='cpu')
    rng = np.random.default_rng(0)
    state.mu.copy_(torch.as_tensor(rng.normal(size=state.mu.numel()).astype(np.float32)))
    state.nu.copy_(torch.as_tensor(rng.uniform(size=state.nu.numel()).astype(np.float32)))
    state.step = 42
    tree = msgpack.unpackb(optimization.state_to_flax_bytes(state), ext_hook=optimization._flax_ext_hook, raw=False)
    assert set(tree) == {'step', 'params', 'opt_state'} and tree['step'] == 42
    assert tree['params']['MLP_0']['Dense_3']['kernel'].shape == (149, 128)
    assert set(tree['opt_state']) == {'0', '1'} and set(tree['opt_state']['0']) == {'count', 'mu', 'nu'}
    assert int(tree['opt_state']['0']['count']) == 42 and tree['opt_state']['0']['nu']['MLP_0']['Dense_4']['bias'].shape == (1,)
    d = str(tmp_path / 'ck')
    for step in (10, 20, 30):
        state.step = step
        optimization.save_checkpoint(d, state, step, keep=2)
    assert sorted(os.listdir(d)) == ['checkpoint_20', 'checkpoint_30']
    fresh = pred.init_state(pred.init_params(seed=9), num_iters=100, lr_init=1e-3, lr_final=1e-5, device='cpu', checkpoint_dir=d)
    assert fresh.step == 30 and torch.equal(fresh.flat, state.flat) and torch.equal(fresh.mu, state.mu) and torch.equal(fresh.nu, state.nu)
    # the earlier pickle payload is still readable
    d2 = str(tmp_path / 'ck2')
    optimization.save_checkpoint(d2, state, 5, fmt='pickle')
    old = pred.init_state(pred.init_params(seed=9), num_iters=100, device='cpu', checkpoint_dir=d2)
    assert torch.equal(old.flat, state.flat) and old.step == 30


def test_flax_numpy_scalar_ext_and_stray_checkpoint_files(tmp_path):
    """(1) flax packs a numpy SCALAR (e.g. state.step / opt_state count as np.int32) as ExtType(3, msgpack((shape=(),
    dtype, bytes))) -- the same triple as an ndarray, unpacked with arr[()]; a hand-built blob of that form must restore.
    (2) files that merely start with 'checkpoint_' (tmp files, foreign suffixes) are ignored by both the keep policy and the
    restore, and a save leaves no temporary file behind."""
    import msgpack
    import torch
    from bhnerf_b200 import network, optimization
    scal = msgpack.ExtType(3, msgpack.packb(([], 'int32', np.asarray(17, dtype=np.int32).tobytes()), use_bin_type=True))
    back = msgpack.unpackb(msgpack.packb({'step': scal}, use_bin_type=True), ext_hook=optimization._flax_ext_hook, raw=False)
    assert back['step'] == 17 and np.asarray(back['step']).dtype == np.int32 and np.ndim(back['step']) == 0
    pred = network.NeRF_Predictor(8.0, 2.0, 8.0, 4.0)
    state = pred.init_state(pred.init_params(seed=3), num_iters=100, device='cpu')
    tree = msgpack.unpackb(optimization.state_to_flax_bytes(state), ext_hook=lambda c, d: msgpack.ExtType(c, d), raw=False)
    tree['step'] = scal
    tree['opt_state']['0']['count'] = scal
    fresh = pred.init_state(pred.init_params(seed=9), num_iters=100, device='cpu')
    optimization.state_from_flax_bytes(fresh, msgpack.packb(tree, use_bin_type=True))
    assert fresh.step == 17 and torch.equal(fresh.flat, state.flat)
    d = str(tmp_path / 'ck')
    os.makedirs(d)
    for stray in ('checkpoint_tmp', 'checkpoint_5.orbax', 'checkpoint_'):
        open(os.path.join(d, stray), 'wb').write(b'junk')
    for step in (1, 2, 3):
        state.step = step
        optimization.save_checkpoint(d, state, step, keep=2)
    assert sorted(os.listdir(d)) == ['checkpoint_', 'checkpoint_2', 'checkpoint_3', 'checkpoint_5.orbax', 'checkpoint_tmp']
    again = pred.init_state(pred.init_params(seed=9), num_iters=100, device='cpu', checkpoint_dir=d)
    assert again.step == 3


def test_trainstep_eht_reference_signature_with_a_duck_typed_observation(monkeypatch):
    """TrainStep.eht(t_frames, obs, image_fov, image_size, chisqdata, pol, scale) -- the reference's signature
    (optimization.py:218-268): per-frame split, chisqdata per frame and polarization, pol axis stacked on axis 1 and
    squeezed, closure phases converted from ehtim's degrees.  ehtim itself is third party and absent: a stub module
    provides make_square and the observation is duck-typed."""
    import types
    from bhnerf_b200 import optimization
    calls = {}
    ehtim = types.ModuleType('ehtim'); image = types.ModuleType('ehtim.image')
    image.make_square = lambda obs, npix, fov: ('prior', npix, fov)
    ehtim.image = image
    monkeypatch.setitem(sys.modules, 'ehtim', ehtim); monkeypatch.setitem(sys.modules, 'ehtim.image', image)
    nt, V, P = 5, 7, 16
    rng = np.random.default_rng(0)

    class Obs:
        def split_obs(self, t_gather):
            calls['t_gather'] = t_gather
            return [('frame', i) for i in range(nt)]

    def chisqdata_vis(obs, prior, mask=[], pol='I'):
        k = 'IQU'.index(pol)
        return (np.full(V, obs[1] + 10 * k, dtype=np.complex64), np.full(V, 1.0 + k, dtype=np.float32),
                np.full((V, P), obs[1] + 100 * k, dtype=np.complex64))

    def chisqdata_cphase(obs, prior, mask=[], pol='I'):
        return (np.full(V, 90.0), np.full(V, 45.0), np.ones((3, V, P), dtype=np.complex64))

    t = np.linspace(0.0, 2.0, nt)
    ts = optimization.TrainStep.eht(t, Obs(), 1e-9, 4, chisqdata_vis)
    assert ts.dtype[0] == 'vis' and abs(calls['t_gather'] - 2.0 * 3600 / (nt + 1)) < 1e-9
    target, sigma, A, tf = ts.args[0].args
    assert target.shape == (nt, V) and sigma.shape == (nt, V) and A.shape == (nt, V, P) and A.dtype == np.complex64
    assert target[3, 0] == 3
    ts3 = optimization.TrainStep.eht(t, Obs(), 1e-9, 4, chisqdata_vis, pol=['I', 'Q', 'U'])
    target, sigma, A, tf = ts3.args[0].args
    assert target.shape == (nt, 3, V) and A.shape == (nt, 3, V, P) and sigma[0, 2, 0] == 3.0 and A[2, 1, 0, 0] == 102
    tsc = optimization.TrainStep.eht(t, Obs(), 1e-9, 4, chisqdata_cphase)
    target, sigma, A, tf = tsc.args[0].args
    assert tsc.dtype[0] == 'cphase' and np.allclose(target, np.pi / 2) and np.allclose(sigma, np.pi / 4) and A.shape == (nt, 3, V, P)
    with pytest.raises(AttributeError):
        optimization.TrainStep.eht(t, Obs(), 1e-9, 4, chisqdata_vis, pol='V')


def test_jax_ffi_binding_is_self_consistent():
    """jaxlib is absent, so the jax.ffi binding cannot run here -- but it must be CONSISTENT: the binding
    parses; every module / name it takes from this repository resolves; the scalar attributes it passes to ffi_call
    (JaxScene.attrs) are exactly the typed attributes the XLA-FFI shim binds and the fields of bhnerf_scene_t; the targets it
    registers are the handler symbols the shim defines; the C functions the shim calls are declared in the header."""
    import ast
    import re
    from bhnerf_b200 import _lib, jax_scene
    src = open(os.path.join(ROOT, 'integration', 'jax_binding.py')).read()
    tree = ast.parse(src)
    repo_imports = [n for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith('bhnerf_b200')]
    assert repo_imports, 'the binding should use bhnerf_b200.jax_scene'
    import importlib
    for n in repo_imports:
        mod = importlib.import_module(n.module)
        for a in n.names:
            assert hasattr(mod, a.name), '%s.%s does not exist' % (n.module, a.name)
    # attributes / methods used on `scene` objects exist on JaxScene
    used = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == 'scene'}
    have = set(dir(jax_scene.JaxScene)) | {'P', 'G', 'S', 'n_active', 'n_pad', 'image_shape', 'polarized', 'scene'}
    assert used and used <= have, used - have
    shim = open(os.path.join(ROOT, 'integration', 'xla_ffi_shim.cc')).read()
    bound = re.findall(r'\.Attr<(\w+)>\("(\w+)"\)', shim)
    want = [('int32_t' if ty == 'int32' else 'float', name) for name, ty in jax_scene.ATTRS]
    assert bound == want, (bound, want)
    hdr = open(os.path.join(ROOT, 'include', 'bhnerf_b200.h')).read()
    struct = re.search(r'typedef struct bhnerf_scene \{(.*?)\} bhnerf_scene_t;', hdr, flags=re.S).group(1)
    fields = re.findall(r'\b(\w+);', re.sub(r'/\*.*?\*/', '', struct, flags=re.S))
    assert fields == ['packed'] + [name for name, _ in jax_scene.ATTRS], fields
    # one Scene field per ctypes field as well
    assert [f[0] for f in _lib.Scene._fields_] == fields
    targets = set(re.findall(r"_lib\.(Bhnerf\w+)", src))
    assert targets == set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),', shim)) and len(targets) == 2
    for fn in set(re.findall(r'\b(bhnerf_[a-z_]+)\(', shim)):
        assert fn in _lib.header_functions(), fn
    assert 'consts' not in shim and 'consts' not in src        # the struct-in-a-float-span attribute is gone
