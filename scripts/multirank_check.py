"""Correctness of the multi-GPU path (one process per GPU, frames sharded, NCCL all-reduce of the gradient, 1/ndev folded
into Adam = jax.lax.pmean + apply_gradients, bhnerf/network.py:620-621) against the oracle golden:
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multirank_check.py
Every rank takes its slice of the golden case's 4 frames through the reference-facing TrainStep; the parameters after one
step must equal Adam applied to the MEAN of the per-rank gradients = (full-batch golden gradient) / world_size."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bhnerf_b200 import network, optimization  # noqa: E402
from oracle import bhnerf_oracle as O  # noqa: E402  (checker only)

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
G = os.path.join(ROOT, 'tests', 'golden')
geo = np.load(os.path.join(G, 'kerr_a0.2_i60_16x16x32.npz'))
for case, kind in (('case_image_full', 'full'), ('case_lc_QU', 'lc')):
    d = np.load(os.path.join(G, case + '.npz'))
    J = d['J'] if 'J' in d.files else 1.0
    pred = network.NeRF_Predictor(float(d['scale']), float(d['rmin']), float(d['rmax']), float(d['z_width']))
    state = pred.init_state(network.unflatten_params(d['params_flat']), num_iters=100, lr_init=1e-3, lr_final=1e-5)
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=J, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    sig = d['sigma'] if 'sigma' in d.files else 1.0
    ts = optimization.TrainStep.image(d['t_frames'], d['target'], sigma=sig, dtype=kind)
    loss, state, images = ts(state, rt, np.arange(4))          # shard() hands this rank its frames
    total = optimization._allreduce_scalar(loss)
    want, _, _ = O.adam_step(d['params_flat'].astype(np.float64), d['grads'] / world, np.zeros(55169), np.zeros(55169), 0,
                             1e-3, 1e-5, 100)
    got = state.flat.cpu().numpy().astype(np.float64)
    upd_err = np.abs((got - d['params_flat']) - (want - d['params_flat'])).max() / 1e-3
    same = torch.stack([state.flat.clone() for _ in range(1)])
    ref = state.flat.clone(); dist.broadcast(ref, src=0)
    if rank == 0:
        print('%s  world=%d  frames/rank=%d  sum of per-rank losses rel err %.2e  Adam update err %.2e of lr  ranks identical: %s'
              % (case, world, images.shape[0], abs(total - float(d['loss'])) / abs(float(d['loss'])), upd_err,
                 bool(torch.equal(ref, state.flat))), flush=True)
    ok = torch.tensor([int(upd_err < 2e-2 and torch.equal(ref, state.flat))], device='cuda')
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    assert ok.item() == 1

# ---- ray sharding: 3 frames on `world` ranks cannot be split by frames -> every rank renders all 3
# frames for its block of rays; partial lightcurves / visibilities are all-reduced before the loss, gradients (SUM) after.
# The result must equal ONE device's step on the 3 frames: oracle value_and_grad + Adam, live on the CPU (checker only).
idx = np.arange(3)
for case, kind in (('case_image_full', 'full'), ('case_lc_QU', 'lc'), ('case_vis', 'vis')):
    d = np.load(os.path.join(G, case + '.npz'))
    J = d['J'] if 'J' in d.files else 1.0
    prd = dict(scale=float(d['scale']), rmin=float(d['rmin']), rmax=float(d['rmax']), z_width=float(d['z_width']))
    pred = network.NeRF_Predictor(prd['scale'], prd['rmin'], prd['rmax'], prd['z_width'])
    state = pred.init_state(network.unflatten_params(d['params_flat']), num_iters=100, lr_init=1e-3, lr_final=1e-5)
    rt = OrderedDict(coords=geo['coords'], Omega=geo['Omega'], J=J, g=geo['g'], dtau=geo['dtau'], Sigma=geo['Sigma'],
                     t_start_obs=float(d['t_start_obs']), t_geos=geo['t_geos'], t_injection=float(d['t_injection']))
    params = O.unflatten_params(d['params_flat'])
    if kind == 'vis':
        ts = optimization.TrainStep.eht_arrays(d['t_frames'], d['target'], d['sigma'], d['A'], dtype='vis')
        ref = O.value_and_grad(params, 'eht', 'vis', d['target'][idx], d['sigma'][idx], d['A'][idx], d['t_frames'][idx], dict(rt), prd)
    else:
        sig = d['sigma'] if 'sigma' in d.files else np.ones_like(d['target'])
        ts = optimization.TrainStep.image(d['t_frames'], d['target'], sigma=sig, dtype=kind)
        ref = O.value_and_grad(params, 'image', kind, d['target'][idx], sig[idx], np.zeros_like(d['target'][idx]),
                               d['t_frames'][idx], dict(rt), prd)
    assert optimization.ray_sharded(len(idx))
    loss, state, images = ts(state, rt, idx)
    want, _, _ = O.adam_step(d['params_flat'].astype(np.float64), ref['grads'], np.zeros(55169), np.zeros(55169), 0, 1e-3, 1e-5, 100)
    got = state.flat.cpu().numpy().astype(np.float64)
    upd_err = np.abs((got - d['params_flat']) - (want - d['params_flat'])).max() / 1e-3
    img_err = np.abs(images.cpu().numpy().reshape(ref['images'].shape) - ref['images']).max() / np.abs(ref['images']).max()
    loss_err = abs(loss.item() - ref['loss']) / abs(ref['loss'])
    chk = state.flat.clone(); dist.broadcast(chk, src=0)
    good = bool(upd_err < 2e-2 and img_err < 1e-4 and loss_err < 2e-4 and torch.equal(chk, state.flat))
    if rank == 0:
        print('ray-sharded %s  world=%d  frames=3 on every rank  loss rel err %.2e  image err %.2e  Adam update err %.2e of lr  %s'
              % (case, world, loss_err, img_err, upd_err, 'ok' if good else 'FAILED'), flush=True)
    ok = torch.tensor([int(good)], device='cuda')
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    assert ok.item() == 1
dist.destroy_process_group()
