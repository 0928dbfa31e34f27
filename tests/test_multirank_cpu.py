"""world_size-2 gloo tests (CPU) of the host logic around the one exchange the path has: frame sharding
(bhnerf/optimization.py:209-216, :360-362) and all-reduce-mean of the gradient (jax.lax.pmean, network.py:620)
with the 1/ndev folded into the optimiser update."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from bhnerf_b200 import network, optimization
        # --- frame sharding: every rank takes its contiguous slice of the sampled batch (pmap in_axes=0)
        nt = 12
        t_frames = np.linspace(0.0, 1.0, nt).astype(np.float32)
        target = np.arange(nt * 3, dtype=np.float32).reshape(nt, 3)
        idx = np.arange(8)
        mine = optimization.shard(target[idx])
        assert mine.shape == (8 // world, 3), mine.shape
        np.testing.assert_array_equal(mine, target[idx][rank * 4:(rank + 1) * 4])
        assert optimization.device_count() == world
        # a batch not divisible by the device count (the reference raises, optimization.py:362) stays whole on every
        # rank and the step shards rays instead
        assert optimization.ray_sharded(7) and not optimization.ray_sharded(8)
        assert optimization.shard(target[:7]).shape == (7, 3)
        # --- all-reduce-mean + update: the same step on both ranks must leave identical parameters, equal to a
        #     single-rank step with the mean gradient
        g = torch.full((16,), float(rank + 1))

        class _State:                                   # minimal stand-in: records what apply_gradients sees
            def apply_gradients(self, grads, grad_scale=1.0):
                self.seen = grads.clone() * grad_scale
        st = _State()
        network._pmean_and_apply(st, g)
        want = torch.full((16,), (1.0 + 2.0) / 2.0)
        assert torch.allclose(st.seen, want), st.seen
        # --- scalar loss reduction / frame gathering used by total_movie_loss
        tot = optimization._allreduce_scalar(torch.tensor(float(rank + 1)))
        assert abs(float(tot) - 3.0) < 1e-6
        frames = optimization._allgather_frames(torch.full((2, 4), float(rank)))
        assert frames.shape == (4, 4) and float(frames[0, 0]) == 0.0 and float(frames[3, 0]) == 1.0
        out[rank] = 'ok'
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 'ok', 1: 'ok'}
