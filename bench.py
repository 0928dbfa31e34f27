#!/usr/bin/env python
"""Benchmark of the bhnerf render/train hot path (BASELINE.json metric: geodesic samples/s, fwd+bwd train step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--kernels auto|simt|tc]
                  [--workload cfg2_lp_flare|cfg1_tutorial3|cfg3_ngeht|cfg4_highres|cfg5_alma] [--frames F]
                  [--scaling weak|strong] [--ensemble] [--vis-head matrix|separable]

One "step" = one fused fwd+bwd train step (render -> loss -> parameter gradient [-> all-reduce-mean over ranks]; the
e2e leg adds Adam and goes through the reference-facing API with HOST buffers) over ALL frames of the workload.
Default workload = BASELINE.json configs[1] (128x128 rays x 128 samples x 100 frames, Q/U lightcurve loss), which fits one
GPU.  cfg3_ngeht times the visibility step (render -> per-frame DFT to V baselines -> chi^2 -> pull-back) and reports the
HBM roofline of the visibility head next to the tensor roofline of the render kernels; with --vis-head separable the same
Fourier kernel is handed over as baselines (u, v) and the head runs as a separable DFT on the tensor cores.

Several ranks (one process per GPU, launched by torchrun):
  --scaling weak   (default) every rank renders its own frames of the workload: global batch = frames x N; the ranks
                   exchange the 220 KB gradient with ONE all-reduce per step (bhnerf_allreduce_mean: NCCL on the kernel
                   stream, below the C ABI), as the reference's pmap/pmean does (network.py:620).
  --scaling strong the workload's frames are split over the ranks (cfg4_highres: 128 frames -> 16 per GPU on 8 GPUs).
  --ensemble       BASELINE.json configs[4]: one independent model per GPU (its own inclination), NO collective.
`value` counts DENSE samples (frames x rays x samples_per_ray -- the reference evaluates all of them); the evaluated
count after dead-sample culling is reported next to it and is what the rooflines use."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_FWD, FLOP_BWD = 109312, 207872          # per evaluated sample
NCAT = 7
CATS = ['render_fwd', 'render_bwd', 'wgrad', 'heads', 'misc', 'comm', 'vis_head']


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--kernels', default=os.environ.get('BHNERF_IMPL', 'auto'))
    ap.add_argument('--workload', default=None)
    ap.add_argument('--frames', type=int, default=None, help='override the number of frames (debug)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--ensemble', action='store_true', help='cfg5: one independent model per GPU, no collective')
    ap.add_argument('--vis-head', default='matrix', choices=['matrix', 'separable'],
                    help="cfg3: 'matrix' = the explicit complex64 A of loss_fn_eht (HBM-bound batched GEMV); 'separable' = "
                         "the same Fourier kernel handed over as baselines (u, v): separable DFT on the tensor cores")
    ap.add_argument('--max-workspace-gb', type=float, default=40.0)
    ap.add_argument('--cpu-frames', type=int, default=None, help='frames in the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    a = ap.parse_args()
    if a.workload is None:
        a.workload = 'cfg5_alma' if a.ensemble else ('cfg4_highres' if a.scaling == 'strong' else 'cfg2_lp_flare')
    return a


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): one persistent
    `nvidia-smi -lms 50` loop started before the warm-up; rows are time-stamped and filtered to [t0, t1]."""
    Q = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.lock = threading.Lock()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            with self.lock:
                self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t0=None, t1=None):
        with self.lock:
            rows = list(self.rows)
        inside = [r for (t, r) in rows if t0 is None or (t0 <= t <= t1 + 0.05)]
        where = 'timed region'
        if len(inside) < 2:                        # region shorter than the sampling period: use the loaded window
            inside = [r for (t, r) in rows if t0 is None or t >= t0 - 2.0]
            where = 'warm-up + timed region'
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in inside:
            try:
                sm.append(float(r[2])); mx.append(float(r[3])); pw.append(float(r[4]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w': float(np.median(pw)) if pw else None, 'reasons': sorted(reasons), 'samples': len(sm), 'window': where}


def cpu_reference_leg(cfg_name, n_frames, threads=None):
    """The reference's CPU path: all-float32 dense restatement (oracle/bhnerf_oracle.py, kind 'port'; JAX is not
    installable offline) of gradient_step_image / gradient_step_eht -- warp, posenc, MLP on EVERY sample (no culling),
    loss, autograd backward -- one frame at a time on the host cores.  Returns (dense samples/s, seconds, cores)."""
    import torch
    from bhnerf_b200 import synthetic
    from oracle import bhnerf_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    c = synthetic.make_config(cfg_name, nt=n_frames)
    params = O.unflatten_params(synthetic.trained_like_flat_params(7))
    kind = c['cfg']['loss']
    dense = c['P'] * c['G']
    t0 = time.perf_counter()
    for b in range(n_frames):
        if kind == 'vis':
            O.value_and_grad(params, 'eht', 'vis', c['target'][b:b + 1], c['sigma'][b:b + 1], c['Amat'][b:b + 1],
                             c['t_frames'][b:b + 1], c['rt'], c['predictor'], dtype=torch.float32)
        else:
            tgt = c['target'][b:b + 1].reshape((1, c['S'], c['A'], c['B']) if kind == 'full' else (1, c['S']))
            sig = c['sigma'][b:b + 1].reshape(tgt.shape)
            O.value_and_grad(params, 'image', kind, tgt, sig, np.zeros_like(tgt), c['t_frames'][b:b + 1], c['rt'],
                             c['predictor'], dtype=torch.float32)
    dt = time.perf_counter() - t0
    return dense * n_frames / dt, dt, threads


def workload_config(args, c):
    return {'workload': args.workload, 'rays': c['n'] * c['n'], 'samples_per_ray': c['G'], 'frames': c['nt'],
            'stokes': c['S'], 'loss': c['loss']}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from bhnerf_b200 import synthetic
    c = dict(synthetic.CONFIGS[args.workload])
    if args.frames:
        c['nt'] = args.frames
    big = c['n'] * c['n'] * c['G'] > 5e6
    nfr = args.cpu_frames or (1 if big else 4)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_leg(args.workload, 1)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, cores = cpu_reference_leg(args.workload, nfr)
        vals.append(v); secs.append(dt)
    v = float(np.mean(vals))
    line = {'metric': 'geodesic samples/s, fwd+bwd train step', 'value': v, 'unit': 'dense samples/s', 'n_gpus': args.gpus,
            'impl': 'reference', 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(secs)),
            'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, c),
            'cpu_baseline': {'value': v, 'unit': 'dense samples/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d of %d frames per step, float32 dense torch-CPU restatement of the '
                                       'reference JAX path (JAX not installable offline)' % (nfr, c['nt'])},
            'e2e': {'value': v, 'unit': 'dense samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from bhnerf_b200 import _lib, engine, network, optimization, synthetic
    from collections import OrderedDict

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    impl = engine.resolve_impl(args.kernels)
    impl_name = 'tc' if impl == engine.IMPL_TC else 'simt'
    exchange = world > 1 and not args.ensemble
    if exchange:
        engine.comm(dev)                                         # collective creation of the C-ABI communicator

    # ---- workload of this rank ----
    base = synthetic.CONFIGS[args.workload]
    span = base['t_span']
    nt_all = args.frames or base['nt']
    if args.scaling == 'strong' and world > 1 and not args.ensemble:
        assert nt_all % world == 0, 'strong scaling: %d frames do not divide over %d ranks' % (nt_all, world)
        c = synthetic.make_config(args.workload, seed=0, nt=nt_all)           # same movie everywhere ...
        sl = slice(rank * (nt_all // world), (rank + 1) * (nt_all // world))  # ... this rank's frames of it
        for k in ('t_frames', 'target', 'sigma', 'offset') + (('Amat',) if 'Amat' in c else ()):
            c[k] = np.ascontiguousarray(c[k][sl])
        scaling, par = 'strong', 'frames of one movie split over %d ranks, gradient all-reduce-mean' % world
    elif args.ensemble:
        # one independent model per GPU: its own inclination (scripts/Fit_ALMA_LP_Apr11_SgrA_Flare.py:83-115 sweeps
        # inclinations, one run each), its own parameters and targets; nothing is exchanged
        inc = float(np.arange(4, 82, 2)[(rank * 5) % 39])
        c = synthetic.make_config(args.workload, seed=rank, nt=nt_all, inc=inc)
        scaling, par = 'weak', 'ensemble: %d independent models (inclinations), one per GPU, no collective' % world
    else:
        c = synthetic.make_config(args.workload, seed=rank, frame_offset=rank * (span[1] - span[0]) * 0.013, nt=nt_all)
        scaling, par = 'weak', 'frames x%d ranks, gradient all-reduce-mean' % world
    cfg, rt, pr = c['cfg'], c['rt'], c['predictor']
    kind = cfg['loss']
    Bt, S, P, Gs = len(c['t_frames']), c['S'], c['P'], c['G']
    dense_per_step = Bt * P * Gs
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    rta = OrderedDict(coords=rt['coords'], Omega=rt['Omega'], J=rt['J'], g=rt['g'], dtau=rt['dtau'], Sigma=rt['Sigma'],
                      t_start_obs=rt['t_start_obs'], t_geos=rt['t_geos'], t_injection=rt['t_injection'])
    scene = network._scene_for(pred, *[rta[k] for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs',
                                                        't_geos', 't_injection')], 'hr', device=dev)
    eval_per_step = Bt * scene.n_active
    params0 = network.unflatten_params(synthetic.trained_like_flat_params(7 + (rank if args.ensemble else 0)))
    state = pred.init_state(params0, num_iters=10000, lr_init=1e-4, lr_final=1e-6, device=dev)
    ws_cap = int(args.max_workspace_gb * 2 ** 30)

    # resident inputs for the kernel-path timing
    tf_d = torch.as_tensor(c['t_frames'], device=dev)
    separable = kind == 'vis' and args.vis_head == 'separable'
    if kind == 'vis':
        if separable:
            A_d = network.SeparableDFT.from_fov(c['uv'], scene.image_shape, c['cfg']['fov'])
            V = c['uv'].shape[1]
        else:
            A_d = torch.as_tensor(c['Amat'], device=dev)
            V = A_d.shape[1]
        tgt_d = torch.as_tensor(c['target'], device=dev); sig_d = torch.as_tensor(c['sigma'], device=dev)
        Bc = engine.frames_per_chunk(scene, Bt, impl, max_workspace=ws_cap)

        def step_resident():
            grads, loss = None, None
            for b0 in range(0, Bt, Bc):
                sl = slice(b0, min(b0 + Bc, Bt))
                images, e, acts = engine.render_fwd(scene, state.flat, tf_d[sl], impl, save_acts=True)
                l, _, dI = network._vis_head(A_d[sl], images.reshape(-1, 1, P), tgt_d[sl], sig_d[sl], 1.0, 'vis')
                dI = dI.reshape(images.shape)
                g = engine.render_bwd(scene, state.flat, tf_d[sl], dI, e, acts, impl, max_workspace=ws_cap)
                grads = g if grads is None else engine.add_inplace(grads, g)
                loss = l if loss is None else engine.add_inplace(loss, l)
            if exchange:
                engine.allreduce_mean(grads)
            return loss
    else:
        tgt_d = torch.as_tensor(c['target'], device=dev); sig_d = torch.as_tensor(c['sigma'], device=dev)
        off_d = torch.as_tensor(c['offset'], device=dev)
        out = (torch.empty(1, device=dev), torch.empty((Bt, S, P), device=dev), torch.empty(55169, device=dev))

        def step_resident():
            loss, images, grads = engine.train_step_image(scene, state.flat, tf_d, tgt_d, sig_d, off_d, 1.0, kind, impl,
                                                          max_workspace=ws_cap, out=out)
            if exchange:
                engine.allreduce_mean(grads)
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # ---- timed region 1: resident inputs (value) + live per-kernel event timing ----
    lib.bhnerf_profile_begin()
    t_region0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_rank = ms
    cat_ms = (ctypes.c_double * NCAT)(); cat_sc = (ctypes.c_int64 * NCAT)(); cat_ln = (ctypes.c_int64 * NCAT)()
    lib.bhnerf_profile_end(cat_ms, cat_sc, cat_ln)
    clocks = sampler.summary(t_region0, time.time())

    # ---- timed region 2: end to end through the reference-facing API with HOST buffers ----
    if kind == 'vis':
        ts = optimization.TrainStep.eht_arrays(c['t_frames'], c['target'], c['sigma'], A_d if separable else c['Amat'], dtype='vis')
        step_fn = network.gradient_step_eht
    else:
        ts = optimization.TrainStep.image(c['t_frames'], c['target'].reshape((Bt, S) if kind == 'lc' else (Bt, S, P)),
                                          sigma=c['sigma'].reshape((Bt, S) if kind == 'lc' else (Bt, S, P)), dtype=kind)
        step_fn = network.gradient_step_image
    # pinned host staging of the per-step inputs (target, sigma, offset | A, t_frames)
    host_args = [a if isinstance(a, network.SeparableDFT) else torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
                 for a in ts.args[0].args]
    h2d = sum(a.uv.nbytes if isinstance(a, network.SeparableDFT) else a.numel() * a.element_size() for a in host_args)
    if args.ensemble and world > 1:          # independent models: the step must not look for peers
        network._dist = lambda: None
        engine._dist_world = lambda: (None, 0, 1)

    def step_e2e():
        # the API takes the HOST buffers, as the reference's does (TemporalBatchedArgs hands numpy arrays to the pmap'd step);
        # every host -> device copy of the step (h2d bytes below) happens inside the call: the image steps stage their inputs
        # into one pinned buffer, the eht step streams A under the render of the chunk it belongs to
        loss, st, images = step_fn(state, 'hr', kind, *host_args, *rta.values(), 1.0, impl=impl)
        return float(loss.item())                      # device -> host read of the step's result

    for _ in range(3):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        last_loss = step_e2e()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    sampler.stop()

    # max over ranks; per-rank breakdown (kernel time by category, all-reduce time incl. the wait for the slowest rank)
    per_rank = None
    if world > 1:
        dist_mod = dist
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        mine = torch.tensor([ms_rank / args.steps] + [cat_ms[i] / args.steps for i in range(NCAT)] +
                            [float(eval_per_step), float(dense_per_step), float(clocks['sm_mhz'] or 0.0)],
                            device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist_mod.all_gather(allr, mine)
        rows = [r.tolist() for r in allr]
        per_rank = [{'rank': i, 'ms_per_step': r[0], 'kernels_ms': {CATS[k]: r[1 + k] for k in range(NCAT) if r[1 + k] > 0},
                     'sm_mhz': r[-1]} for i, r in enumerate(rows)]
        eval_total = sum(r[1 + NCAT] for r in rows)
        dense_total = sum(r[2 + NCAT] for r in rows)
    else:
        eval_total, dense_total = float(eval_per_step), float(dense_per_step)
    value = dense_total * args.steps / (ms * 1e-3)
    e2e_value = dense_total * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        tensor_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md)'
        ms_by = {n: cat_ms[i] for i, n in enumerate(CATS)}
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tc_on = impl == engine.IMPL_TC
        # ---- bytes.  SURVEY.md s8d's ALGORITHMIC bytes of the render = the packed sample stream, P_active*(24+4S) per
        # step (frame independent, L2 resident by design: 8 MB for cfg2).  What the kernels really move through HBM is the
        # design's own save/restore of activations between forward and backward: per evaluated
        # sample-frame the forward writes h0..h3 (4 x 256 B fp16) + features 48 + ReLU masks 64 + e 4, the fused backward
        # reads them back + d loss/d o 4.  Both are reported, and their ratio, instead of calling the second "algorithmic".
        stream_bytes = scene.n_active * (24 + 4 * S)
        save_fwd, save_bwd = (1140, 1144) if tc_on else (2136, 4184)
        spec = {'render_fwd': dict(bound='tensor', flop=FLOP_FWD, design_bytes=save_fwd),
                'render_bwd': dict(bound='tensor', flop=FLOP_BWD if tc_on else 49280 * 2, design_bytes=save_bwd),
                'wgrad': dict(bound='tensor', flop=54656 * 2, design_bytes=4180)}
        try:
            ncu_traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        except Exception:
            ncu_traffic = {}
        kernels = {}
        for n, sp in spec.items():
            launches = max(int(cat_sc[CATS.index(n)]), 1)
            per_launch_s = ms_by[n] * 1e-3 / launches
            eval_per_launch = eval_per_step * args.steps / launches
            if per_launch_s <= 0:
                continue
            tf = eval_per_launch * sp['flop'] / per_launch_s / 1e12
            design_gbs = eval_per_launch * sp['design_bytes'] / per_launch_s / 1e9
            tr = ncu_traffic.get(n + '_' + impl_name, {}).get('dram_bytes_per_eval_sample')
            kernels[n + '_' + impl_name] = {
                'bound': 'tensor', 'avg_launch_ms': per_launch_s * 1e3, 'ms_per_step': ms_by[n] / args.steps,
                'achieved': tf, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': tf / tensor_peak, 'algorithmic_tflops': tf,
                'design_hbm_gbs': design_gbs, 'design_hbm_frac': design_gbs / hbm_peak,
                'design_bytes_per_evaluated_sample': sp['design_bytes'],
                'traffic': tr * eval_per_launch if tr else None}
        if separable and ms_by['vis_head'] > 0:
            # separable DFT on the tensor cores: nothing of size V x P is read; algorithmic work = one real-by-complex GEMM per
            # direction (cos and sin parts): 2 directions x 2 parts x 2 V P flop per frame
            head_flop = 8.0 * V * P * Bt
            head_s = ms_by['vis_head'] * 1e-3 / args.steps
            kernels['vis_dft'] = {'bound': 'tensor', 'ms_per_step': head_s * 1e3, 'achieved': head_flop / head_s / 1e12,
                                  'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': head_flop / head_s / 1e12 / tensor_peak,
                                  'traffic': None, 'explicit_matrix_hbm_floor_ms': 2 * 8.0 * V * P * Bt / hbm_peak / 1e6,
                                  'note': 'latency-bound at this size: the whole head is %.0f MFLOP; the point is the %.1f GB of '
                                          'A it does not read' % (head_flop / 1e6, 2 * 8.0 * V * P * Bt / 1e9),
                                  'launches_per_step': int(cat_sc[CATS.index('vis_head')]) // args.steps}
        elif kind == 'vis' and ms_by['vis_head'] > 0:
            # the visibility head is a batched GEMV over a per-frame complex64 A: algorithmic bytes = 8*V*P per frame per
            # pass, two passes (A I and A^H d_vis) -- SURVEY.md s8d; the second pass is served by the L2 (bhnerf_vis_head)
            head_bytes = 2 * 8.0 * V * P * Bt + 2 * 4.0 * P * Bt
            head_s = ms_by['vis_head'] * 1e-3 / args.steps
            gbs = head_bytes / head_s / 1e9
            kernels['vis_head'] = {'bound': 'hbm', 'ms_per_step': head_s * 1e3, 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                                   'frac': gbs / hbm_peak, 'algorithmic_bytes_per_step': head_bytes,
                                   'note': 'achieved = algorithmic bytes (two passes over A: A I and A^H d_vis) / head time',
                                   'launches_per_step': int(cat_sc[CATS.index('vis_head')]) // args.steps}
        kernels = {n: k for n, k in kernels.items() if k['ms_per_step'] > 0}
        dom = max((n for n in kernels if kernels[n]['bound'] == 'tensor'), key=lambda n: kernels[n]['ms_per_step'])
        kd = kernels[dom]
        step_tf = eval_per_step * (FLOP_FWD + FLOP_BWD) * args.steps / (ms_rank * 1e-3) / 1e12
        design_bytes_step = eval_per_step * (save_fwd + save_bwd)
        roofline = {'bound': kd['bound'], 'kernel': dom, 'achieved': kd['achieved'], 'peak': kd['peak'], 'unit': kd['unit'],
                    'frac': kd['frac'], 'traffic': kd['traffic'], 'peak_source': peak_src,
                    'avg_launch_ms': kd['avg_launch_ms'], 'kernels': kernels,
                    'kernel_ms_per_step': {n: ms_by[n] / args.steps for n in CATS},
                    'step_algorithmic_tflops': step_tf, 'step_frac_of_tensor_peak': step_tf / tensor_peak,
                    'bytes': {'algorithmic_sample_stream_per_step': stream_bytes,
                              'design_save_restore_per_step': design_bytes_step,
                              'design_over_algorithmic': design_bytes_step / max(stream_bytes, 1),
                              'design_hbm_gbs_over_step': design_bytes_step * args.steps / (ms_rank * 1e-3) / 1e9,
                              'design_hbm_frac_over_step': design_bytes_step * args.steps / (ms_rank * 1e-3) / 1e9 / hbm_peak,
                              'note': 'the sample stream (SURVEY s8d) is L2-resident and reused by every frame; the HBM traffic '
                                      'of the step is the save/restore of activations the design chose over recomputation'}}
        cpu = None
        if not args.no_cpu_baseline:
            big = P * Gs > 5e6
            nfr = args.cpu_frames or (1 if big else (4 if kind == 'vis' else 8))
            v, dt, cores = cpu_reference_leg(args.workload, nfr)
            cpu = {'value': v, 'unit': 'dense samples/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d of %d frames (%.1f s), float32 dense torch-CPU restatement of the reference JAX path'
                             % (nfr, Bt, dt)}
        wc = workload_config(args, dict(base, nt=nt_all))
        wc.update({'frames_per_gpu': Bt, 'mlp': '4x128 relu + skip, posenc deg 3', 'kernels': impl_name, 'parallelism': par,
                   'active_fraction': scene.n_active / (P * Gs),
                   'l2_policy': 'inputs larger than L2: each step streams the per-frame activation workspace '
                                '(>> 126 MB); the packed geodesic set is L2-resident by design'})
        if kind == 'vis':
            wc['visibilities_per_frame'] = V
        line = {
            'metric': 'geodesic samples/s, fwd+bwd train step', 'value': value, 'unit': 'dense samples/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
            'dtype': 'f16x3 (fwd) / f16 scaled cotangents (bwd) split operands, f32 accumulate (tcgen05)' if tc_on else 'f32 (FFMA)',
            'data': 'synthetic', 'config': wc,
            'evaluated_samples_per_s': eval_total * args.steps / (ms * 1e-3),
            'e2e': {'value': e2e_value, 'unit': 'dense samples/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'bhnerf_b200.network.%s (train step + all-reduce + Adam)' % step_fn.__name__},
            'gpu_launches': int(sum(cat_ln)),
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu, 'last_loss': last_loss,
        }
        if per_rank is not None:
            line['per_rank'] = per_rank
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
