#!/usr/bin/env python
"""Benchmark of the bhnerf render/train hot path (BASELINE.json metric: geodesic samples/s, fwd+bwd train step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--kernels auto|simt|tc]
                  [--workload cfg2_lp_flare] [--frames F]

One "step" = one fused fwd+bwd train step (render -> loss -> parameter gradient [-> all-reduce-mean -> Adam in
the e2e leg, where network.gradient_step_image replays the captured CUDA graph of the step at N=1]) over ALL frames
of the workload.  Default workload = BASELINE.json configs[1]
(128x128 rays x 128 samples x 100 frames, Q/U lightcurve loss), which fits one GPU.  With N ranks every rank
renders its own 100 frames (weak scaling: global batch = 100*N frames) and the ranks exchange the 220 KB
gradient with one NCCL all-reduce per step, as the reference's pmap/pmean does (network.py:620).
`value` counts DENSE samples (frames*rays*samples_per_ray -- the reference evaluates all of them); the
evaluated count after dead-sample culling is reported next to it and is what the roofline uses."""
import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--kernels', default=os.environ.get('BHNERF_IMPL', 'auto'))
    ap.add_argument('--workload', default='cfg2_lp_flare')
    ap.add_argument('--frames', type=int, default=None, help='override the number of frames (debug)')
    ap.add_argument('--max-workspace-gb', type=float, default=40.0)
    ap.add_argument('--cpu-frames', type=int, default=None, help='frames in the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


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
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in inside:
            try:
                sm.append(float(r[2])); mx.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': where}


def cpu_reference_leg(cfg_name, n_frames, threads=None):
    """The reference's CPU path: all-float32 dense restatement (oracle/bhnerf_oracle.py, kind 'port'; JAX is not
    installable offline) of gradient_step_image -- warp, posenc, MLP on EVERY sample (no culling), loss, autograd
    backward -- one frame at a time on the host cores.  Returns (dense samples/s, seconds, cores)."""
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


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from bhnerf_b200 import synthetic
    c = synthetic.CONFIGS[args.workload]
    nfr = args.cpu_frames or 4
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_leg(args.workload, 1)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, cores = cpu_reference_leg(args.workload, nfr)
        vals.append(v); secs.append(dt)
    v = float(np.mean(vals))
    line = {'metric': 'geodesic samples/s, fwd+bwd train step', 'value': v, 'unit': 'dense samples/s', 'n_gpus': args.gpus,
            'impl': 'reference', 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean(secs)),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'rays': c['n'] * c['n'], 'samples_per_ray': c['G'], 'frames': c['nt'],
                       'stokes': c['S'], 'loss': c['loss']},
            'cpu_baseline': {'value': v, 'unit': 'dense samples/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d of %d frames per step, float32 dense torch-CPU restatement of the '
                                       'reference JAX path (JAX not installable offline)' % (nfr, c['nt'])},
            'e2e': {'value': v, 'unit': 'dense samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)
    import ctypes
    import torch
    import torch.distributed as dist
    from bhnerf_b200 import _lib, constants, engine, network, optimization, synthetic

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

    # ---- workload: every rank renders its own frames of the same scene (weak scaling) ----
    span = synthetic.CONFIGS[args.workload]['t_span']
    c = synthetic.make_config(args.workload, seed=rank, frame_offset=rank * (span[1] - span[0]) * 0.013, nt=args.frames)
    cfg, rt, pr = c['cfg'], c['rt'], c['predictor']
    kind = cfg['loss']
    if kind == 'vis':
        raise SystemExit('bench.py times the image/lightcurve train step; use tests for the visibility head')
    Bt, S, P, Gs = len(c['t_frames']), c['S'], c['P'], c['G']
    dense_per_step = Bt * P * Gs
    pred = network.NeRF_Predictor(pr['scale'], pr['rmin'], pr['rmax'], pr['z_width'])
    from collections import OrderedDict
    rta = OrderedDict(coords=rt['coords'], Omega=rt['Omega'], J=rt['J'], g=rt['g'], dtau=rt['dtau'], Sigma=rt['Sigma'],
                      t_start_obs=rt['t_start_obs'], t_geos=rt['t_geos'], t_injection=rt['t_injection'])
    scene = network._scene_for(pred, *[rta[k] for k in ('coords', 'Omega', 'J', 'g', 'dtau', 'Sigma', 't_start_obs',
                                                        't_geos', 't_injection')], 'hr', device=dev)
    eval_per_step = Bt * scene.n_active
    params0 = network.unflatten_params(synthetic.trained_like_flat_params(7))
    state = pred.init_state(params0, num_iters=10000, lr_init=1e-4, lr_final=1e-6, device=dev)
    ws_cap = int(args.max_workspace_gb * 2 ** 30)

    # resident inputs for the kernel-path timing
    tf_d = torch.as_tensor(c['t_frames'], device=dev)
    tgt_d = torch.as_tensor(c['target'], device=dev); sig_d = torch.as_tensor(c['sigma'], device=dev)
    off_d = torch.as_tensor(c['offset'], device=dev)
    out = (torch.empty(1, device=dev), torch.empty((Bt, S, P), device=dev), torch.empty(55169, device=dev))

    def step_resident():
        loss, images, grads = engine.train_step_image(scene, state.flat, tf_d, tgt_d, sig_d, off_d, 1.0, kind, impl,
                                                      max_workspace=ws_cap, out=out)
        if world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM)
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
    cat_ms = (ctypes.c_double * 5)(); cat_sc = (ctypes.c_int64 * 5)(); cat_ln = (ctypes.c_int64 * 5)()
    lib.bhnerf_profile_end(cat_ms, cat_sc, cat_ln)
    clocks = sampler.summary(t_region0, time.time())

    # ---- timed region 2: end to end through the reference-facing API with HOST buffers ----
    ts = optimization.TrainStep.image(c['t_frames'], c['target'].reshape((Bt, S) if kind == 'lc' else (Bt, S, P)),
                                      sigma=c['sigma'].reshape((Bt, S) if kind == 'lc' else (Bt, S, P)), dtype=kind)
    # pinned host staging of the per-step inputs (target, sigma, offset, t_frames)
    host_args = [torch.as_tensor(np.ascontiguousarray(a)).pin_memory() for a in ts.args[0].args]
    h2d = sum(a.numel() * a.element_size() for a in host_args)
    idx = np.arange(Bt)

    def step_e2e():
        dev_args = [a.to(dev, non_blocking=True) for a in host_args]
        loss, st, images = network.gradient_step_image(state, 'hr', kind, *dev_args, *rta.values(), 1.0, impl=impl)
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

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        ev = torch.tensor([float(eval_per_step)], device=dev, dtype=torch.float64)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
        eval_total = ev.item()
    else:
        eval_total = float(eval_per_step)
    value = world * dense_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * dense_per_step * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        tensor_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md)'
        # per-kernel rooflines; the reported one is the kernel with the most device time
        names = ['render_fwd', 'render_bwd', 'wgrad', 'heads', 'misc']
        ms_by = {n: cat_ms[i] for i, n in enumerate(names)}
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tc_on = impl == engine.IMPL_TC
        # algorithmic work per evaluated sample: FLOPs of the MLP stage; bytes that must cross HBM
        # tcgen05 family: forward writes h0..h3 (4 x 256 B bf16) + features 64 + ReLU masks 64 + e 4 and reads the 28 B
        # packed sample (L2-resident); the fused backward (dgrad + wgrad in one launch, wgrad category empty) reads
        # those 1152 B + d loss/d o 4 and writes nothing per sample (cotangents stay in the L2-resident ring).  Both
        # are contractions: the roofline is the tensor pipe; the HBM fraction is reported next to it.
        spec = {'render_fwd': dict(bound='tensor', flop=FLOP_FWD, bytes=1184 if tc_on else 2136),
                'render_bwd': dict(bound='tensor', flop=FLOP_BWD if tc_on else 49280 * 2, bytes=1160 if tc_on else 4184),
                'wgrad': dict(bound='tensor', flop=54656 * 2, bytes=4180)}
        try:
            ncu_traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        except Exception:
            ncu_traffic = {}
        kernels = {}
        for n, sp in spec.items():
            launches = max(int(cat_sc[names.index(n)]), 1)
            per_launch_s = ms_by[n] * 1e-3 / launches
            eval_per_launch = eval_per_step * args.steps / launches
            if per_launch_s <= 0:
                continue
            tf = eval_per_launch * sp['flop'] / per_launch_s / 1e12
            gbs = eval_per_launch * sp['bytes'] / per_launch_s / 1e9
            tr = ncu_traffic.get(n + '_' + impl_name, {}).get('dram_bytes_per_eval_sample')
            k = {'bound': sp['bound'], 'avg_launch_ms': per_launch_s * 1e3, 'algorithmic_tflops': tf, 'algorithmic_gbs': gbs,
                 'ms_per_step': ms_by[n] / args.steps, 'traffic': tr * eval_per_launch if tr else None}
            k['hbm_frac'] = gbs / hbm_peak
            if sp['bound'] == 'tensor':
                k.update(achieved=tf, peak=tensor_peak, unit='TFLOP/s', frac=tf / tensor_peak)
            else:
                k.update(achieved=gbs, peak=hbm_peak, unit='GB/s', frac=gbs / hbm_peak)
            kernels[n + '_' + impl_name] = k
        kernels = {n: k for n, k in kernels.items() if k['ms_per_step'] > 0}
        dom = max(kernels, key=lambda n: kernels[n]['ms_per_step'])
        kd = kernels[dom]
        roofline = {'bound': kd['bound'], 'kernel': dom, 'achieved': kd['achieved'], 'peak': kd['peak'], 'unit': kd['unit'],
                    'frac': kd['frac'], 'traffic': kd['traffic'],
                    'peak_source': peak_src + ('; hbm_gbs' if kd['bound'] == 'hbm' else ''),
                    'avg_launch_ms': kd['avg_launch_ms'], 'kernels': kernels,
                    'kernel_ms_per_step': {n: ms_by[n] / args.steps for n in names},
                    'step_algorithmic_tflops': eval_per_step * (FLOP_FWD + FLOP_BWD) * args.steps / (ms * 1e-3) / 1e12,
                    'step_frac_of_tensor_peak': eval_per_step * (FLOP_FWD + FLOP_BWD) * args.steps / (ms * 1e-3) / 1e12 / tensor_peak}
        cpu = None
        if not args.no_cpu_baseline:
            nfr = args.cpu_frames or 8
            v, dt, cores = cpu_reference_leg(args.workload, nfr)
            cpu = {'value': v, 'unit': 'dense samples/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d of %d frames (%.1f s), float32 dense torch-CPU restatement of the reference JAX path'
                             % (nfr, Bt, dt)}
        line = {
            'metric': 'geodesic samples/s, fwd+bwd train step', 'value': value, 'unit': 'dense samples/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16x3 (fwd) / bf16 (bwd) split operands, f32 accumulate (tcgen05)' if impl == engine.IMPL_TC else 'f32 (FFMA)',
            'data': 'synthetic',
            'config': {'workload': args.workload, 'rays': P, 'samples_per_ray': Gs, 'frames_per_gpu': Bt,
                       'stokes': S, 'loss': kind, 'mlp': '4x128 relu + skip, posenc deg 3', 'kernels': impl_name,
                       'parallelism': 'frames x%d ranks, gradient all-reduce-mean' % world,
                       'active_fraction': scene.n_active / (P * Gs),
                       'l2_policy': 'inputs larger than L2: each step streams the per-frame activation workspace '
                                    '(>> 126 MB); the 8 MB packed geodesic set is L2-resident by design'},
            'evaluated_samples_per_s': eval_total * args.steps / (ms * 1e-3),
            'e2e': {'value': e2e_value, 'unit': 'dense samples/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'bhnerf_b200.network.gradient_step_image (train step + all-reduce + Adam)'},
            'gpu_launches': int(sum(cat_ln)),
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu, 'last_loss': last_loss,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
