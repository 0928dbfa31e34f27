"""Probes behind the fused-backward design (bhnerf_b200/csrc/dsmem_probe.cu): DSMEM push bandwidth of a CTA pair
and legality/exactness of mixed bf16 x fp16 operands in one tcgen05.mma.  Usage: python scripts/run_dsmem_probe.py"""
import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, 'bhnerf_b200', 'lib', 'libbhnerf_dsmem_probe.so'))
lib.dsmem_bw_run.restype = C.c_int
lib.dsmem_bw_run.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
lib.mixed_fmt_run.restype = C.c_int
lib.mixed_fmt_run.argtypes = [C.c_void_p] * 3 + [C.c_int] * 2 + [C.c_void_p, C.c_void_p]


def bw(mode, nthreads, reps, chunk=0):
    out = torch.zeros(4, dtype=torch.int64, device='cuda')
    for _ in range(2):
        rc = lib.dsmem_bw_run(mode, nthreads, reps, chunk, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    cyc, nbytes, flag, _ = out.tolist()
    print('mode %d (%s) threads %4d chunk %5d: %8d cycles for %8d B -> %6.1f B/clk  rc=%d flag=%d' % (
        mode, 'st.shared::cluster.v4' if mode == 0 else 'cp.async.bulk smem->dsmem', nthreads, chunk, cyc, nbytes,
        nbytes / max(cyc, 1), rc, flag), flush=True)


def mixed(a_fmt, b_fmt):
    g = torch.Generator(device='cpu').manual_seed(1)
    A = torch.randn(128, 64, generator=g); B = torch.randn(64, 64, generator=g)
    D = torch.full((128, 64), float('nan'), device='cuda'); st = torch.zeros(1, dtype=torch.int32, device='cuda')
    rc = lib.mixed_fmt_run(A.cuda().data_ptr(), B.cuda().data_ptr(), D.data_ptr(), a_fmt, b_fmt, st.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print('a_fmt %d b_fmt %d: launch failed: %s' % (a_fmt, b_fmt, e)); return
    ra = (A.to(torch.bfloat16) if a_fmt else A.to(torch.float16)).double()
    rb = (B.to(torch.bfloat16) if b_fmt else B.to(torch.float16)).double()
    ref = ra @ rb
    err = ((D.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    print('A %s x B %s: rc=%d status=%d max rel err vs exact product of the rounded operands %.2e' % (
        'bf16' if a_fmt else 'fp16', 'bf16' if b_fmt else 'fp16', rc, int(st.item()), err), flush=True)


def egress2():
    """are the DSMEM path and the L2 path out of an SM independent?"""
    lib.egress2_run.restype = C.c_int
    lib.egress2_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    gbuf = torch.zeros(74 * 262144, dtype=torch.uint8, device='cuda')
    o2 = torch.zeros(4 * 74, dtype=torch.int64, device='cuda')
    reps = 24
    print('\nSM egress, 74 CTA pairs, %d x 32 KB per path:' % reps)
    for flags, name in ((1, 'DSMEM alone'), (2, 'global (L2) alone'), (3, 'both at once')):
        for _ in range(2):
            o2.zero_()
            rc = lib.egress2_run(flags, reps, 74, gbuf.data_ptr(), o2.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        assert rc == 0
        r = o2.view(74, 4).double()
        nb = reps * 32768
        line = '  %-20s' % name
        if flags & 1: line += '  DSMEM %6.1f B/clk' % (nb / r[:, 0].mean().item())
        if flags & 2: line += '  global %6.1f B/clk' % (nb / r[:, 1].mean().item())
        print(line, flush=True)

if __name__ == '__main__':
    for nt in (128, 256, 512, 1024):
        bw(0, nt, 64)
    for chunk in (32768, 8192, 2048, 512):
        bw(1, 128, 16, chunk)
    egress2()
    for a, b in ((1, 1), (0, 0), (1, 0), (0, 1)):      # last: an illegal combination would poison the context
        mixed(a, b)
