"""Runs the UMMA operand-form probe (bhnerf_b200/csrc/umma_probe.cu) on the GPU and prints a table of
max relative errors vs a float64 matmul of the bf16-rounded operands.  Usage: python scripts/run_umma_probe.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, 'bhnerf_b200', 'lib', 'libbhnerf_umma_probe.so'))
lib.umma_probe_run.restype = C.c_int
lib.umma_probe_run.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def run(K, N, a_mode, b_mode, swap_a=0, swap_b=0, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    A = torch.randn(128, K, generator=g); B = torch.randn(K, N, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((128, N), float('nan'), device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    rc = lib.umma_probe_run(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, N, a_mode, b_mode, swap_a, swap_b,
                            st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = bf(A) @ bf(B)
    err = ((D.cpu().double() - ref).abs().max() / ref.abs().max()).item()
    return rc, int(st.item()), err


if __name__ == '__main__':
    names_a = {0: 'A smem K-major', 1: 'A tmem', 2: 'A smem MN-major'}
    names_b = {0: 'B smem MN-major', 1: 'B smem K-major'}
    ok_all = True
    for a_mode in (0, 1, 2):
        for b_mode in (0, 1):
            for (K, N) in ((16, 64), (32, 64), (128, 64), (160, 64), (64, 32), (64, 128), (64, 160), (128, 16)):
                res = []
                for swa, swb in ((0, 0),):
                    rc, st, err = run(K, N, a_mode, b_mode, swa, swb)
                    res.append('%s%s:%.1e%s' % ('a' if swa else '-', 'b' if swb else '-', err, '' if (rc == 0 and st == 0) else '!rc%d st%d' % (rc, st)))
                good = res[0].split(':')[1].startswith(('0.0e', '1.', '2.', '3.', '4.', '5.', '6.', '7.', '8.', '9.')) and float(res[0].split(':')[1].split('!')[0]) < 1e-5
                ok_all &= good
                print('%-16s %-16s K=%3d N=%3d  %s  %s' % (names_a[a_mode], names_b[b_mode], K, N, 'OK ' if good else 'BAD', '  '.join(res)), flush=True)
    print('ALL OK' if ok_all else 'SOME BAD')
