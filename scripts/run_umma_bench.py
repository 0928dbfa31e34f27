"""UMMA throughput micro-benchmark (bhnerf_b200/csrc/umma_probe.cu: umma_bench_kernel): cycles per
tcgen05.mma (M=128, N, K=16) for the operand forms and SWIZZLE_NONE core-matrix arrangements the kernels use.
Ideal = N/2 cycles (B300_MICROARCH: 128xNx16 at 8192 FLOP/clk/SM)."""
import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, 'bhnerf_b200', 'lib', 'libbhnerf_umma_probe.so'))
lib.umma_bench_run.restype = C.c_int
lib.umma_bench_run.argtypes = [C.c_int] * 6 + [C.c_uint32] * 4 + [C.c_int, C.c_void_p, C.c_void_p]


def run(name, N, nks, a_mode, b_mn, ak, am, bk, bn, f16=0, grid=148, reps=200):
    out = torch.zeros(grid, dtype=torch.int64, device='cuda')
    for _ in range(2):
        rc = lib.umma_bench_run(grid, N, nks, reps, a_mode, b_mn, ak, am, bk, bn, f16, out.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    cyc = out.double().mean().item() / (reps * nks)
    print('%-58s N=%3d  %7.1f cycles/MMA (ideal %5.1f) -> %4.0f%% of peak' % (name, N, cyc, N / 2, 100 * (N / 2) / cyc), flush=True)
    return cyc


if __name__ == '__main__':
    run('UNROLLED raw: A tmem , B MN-major', 128, 8, 1, 1, 0, 0, 128, 2048, f16=8)
    run('UNROLLED raw: A smemK, B MN-major', 128, 8, 0, 1, 2048, 128, 128, 2048, f16=8)
    run('UNROLLED raw: A tmem , B MN-major N=64', 64, 8, 1, 1, 0, 0, 128, 2048, f16=8)
    run('UNROLLED raw: A tmem , B MN-major N=256', 256, 8, 1, 1, 0, 0, 128, 2048, f16=8)
    K8 = lambda rows: rows // 8 * 128
    # forward: A smem K-major [s][k] (features) or TMEM, B = W image [k][n] MN-major
    for N in (128,):
        run('fwd   A tmem , B MN-major k-groups adjacent (current)', N, 8, 1, 1, 0, 0, 128, K8(128))
        run('fwd   A tmem , B MN-major n-groups adjacent', N, 8, 1, 1, 0, 0, K8(N), 128)
        run('fwd   A smemK (k adj), B MN-major k-groups adjacent', N, 8, 0, 1, 128, K8(128), 128, K8(128))
        run('fwd   A smemK (m adj), B MN-major k-groups adjacent (current feat)', N, 8, 0, 1, K8(128), 128, 128, K8(128))
        run('fwd   A smemK (m adj), B MN-major n-groups adjacent', N, 8, 0, 1, K8(128), 128, K8(N), 128)
        run('fwd   A tmem , B MN-major k adj, fp16', N, 8, 1, 1, 0, 0, 128, K8(128), f16=1)
        # dgrad: A tmem, B = W image [k][n] read K-major: K' groups = column groups (stride K8(rows)), N' groups = row groups (128)
        run('dgrad A tmem , B K-major  K-stride 2048, N-stride 128 (current)', N, 8, 1, 0, 0, 0, K8(128), 128)
        run('dgrad A tmem , B K-major  K-stride 128, N-stride 2048', N, 8, 1, 0, 0, 0, 128, K8(128))
        # wgrad: A, B = sample images [s][c] MN-major: K groups = row groups (128), MN groups = column groups (2048)
        run('wgrad A smemMN (k adj), B MN-major (k adj) (current)', N, 8, 2, 1, 128, K8(128), 128, K8(128))
        run('wgrad A smemMN (m adj), B MN-major (n adj)', N, 8, 2, 1, K8(128), 128, K8(128), 128)
    for N in (256, 64, 32):
        run('fwd   A tmem , B MN-major k-groups adjacent', N, 8, 1, 1, 0, 0, 128, K8(128))
        run('fwd   A tmem , B MN-major n-groups adjacent', N, 8, 1, 1, 0, 0, K8(N), 128)
    run('1 CTA only: A tmem, B MN-major k adj', 128, 8, 1, 1, 0, 0, 128, K8(128), grid=1)
    run('A tmem: two alternating accumulators', 128, 8, 1, 1, 0, 0, 128, K8(128), f16=2)
    run('A tmem: no accumulate (acc=0)', 128, 8, 1, 1, 0, 0, 128, K8(128), f16=4)
    run('A smemK: two alternating accumulators', 128, 8, 0, 1, K8(128), 128, 128, K8(128), f16=2)
    run('A tmem: two alternating accumulators N=64', 64, 8, 1, 1, 0, 0, 128, K8(128), f16=2)
