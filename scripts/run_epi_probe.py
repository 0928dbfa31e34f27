"""Issue-rate probe of the epilogue instruction mix (bhnerf_b200/csrc/epi_probe.cu): cycles per warp-instruction per
SMSP for each op alone and for pairs (sum of the single rates = same pipe, max = different pipes).
Usage (GPU box): python scripts/run_epi_probe.py"""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

so = os.path.join(ge.LIBDIR, 'libbhnerf_epi_probe.so')
src = os.path.join(ge.CSRC, 'epi_probe.cu')
if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
    subprocess.check_call(['/usr/local/cuda/bin/nvcc'] + ge.NVCC_FLAGS + ['-o', so, src])
lib = C.CDLL(so)
lib.epi_probe_run.restype = C.c_int
lib.epi_probe_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
NAMES = {0: 'F2FP f16x2 rn', 1: 'F2FP f16x2 rz.relu', 2: 'F2FP bf16x2 rn.relu', 3: 'LOP3', 4: 'FADD', 5: 'HSET2 (set.gt.u32.f16x2)',
         6: 'HADD2.F32 (cvt.f32.f16)', 7: 'PRMT', 8: 'FFMA', 9: 'FMNMX', 10: 'IMAD', 11: 'SHF', 12: 'F2FP e4m3x2',
         20: 'F2FP + LOP3', 21: 'F2FP + FADD', 22: 'F2FP + HSET2', 23: 'F2FP + HADD2.F32', 24: 'LOP3 + FADD', 25: 'FADD + HSET2',
         26: 'FADD + HADD2.F32', 27: 'LOP3 + PRMT', 28: 'F2FP + PRMT', 29: 'F2FP + FFMA', 30: 'LOP3 + HSET2', 31: 'LOP3 + HADD2.F32',
         32: 'HSET2 + HADD2.F32', 33: 'F2FP f16 + F2FP bf16', 34: 'LOP3 + IMAD', 35: 'FADD + IMAD', 36: 'F2FP + IMAD',
         50: 'fwd epilogue mix today (3 F2FP + 4 LOP3 + 2 FADD + HSET2)', 51: 'fp16 saves (2 F2FP + 3 LOP3 + 2 FADD + HSET2)',
         52: 'fp16 saves, one residual via HADD2.F32 (2 F2FP + 2 LOP3 + 2 FADD + HSET2 + HADD2.F32)',
         53: 'dgrad epilogue mix (F2FP + LOP3 + PRMT)'}
seed = torch.randint(0, 2 ** 31 - 1, (256,), dtype=torch.int32, device='cuda')
seed = (seed & 0x007fffff) | 0x3f000000          # floats in [0.5, 1): no NaN / denormal slow paths
out = torch.zeros(148 * 512, dtype=torch.int32, device='cuda')
cyc = torch.zeros(148, dtype=torch.int64, device='cuda')
iters = 2000
st = torch.cuda.current_stream().cuda_stream
print('%-90s %10s %10s' % ('mix', 'cyc/instr', 'instr/clk'))
for i in sorted(NAMES):
    for _ in range(2):
        n = lib.epi_probe_run(i, seed.data_ptr(), out.data_ptr(), iters, cyc.data_ptr(), st)
        torch.cuda.synchronize()
    assert n > 0
    c = cyc.double().mean().item()
    per_smsp = c / (iters * n * 4)      # 4 warps per SMSP issue n instructions per iteration each
    print('%-90s %10.3f %10.3f' % (NAMES[i], per_smsp, 1.0 / per_smsp), flush=True)

# ---- stand-alone replica of the forward layer epilogue: cycles per tile-layer pair (both slots in parallel) ----
lib.epi_full_run.restype = C.c_int
lib.epi_full_run.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
gbuf = torch.empty(4 << 30, dtype=torch.uint8, device='cuda')
sink = torch.zeros(148 * 512, dtype=torch.int32, device='cuda')
FL = {63: 'all (LDTM, hi+lo split, STTM, masks, STG)', 47: 'no STG', 55: 'no masks', 39: 'no masks, no STG', 59: 'no STTM',
      62: 'no LDTM', 1: 'LDTM only', 5: 'LDTM + STTM (no conversion)', 34: 'split only (hi + lo)', 31: 'no lo plane',
      17: 'LDTM + STG', 16: 'STG only', 42: 'split + masks'}
print('\n%-60s %12s' % ('epilogue replica: parts', 'cycles per tile-layer pair'))
it2 = 400
for f, name in FL.items():
    for _ in range(2):
        rc = lib.epi_full_run(f, gbuf.data_ptr(), gbuf.numel(), it2, cyc.data_ptr(), sink.data_ptr(), st)
        torch.cuda.synchronize()
    assert rc == 0, rc
    print('%-60s %12.0f' % (name, cyc.double().mean().item() / it2), flush=True)

# ---- SM egress to an L2-resident region ----
lib.egress_run.restype = C.c_int
lib.egress_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
nbytes = torch.zeros(148, dtype=torch.int64, device='cuda')
MODES = {0: 'st.global.v4 from 16 warps', 1: 'cp.async.bulk smem->global, one 32 KB copy at a time', 2: 'cp.async.bulk, four 32 KB copies in flight',
         3: 'st.global + cp.async.bulk together', 4: 'cp.async.bulk, eight 16 KB copies from eight lanes',
         5: 'cp.async.bulk, 2 x 4 KB per group, wait_group.read 0 each', 6: 'cp.async.bulk, 2 x 4 KB per group, one group ahead'}
print('\n%-60s %8s %14s' % ('SM egress to an L2-resident region', 'CTAs', 'B/clk per SM'))
for nblk in (148, 74, 1):
    for m, name in MODES.items():
        for _ in range(2):
            rc = lib.egress_run(gbuf.data_ptr(), 200, m, nblk, cyc.data_ptr(), nbytes.data_ptr(), st)
            torch.cuda.synchronize()
        assert rc == 0, rc
        bw = nbytes[:nblk].double() / cyc[:nblk].double()
        print('%-60s %8d %14.1f   (min %.1f, max %.1f over the SMs)' % (name, nblk, bw.mean().item(), bw.min().item(), bw.max().item()), flush=True)

# ---- SM ingress ----
lib.ingress_run.restype = C.c_int
lib.ingress_run.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
print('\n%-60s %8s %14s' % ('SM ingress, cp.async.bulk global->smem, 4 x 32 KB in flight', 'CTAs', 'B/clk per SM'))
for nblk in (148, 74, 1):
    for m, name in ((0, 'L2-resident private 256 KB region'), (1, 'streaming a %d MB buffer (HBM)' % (gbuf.numel() >> 20))):
        for _ in range(2):
            rc = lib.ingress_run(gbuf.data_ptr(), gbuf.numel(), 800, m, nblk, cyc.data_ptr(), nbytes.data_ptr(), st)
            torch.cuda.synchronize()
        assert rc == 0, rc
        print('%-60s %8d %14.1f' % (name, nblk, (nbytes[:nblk].double() / cyc[:nblk].double()).mean().item()), flush=True)

# ---- SM egress, STREAMING to a buffer far larger than L2 (write-allocate + eviction to HBM) ----
lib.egress_stream_run.restype = C.c_int
lib.egress_stream_run.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
print('\n%-60s %8s %14s' % ('SM egress, streaming through a %d MB buffer' % (gbuf.numel() >> 20), 'CTAs', 'B/clk per SM'))
for nblk in (148, 74):
    for m, name in ((0, 'st.global.v4 from 16 warps'), (1, 'cp.async.bulk, 2 x 4 KB per group, one group ahead'),
                    (2, 'cp.async.bulk, 32 KB per group, one group ahead')):
        for _ in range(2):
            rc = lib.egress_stream_run(gbuf.data_ptr(), gbuf.numel(), 200, m, nblk, cyc.data_ptr(), nbytes.data_ptr(), st)
            torch.cuda.synchronize()
        assert rc == 0, rc
        bw = nbytes[:nblk].double() / cyc[:nblk].double()
        print('%-60s %8d %14.1f   (min %.1f, max %.1f over the SMs)' % (name, nblk, bw.mean().item(), bw.min().item(), bw.max().item()), flush=True)
