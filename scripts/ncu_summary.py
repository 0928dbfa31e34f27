"""Text summary of one `ncu --set full` report (key raw metrics + the hottest SASS instructions by stall samples).
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md   (runs `ncu -i` locally; no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def ncu(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = ncu(rep, 'raw')
    hdr, units, vals = raw[0], raw[1], raw[2]
    print('# ncu --set full summary of `%s`\n' % rep.split('/')[-1])
    print('kernel: `%s`\n' % vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    print('| metric | value | unit |\n|---|---|---|')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('| %s | %s | %s |' % (k, vals[i], units[i]))
    src = ncu(rep, 'source')
    h = src[1]; data = src[2:]
    isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    tot = sum(int(r[isamp]) for r in data)
    print('\nHottest SASS instructions by warp-stall samples (%d samples, %d instructions):\n' % (tot, len(data)))
    print('| samples | share | executed | SASS |\n|---|---|---|---|')
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:15]:
        print('| %d | %.1f%% | %s | `%s` |' % (int(r[isamp]), 100.0 * int(r[isamp]) / max(tot, 1), r[iex], r[isrc].strip()[:90]))
    ops = {}
    for r in data:
        op = r[isrc].strip().split()
        op = [t for t in op if not t.startswith('@')]
        if op:
            ops[op[0].split('.')[0]] = ops.get(op[0].split('.')[0], 0) + int(r[iex])
    print('\nExecuted warp-instructions by opcode (top 12): ' + ', '.join('%s %d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:12]))
    tc = {k: v for k, v in ops.items() if k in ('UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'SYNCS')}
    print('\nBlackwell-path opcodes executed: ' + ', '.join('%s %d' % kv for kv in sorted(tc.items())))


if __name__ == '__main__':
    main()
