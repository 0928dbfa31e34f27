#!/bin/bash
# Final single-GPU evidence of the round (run under gpurun): GPU test log, smoke, bench lines (default cfg2, cfg3 with the
# visibility head, the reference arm), all BASELINE.json configs, the reference-shaped training loops, the visibility heads.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > $O/r2_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1
timeout 600 python bench.py > $O/r2_bench_cfg2.json 2> $O/r2_bench_cfg2.err
timeout 600 python bench.py --workload cfg3_ngeht > $O/r2_bench_cfg3.json 2> $O/r2_bench_cfg3.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
timeout 900 python scripts/bench_configs.py > $O/r2_configs.jsonl 2> $O/r2_configs.err
timeout 600 python scripts/train_loop_bench.py > $O/r2_train_loop.jsonl 2> $O/r2_train_loop.err
timeout 300 python scripts/vis_head_bench.py > $O/r2_vis_head_bench.log 2>&1
BHNERF_PRECISION=fast timeout 300 python bench.py --no-cpu-baseline > $O/r2_bench_cfg2_fast.json 2> $O/r2_bench_cfg2_fast.err
echo done
