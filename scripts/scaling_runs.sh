#!/bin/bash
# One-box scaling runs of the round (launch under gpurun --gpus 8): weak cfg2 at N = 1,2,4,8 (the driver's own sequence),
# strong cfg4 (128 frames over 8 ranks; cfg4 at N=1 for its baseline is a 16-frame shard = what one rank does), ensemble cfg5,
# and the 2-rank correctness check.  Lines go to gpurun_out/r2_scale_*.json.
set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_scale_weak_n1.json 2> $O/r2_scale_weak_n1.err
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 > $O/r2_scale_weak_n$n.json 2> $O/r2_scale_weak_n$n.err
done
timeout 500 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 3 --warmup 3 --scaling strong --workload cfg4_highres > $O/r2_scale_strong_cfg4_n8.json 2> $O/r2_scale_strong_cfg4_n8.err
timeout 500 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 --scaling strong --workload cfg4_highres > $O/r2_scale_strong_cfg4_n4.json 2> $O/r2_scale_strong_cfg4_n4.err
timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --ensemble > $O/r2_scale_ensemble_cfg5_n8.json 2> $O/r2_scale_ensemble_cfg5_n8.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload cfg5_alma --no-cpu-baseline > $O/r2_scale_ensemble_cfg5_n1.json 2> $O/r2_scale_ensemble_cfg5_n1.err
timeout 300 $TR --nproc-per-node 2 --master-port 29523 scripts/multirank_check.py > $O/r2_multirank_check.log 2>&1
echo done
