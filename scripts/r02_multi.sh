#!/bin/bash
# Round-2 multi-GPU pass (run under gpurun --gpus N): bench at N, rank-mode check, whole-job runs in both launch modes.
#   bash scripts/r02_multi.sh 8          bench + whole job (rank and local mode)
#   bash scripts/r02_multi.sh 8 all      ... plus the rank-mode check and the density-fitted whole jobs
set -x
cd "$(dirname "$0")/.."
N=${1:-8}
ALL=${2:-}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 5 --warmup 2 > $O/r02_bench_n$N.log 2>&1
tail -1 $O/r02_bench_n$N.log | cut -c1-300
timeout 400 $TR --master-port 29523 scripts/full_job.py > $O/r02_fulljob_ranks$N.log 2>&1
tail -1 $O/r02_fulljob_ranks$N.log
timeout 400 python scripts/full_job.py --ngpu $N --verbose 2 > $O/r02_fulljob_n$N.log 2>&1
tail -3 $O/r02_fulljob_n$N.log
if [ -n "$ALL" ]; then
  timeout 300 $TR --master-port 29522 scripts/check_rank_mode.py > $O/r02_rankmode_n$N.log 2>&1
  tail -1 $O/r02_rankmode_n$N.log | cut -c1-300
  timeout 400 $TR --master-port 29524 scripts/full_job.py --df > $O/r02_fulljob_ranks${N}_df.log 2>&1
  tail -1 $O/r02_fulljob_ranks${N}_df.log
  timeout 400 $TR --master-port 29525 scripts/full_job.py --df --df-block 6 > $O/r02_fulljob_ranks${N}_df_b6.log 2>&1
  tail -1 $O/r02_fulljob_ranks${N}_df_b6.log
fi
