#!/bin/bash
# GPU box (2 GPUs): why do the host-facing legs lose throughput under torchrun?
nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
echo "--- single process, default threads: refined LRT 5000 sims"
timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1
echo "--- single process, OMP_NUM_THREADS=1"
OMP_NUM_THREADS=1 timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1
echo "--- two concurrent processes (one per GPU), default threads"
( CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1 ) &
( CUDA_VISIBLE_DEVICES=1 timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1 ) &
wait
echo "--- two concurrent processes, OMP_NUM_THREADS=1"
( OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1 ) &
( OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=1 timeout 200 python tools/lrt_probe.py 5000 fast 0 3 30 500 2>&1 | tail -1 ) &
wait
