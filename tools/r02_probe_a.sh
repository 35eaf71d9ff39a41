#!/bin/bash
# GPU box: round-2 opening probe -- L-BFGS active counts per round, K2 / K9 latency over batch sizes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_smi.txt 2>&1
( MTG_LBFGS_DEBUG=1 timeout 300 python tools/lrt_probe.py 1250 fast 0 1 ) > gpurun_out/r02a_lrt1250.log 2> gpurun_out/r02a_lrt1250.err
( MTG_LBFGS_DEBUG=1 timeout 300 python tools/lrt_probe.py 10000 fast 10000 1 ) > gpurun_out/r02a_lrt10000.log 2> gpurun_out/r02a_lrt10000.err
( timeout 300 python tools/k2_probe.py 128 1024 4096 10000 37888 ) > gpurun_out/r02a_k2.log 2>&1
( timeout 300 python tools/k9_probe.py 1000 ) > gpurun_out/r02a_k9.log 2>&1
tail -3 gpurun_out/r02a_lrt1250.log; cat gpurun_out/r02a_k2.log
