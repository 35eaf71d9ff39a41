#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_lbfgs_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15
( MTG_LBFGS_DEBUG=1 timeout 300 python tools/lrt_probe.py 1250 fast 0 1 ) > gpurun_out/r02b_lrt1250.log 2> gpurun_out/r02b_lrt1250.err
( timeout 300 python tools/lrt_probe.py 10000 fast 10000 1 ) > gpurun_out/r02b_lrt10000.log 2> gpurun_out/r02b_lrt10000.err
( MTG_LBFGS_SWITCH=0 timeout 300 python tools/lrt_probe.py 1250 fast 0 1 ) > gpurun_out/r02b_lrt1250_noswitch.log 2>&1
for sw in 1024 2048 8192; do ( MTG_LBFGS_SWITCH=$sw timeout 300 python tools/lrt_probe.py 1250 fast 0 1 ) 2>&1 | tail -1 | sed "s/^/switch $sw: /"; done
tail -2 gpurun_out/r02b_lrt1250.log; tail -1 gpurun_out/r02b_lrt10000.log; tail -1 gpurun_out/r02b_lrt1250_noswitch.log
grep persistent gpurun_out/r02b_lrt1250.err | tail -4
