#!/bin/bash
# GPU box only: round-2 standing check -- gpu tests, smoke, default bench, reference arm (tag = $1)
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/${TAG}_smoke.log 2>&1
tail -12 gpurun_out/${TAG}_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.log; tail -5 gpurun_out/${TAG}_bench.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -c 400 gpurun_out/${TAG}_bench_ref.log
