python -m pytest tests/test_grad_gpu.py tests/test_lbfgs_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -5
for kb in 110 74 55 224; do echo "== smem budget $kb KB"; MTG_K2_SMEM_KB=$kb python tools/k2_probe.py 32 4096 16384 65536 151552 2>&1 | grep K2; done
python tools/lrt_probe.py 10000 fast 4096 2>&1 | tail -2; python tools/lrt_probe.py 1250 fast 4096 2>&1 | tail -1
