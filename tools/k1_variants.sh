#!/bin/bash
# run the K1 bench with each experimental library variant (exp_libs/*.so) -- GPU box only
for lib in "" exp_libs/mb6.so exp_libs/mb8_64.so exp_libs/t256.so; do
  for ens in 5920 7104; do
    MTG_LIB=${lib:+$PWD/$lib} python bench.py --steps 10 --no-cpu-baseline --ensembles $ens 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', $ens, '%.4g evals/s  %.3f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
  done
done
