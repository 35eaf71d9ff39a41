#!/bin/bash
# run the K1 bench with each experimental library variant (exp_libs/*.so) -- GPU box only
for lib in "" $(ls exp_libs/*.so 2>/dev/null); do
  for ens in ${ENS:-5920}; do
    MTG_LIB=${lib:+$PWD/$lib} python bench.py --steps 10 --no-cpu-baseline --ensembles $ens 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', $ens, '%.4g evals/s  %.3f ms  frac %.3f ok=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['results_ok']))"
  done
done
