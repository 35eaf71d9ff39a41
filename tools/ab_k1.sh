#!/bin/bash
# GPU box only: K1 headline A/B over the experimental libraries in exp_libs/ (MTG_LIB override)
for lib in "" $(ls exp_libs/*.so 2>/dev/null); do
  MTG_LIB=${lib:+$PWD/$lib} python bench.py --steps 20 --no-extras --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', '%.4g evals/s  %.3f ms  frac %.3f  e2e %.4g ok=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['results_ok']))"
done
