#!/bin/bash
# GPU box only: K1 headline over the library variants in mind_the_gaps_b200/csrc/_variants (MTG_LIB override); $ENS = ensembles per variant name
run() { lib=$1; ens=$2; MTG_LIB=${lib:+$PWD/$lib} python bench.py --steps 10 --no-extras --no-cpu-baseline --ensembles $ens 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', $ens, '%.4g evals/s  %.3f ms  frac %.3f ok=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['results_ok']))"; }
run "" 5920
run mind_the_gaps_b200/csrc/_variants/lib_t256b2u1.so 5920
run mind_the_gaps_b200/csrc/_variants/lib_t256b2u4.so 5920
run mind_the_gaps_b200/csrc/_variants/lib_t256b3.so 7104
run mind_the_gaps_b200/csrc/_variants/lib_t384b2.so 7104
run mind_the_gaps_b200/csrc/_variants/lib_t192b4.so 7104
