#!/bin/bash
# GPU box only: chunk / pipeline policy of lrt.simulate_and_fit for a shard of $1 simulations (plain and MCMC-refined)
n=${1:-10000}
for cfg in "0 3" "$n 1" "$((n/2+1)) 2" "$((n/3+1)) 3"; do
  set -- $cfg
  python tools/lrt_probe.py $n fast $1 $2 2>/dev/null | tail -1 | cut -c1-60 | sed "s/^/plain   n=$n chunk=$1 pipe=$2: /"
  python tools/lrt_probe.py $n fast $1 $2 30 500 2>/dev/null | tail -1 | cut -c1-60 | sed "s/^/refined n=$n chunk=$1 pipe=$2: /"
done
