#!/bin/bash
# GPU box only: chunk / pipeline policy of lrt.simulate_and_fit for a shard of $1 simulations (plain; refined with $2 = refined)
n=${1:-10000}
for cfg in "0 3" "$n 1" "$((n/2+1)) 2" "$((n/4+1)) 1" "$((n/4+1)) 2"; do
  set -- $cfg
  timeout 300 python tools/lrt_probe.py $n fast $1 $2 2>/dev/null | tail -1 | cut -c1-250 | sed "s/^/plain   n=$n chunk=$1 pipe=$2: /"
done
