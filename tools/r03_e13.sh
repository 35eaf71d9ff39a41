#!/bin/bash
# GPU box only: E13 (K6): tests, throughput per slot count (MTG_E13_SLOTS), cuFFT pair time per candidate window length, launch list.  tag = $1
TAG=${1:-r03h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_simulator_gpu.py tests/test_simulator_stats_gpu.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log
for SL in 1 2 3 4; do
echo "--- fast grid, slots $SL"; MTG_E13_SLOTS=$SL timeout 300 python tools/e13_probe.py 256 fast 2>&1 | tail -2
done
echo "--- FFT lengths"; timeout 200 python tools/e13_fft_len_probe.py 2>&1 | tail -14
echo "--- exact grid"; timeout 300 python tools/e13_probe.py 128 exact 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${TAG}_launches_e13.csv \
    python tools/e13_probe.py 28 fast > gpurun_out/${TAG}_launches_e13.log 2>&1
python - $TAG <<'PY'
import csv, collections, sys
rows = list(csv.reader(open("gpurun_out/%s_launches_e13.csv" % sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r03h_launches_e13.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    d[r[ik][:60]][0] += 1; d[r[ik][:60]][1] += float(r[iv]) / 1e3
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-62s n %4d  total us %9.1f  %.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
