#!/bin/bash
# Evidence pass of a round on the GPU box (one B200): bench lines, launch list, ncu captures of
# the two dominant kernels, model-vs-hull disagreement.  Everything lands in gpurun_out/;
# tools/summarise_profiles.py turns it into profiles/<round>_*.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r01'
set -u
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err || echo "bench failed"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err || echo "reference arm failed"
# launch list of one short bench run (cold-cache, serialised: compare shares)
python bench.py --steps 1 --warmup 1 --nodes 250000 --no-cpu > $O/plain_launches.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_raw.csv \
      python bench.py --steps 1 --warmup 1 --nodes 250000 --no-cpu > $O/ncu_launches.log 2>&1
# full captures of the two dominant kernels at the bench size
python tools/knn_one.py 1000000 fk 1 12 > $O/plain_knn.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:knn_scan_kernel -c 1 -f -o $O/${R}_knn_scan \
      python tools/knn_one.py 1000000 fk 1 12 > $O/ncu_knn.log 2>&1
python tools/edge_probe.py 1000000 > $O/plain_edge.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:collide_edges_(adaptive|hard)" -c 2 -f -o $O/${R}_edge \
      python tools/edge_probe.py 1000000 > $O/ncu_edge.log 2>&1
python tools/disagreement.py 20000 > $O/${R}_disagreement.txt 2> $O/disagreement.err
python tools/phase_probe.py 1000000 > $O/${R}_phase_probe.txt 2>&1
python tools/configs_probe.py > $O/${R}_configs_probe.txt 2>&1
tail -2 $O/${R}_bench_1gpu.err
head -c 600 $O/${R}_bench_1gpu.json
