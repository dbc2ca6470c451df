O=gpurun_out; R=r01
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/knn_one.py 1000000 fk 1 12 > $O/plain_knn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knn_scan_kernel -c 1 -f -o $O/${R}_knn_scan python tools/knn_one.py 1000000 fk 1 12 > $O/ncu_knn.log 2>&1
python bench.py --steps 1 --warmup 1 --nodes 250000 --no-cpu > $O/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_raw.csv python bench.py --steps 1 --warmup 1 --nodes 250000 --no-cpu > $O/ncu_launches.log 2>&1
python tools/summarise_profiles.py $R > $O/summarise.log 2>&1
python bench.py > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err || echo "bench failed"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err || echo "reference arm failed"
python tools/phase_probe.py 1000000 > $O/${R}_phase_probe.txt 2>&1
python tools/configs_probe.py > $O/${R}_configs_probe.txt 2>&1
tail -3 $O/plain_knn.log
python -c "
import json
d=json.loads(open('$O/${R}_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['traffic'], d['gpu_launches']); print(d['phases_ms_per_step'])"
