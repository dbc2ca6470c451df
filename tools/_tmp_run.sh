python bench.py > gpurun_out/r02_bench_1gpu_v4.json 2> gpurun_out/r02_bench_1gpu_v4.err; tail -c 300 gpurun_out/r02_bench_1gpu_v4.err
