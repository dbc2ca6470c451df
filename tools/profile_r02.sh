#!/bin/bash
# ncu evidence of round 2 (one B200): `--set full` plus the FP32-pipe / SASS-op counters north_star
# asks for, one launch of every hot kernel at the bench size.  Each probe first runs WITHOUT ncu
# (exit code checked, its COUNTERS line kept), then under ncu.  tools/kernel_flops.py turns the
# reports into profiles/r02_*_full.csv and profiles/r02_kernel_flops.json.
#   gpurun --timeout 1500 -- 'bash tools/profile_r02.sh r02 "scan edge configs fk scan14 accept"'
set -u
R=${1:-r02}
PROBES=${2:-"scan edge configs fk scan14 accept"}
O=gpurun_out
mkdir -p $O
MET=sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__sass_thread_inst_executed_op_ffma2_pred_on.sum,sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fadd2_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_fmul2_pred_on.sum,sm__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__sass_thread_inst_executed_op_fp64_pred_on.sum,sm__sass_thread_inst_executed_op_integer_pred_on.sum,sm__sass_thread_inst_executed_op_memory_pred_on.sum,sm__sass_thread_inst_executed_op_control_pred_on.sum,sm__sass_thread_inst_executed_op_conversion_pred_on.sum
for p in $PROBES; do
  case $p in
    scan|scan14) K="regex:knn_scan_kernel"; C=1 ;;
    edge) K="regex:collide_edges_(adaptive|hard)"; C=2 ;;
    configs) K="regex:collide_configs_kernel"; C=1 ;;
    fk) K="regex:fk_chain_kernel"; C=1 ;;
    accept) K="regex:greedy_rounds_kernel"; C=1 ;;
  esac
  # the captured launch must be the probe's OWN: skip the launches of the set-up (sampling's accept
  # test also runs collide_configs; the builder's sample_free does not run the scan or edge kernels)
  S=0
  if [ $p = configs ]; then S=0; fi
  python tools/kernel_probe.py $p > $O/${R}_${p}_plain.log 2>&1 &&
    ncu --set full --metrics $MET --clock-control none --import-source on -k "$K" -s $S -c $C -f -o $O/${R}_${p} \
        python tools/kernel_probe.py $p > $O/${R}_${p}_ncu.log 2>&1
  echo "== $p: $(grep COUNTERS $O/${R}_${p}_plain.log | cut -c1-400)"; tail -2 $O/${R}_${p}_ncu.log
done
