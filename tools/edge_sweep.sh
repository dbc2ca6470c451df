#!/bin/bash
# times the edge-kernel variants under mmmp_b200/lib/variants/ on one B200 (10 M candidate edges of one C5 arm)
O=gpurun_out/r02_edge_sweep.log
: > $O
for v in "" $(ls mmmp_b200/lib/variants/ | sed -e 's/libmmmp_b200_//' -e 's/\.so//'); do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}" >> $O
  for i in 1 2; do python tools/kernel_probe.py edge 2>&1 | grep COUNTERS | cut -c1-110 >> $O; done
  python tools/kernel_probe.py configs 2>&1 | grep COUNTERS | cut -c1-110 >> $O
done
cat $O
