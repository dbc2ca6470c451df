#!/bin/bash
# times the scan variants under mmmp_b200/lib/variants/ (tools/build_variant.py) on one B200:
# FK metric with dual / centroid filter rows, 7-D L2, 14-D composite rows
O=gpurun_out/r02_scan_sweep.log
: > $O
for v in "" $(ls mmmp_b200/lib/variants/ | sed -e 's/libmmmp_b200_//' -e 's/\.so//'); do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}" >> $O
  for m in fk l2; do python tools/knn_one.py 1000000 $m 1 12 2>&1 | tail -3 | tr '\n' ' ' >> $O; echo >> $O; done
  python tools/kernel_probe.py scan14 2>&1 | grep COUNTERS | cut -c1-120 >> $O
done
cat $O
