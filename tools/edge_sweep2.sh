O=gpurun_out/r02_edge_sweep2.log
: > $O
for v in "" e8 e2 d1 d3; do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}" >> $O
  python tools/kernel_probe.py edge 2>&1 | grep COUNTERS | cut -c1-220 >> $O
done
unset MMMP_B200_LIB
for h in 15 22 30 38 45; do echo "== hard_after $h" >> $O; MMMP_EDGE_HARD=$h python tools/kernel_probe.py edge 2>&1 | grep COUNTERS | cut -c1-220 >> $O; done
cat $O
