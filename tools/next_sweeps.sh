#!/bin/bash
# First measurements to take when GPU time is available again.  Builds the
# tuning variants HERE (no GPU needed), then prints the one gpurun command that times them.
#   bash tools/next_sweeps.sh
# Baselines of the tree this script was written for (one B200, tools/knn_one.py 1000000 fk 1 10):
#   default build 27.3 ms, checksum 2499110023493734664 (every variant must print the same checksum)
set -eu
cd "$(dirname "$0")/.."
build() { python tools/build_variant.py "$@" | tail -1; }
# 8-GPU wave quantisation: CTAs of 64 queries on 64-row tiles (twice the CTAs per slice)
build cw2t64 -DMMMP_KNN_CWARPS=2 -DMMMP_KNN_TILE=64 -DMMMP_KNN_MINB=9
build cw2 -DMMMP_KNN_CWARPS=2 -DMMMP_KNN_MINB=6
# lock step inside a CTA: a third stage at 5 CTAs per SM needs a smaller queue
build s3q4 -DMMMP_KNN_STAGES=3 -DMMMP_KNN_QSLACK=4 -DMMMP_KNN_MINB=4
# joint-space L2 (8-column filter rows): 3 CTAs per SM without spills (default) against 4 / 5 with
build l2mb4 -DMMMP_KNN_MINB2=4
build l2mb5 -DMMMP_KNN_MINB2=5
# edge kernels: lanes per edge / rounds before an edge is deferred (flat around 4 / 2 so far)
build e8 -DMMMP_EDGE_LANES=8
build d3 -DMMMP_EDGE_DEFER_ROUNDS=3
cat <<'EOF'
gpurun --timeout 600 -- 'for v in "" cw2t64 cw2 s3q4; do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}"; python tools/knn_one.py 1000000 fk 1 10 | tail -3
  python tools/shard_probe.py 2>&1 | tail -6             # per-slice scan times at 2/4/8 shards of the same 1 M rows
done
for v in "" l2mb4 l2mb5; do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}"; python tools/knn_one.py 1000000 l2 1 10 | tail -3
done
for v in "" e8 d3; do
  if [ -n "$v" ]; then export MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so; else unset MMMP_B200_LIB; fi
  echo "== ${v:-default}"; python tools/edge_probe.py 1000000 | tail -4
done'
EOF
