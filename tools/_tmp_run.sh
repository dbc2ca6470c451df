python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "rerank or knn" 2>&1 | tail -3
for v in walk1 walk4 ru16 st3; do echo == $v; MMMP_B200_LIB=mmmp_b200/lib/variants/libmmmp_b200_$v.so python tools/knn_one.py 1000000 fk 1 12 2>&1 | tail -2 | head -1; done
echo == default; python tools/knn_one.py 1000000 fk 1 12 2>&1 | tail -2 | head -1
