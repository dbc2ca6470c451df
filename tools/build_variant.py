#!/usr/bin/env python
"""Build a tuning variant of the library with extra nvcc defines:

    python tools/build_variant.py NAME -DMMMP_KNN_STAGES=3 -DMMMP_KNN_QSLACK=16 ...

writes mmmp_b200/lib/variants/libmmmp_b200_NAME.so (git-ignored, travels with gpurun);
select it at run time with MMMP_B200_LIB=<path>.
Knobs: MMMP_KNN_{STAGES,TILE,RU,RU2,RU4,CWARPS,QSLACK,MINB,MINB2,MINB4,WALK,SPIN_NS,PREFETCH,HILBERT}
(csrc/knn_scan.cu; run-time: MMMP_KNN_BITS=<bits per axis of the sort grid>, MMMP_KNN_ORDER=1 heaviest-first launch),
MMMP_EDGE_{LANES,DEFER_ROUNDS} (csrc/collide.cu); tools/knn_one.py prints a checksum of the answer
so that variants can be compared for equality as well as for time.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mmmp_b200 import _build as B  # noqa: E402

name, defines = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.LIBDIR, "variants")
obj_dir = os.path.join(out_dir, name)
os.makedirs(obj_dir, exist_ok=True)


def compile_one(src):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    r = subprocess.run([B.NVCC, *B.FLAGS, *defines, "-c", os.path.join(B.CSRC, src), "-o", obj],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return obj


with ThreadPoolExecutor(max_workers=8) as ex:
    objs = list(ex.map(compile_one, B.SOURCES))
lib = os.path.join(out_dir, f"libmmmp_b200_{name}.so")
subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
for o in objs:
    os.remove(o)
os.rmdir(obj_dir)
print(lib)
