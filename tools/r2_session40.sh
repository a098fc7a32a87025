#!/bin/bash
# oriented-box kernel: 256 / 512 / 1024 threads per object
for t in 256 512 1024; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_obb$t.so
  echo "== $t threads"; python tools/time_obb.py 2>&1 | grep "n="
done
export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_obb512.so
python -m pytest tests/test_postproc_gpu.py -m gpu -q -x 2>&1 | tail -2
