#!/bin/bash
# time each A/B library variant on the four bench configs:  tools/ab_run.sh old g8 g4 ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  export ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_$v.so
  echo "=== $v"
  python tools/prof_run.py --config 5 --objects 9472 --iters 40 --launches 3 | grep "launch 2"
  python tools/prof_run.py --config 3 --iters 40 --launches 3 | grep "launch 2"
  python tools/prof_run.py --config 4 --iters 40 --launches 3 | grep "launch 2"
  python tools/prof_run.py --config 2 --launches 4 | grep "launch 3"
done
