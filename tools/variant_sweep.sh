#!/bin/bash
# Link each pre-compiled variant of score_tcgen05.o (variants/score_*.o, built locally with different -D tuning
# macros) into the library and time the scorer microbenchmark.  Timing only -- the default build is restored by
# `python -c "from msha_gnn_b200 import _lib; _lib.build(force=True)"`.
cd "$(dirname "$0")/.."
OBJS=$(ls msha_gnn_b200/build/*.o | grep -v score_tcgen05.o)
for v in variants/score_*.o; do
  echo "== $v"
  nvcc -shared -o msha_gnn_b200/libmsha_b200.so $OBJS "$v" -lcudart -lcuda || continue
  timeout 120 python tools/score_bench.py 2>&1 | tail -2
done
