#!/bin/bash
# One gpurun call that settles what round 1 left pending:
#   * the whole GPU suite on the final code
#   * the default step with the pair order by label (default) and by (label, src) (MSHA_NLL_ORDER=src), with the per-step trace
#     of the device-timed loop (host enqueue ms, GPU ms, allocator reserve, GC) -- the src order becomes the default if its
#     trace shows no allocator growth inside the timed loop and dev ms/step is ~4.05
#   usage:  gpurun --timeout 300 -- 'bash tools/next_round_checks.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/next_gpu_tests.log
for order in label src; do
    MSHA_BENCH_TRACE=1 MSHA_NLL_ORDER=$order python bench.py --no-cpu-baseline --steps 20 \
        > gpurun_out/next_bench_$order.json 2> gpurun_out/next_bench_$order.trace
    tail -1 gpurun_out/next_bench_$order.json | python tools/bench_summary.py "order=$order"
    grep "dev loop trace" gpurun_out/next_bench_$order.trace | cut -c1-700
done
