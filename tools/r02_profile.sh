#!/bin/bash
# Round-2 evidence run on ONE B200: the full GPU suite, smoke(), the default bench line (with the CPU arm and the R-MAT block),
# the ncu launch list of the default step, and ncu --set full of the scorer kernels (ddi) and of the GAT edge kernels (rmat-s).
# Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
python bench.py --steps 20 > gpurun_out/r02_bench_ddi_n1.json 2> gpurun_out/r02_bench_ddi_n1.err; python tools/bench_summary.py ddi-n1 < gpurun_out/r02_bench_ddi_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-scaling"
$CMD > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on \
    -k "regex:score_fwd_kernel|score_bwd_dz_kernel|score_bwd_dw_kernel" -s 6 -c 3 -f -o gpurun_out/r02_scorer $CMD > gpurun_out/ncu2.log 2>&1
CMD2="python bench.py --workload rmat-s --steps 2 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on \
    -k "regex:gat_fwd_kernel|gat_bwd_rows_kernel|spmm_csc_kernel" -s 15 -c 5 -f -o gpurun_out/r02_gat_rmats $CMD2 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
