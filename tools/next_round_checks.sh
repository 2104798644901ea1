#!/bin/bash
# Measurements round 2 left open (DESIGN.md sections 7 / 8); each line is one gpurun call on an 8-GPU box:
#   * the partitioned MSHA layer with the halo exchange at the named cfg-5 shape (built, parity-tested, measured at 2 GPUs only)
#   * the R-MAT strong-scaling step with a phase trace on the final (sequential halo) code
#   usage:  gpurun --gpus 8 --timeout 900 -- 'bash tools/next_round_checks.sh'
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530"
for halo in 0 1; do
    MSHA_MSHA_HALO=$halo $T bench.py --gpus 8 --workload ours --steps 5 > gpurun_out/next_ours_n8_halo$halo.json 2> gpurun_out/next_ours_n8_halo$halo.err
    python tools/bench_summary.py "ours n8 halo=$halo" < gpurun_out/next_ours_n8_halo$halo.json
done
bash tools/rmat_n.sh 8 next_trace MSHA_P2P_TRACE=1
grep "p2p trace rank 0" gpurun_out/rmat_next_trace.err | cut -c1-2000
