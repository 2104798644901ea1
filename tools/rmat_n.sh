#!/bin/bash
# usage: tools/rmat_n.sh N tag [ENV=VAL ...]   -- the 100 M-edge strong-scaling workload on N GPUs, one summary line
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
    bench.py --gpus $N --workload rmat --steps 5 --no-cpu-baseline > gpurun_out/rmat_$tag.json 2> gpurun_out/rmat_$tag.err || tail -c 1500 gpurun_out/rmat_$tag.err
python - "$tag" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/rmat_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = {r["call"]: r for r in d["kernels"]}
def tot(n): return round(k[n]["avg_ms"] * k[n]["launches_per_step"], 2) if n in k else None
print(sys.argv[1], f"step {d['ms_per_step']:.2f} ms e2e {d['e2e']['ms_per_step']:.2f}", "parity", f"{d['parity']['worst_over_ranks']:.1e}",
      {n[5:]: tot(n) for n in ("msha_gat_fwd", "msha_gat_bwd_rows", "msha_spmm_csc", "msha_gemm_tf32x3", "msha_peer_signal", "msha_peer_wait", "msha_peer_sum", "msha_score_mlp_fwd", "msha_score_mlp_nll_bwd")},
      "kernel sum", round(sum(r["avg_ms"] * r["launches_per_step"] for r in d["kernels"]), 1))
PY
