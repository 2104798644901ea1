#!/bin/bash
# unroll / min-blocks variants of the forward and column kernels, edges-in-flight of the row pass (WPC = 1), 1 GPU
mkdir -p gpurun_out
run() {
    tag=$1; shift
    env "$@" MSHA_GAT_WPC=1 python bench.py --workload ${WL:-rmat-s} --no-cpu-baseline --steps 5 > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err
    python - "$tag" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/var_{sys.argv[1]}.json").read().strip().splitlines()[-1])
k = {r["call"]: r for r in d["kernels"]}
print(sys.argv[1], f"step {d['ms_per_step']:.2f} ms", {n[5:]: (k[n]["ms_per_unit"], k[n]["frac_hbm"]) for n in ("msha_gat_fwd", "msha_gat_bwd_rows", "msha_spmm_csc") if n in k},
      "layer frac", d["gat_layer_roofline"]["frac"], "parity", f"{d['parity']['worst_over_ranks']:.1e}")
PY
}
run base MSHA_GAT_ROWS_EB=4
run fwd1 MSHA_GAT_ROWS_EB=4 MSHA_GAT_FWD_VAR=1 MSHA_GAT_CSC_VAR=1
run fwd2 MSHA_GAT_ROWS_EB=4 MSHA_GAT_FWD_VAR=2 MSHA_GAT_CSC_VAR=2
run fwd3 MSHA_GAT_ROWS_EB=2 MSHA_GAT_FWD_VAR=3 MSHA_GAT_CSC_VAR=3
