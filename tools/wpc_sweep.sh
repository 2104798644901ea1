#!/bin/bash
# warps-per-CTA / edges-in-flight variants of the GAT edge kernels on the power-law workload (1 GPU)
mkdir -p gpurun_out
for v in "8 8" "2 8" "1 8" "1 4" "2 4" "8 4"; do
    set -- $v
    MSHA_GAT_WPC=$1 MSHA_GAT_ROWS_EB=$2 python bench.py --workload ${WL:-rmat-s} --no-cpu-baseline --steps 5 > gpurun_out/wpc_$1_$2.json 2> gpurun_out/wpc_$1_$2.err
    python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/wpc_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
k = {r["call"]: r for r in d["kernels"]}
print(f"wpc={sys.argv[1]} eb={sys.argv[2]} step {d['ms_per_step']:.2f} ms", {n: (k[n]["ms_per_unit"], k[n]["frac_hbm"]) for n in ("msha_gat_fwd", "msha_gat_bwd_rows", "msha_spmm_csc") if n in k},
      "layer frac", d["gat_layer_roofline"]["frac"], "parity", d["parity"]["worst_over_ranks"])
PY
done
