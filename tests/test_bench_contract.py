"""bench.py's output contract: the CPU (reference) arm is run here on the small workload; the GPU arm's line is checked on
the committed end-of-round measurements (profiles/r02_bench_ddi_n{1,8}.json).  CPU only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "gat_fwd_bwd_layer_edges_per_sec" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None      # BASELINE.md publishes no number
    assert isinstance(d["config"]["workload"], str) and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["value"] > 0 and d["e2e"]["value"] > 0


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                  # ONE json line
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("small: 1000 nodes")          # same wording as the GPU arm's config


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def _load_line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_committed_gpu_line_has_the_contract_keys():
    d = _load_line("r02_bench_ddi_n1.json")
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] in ("hbm", "tensor")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    # frac is the SURVEY 8d algorithmic figure; the tensor-pipe view counts the algorithmic flops once (x3 = issued)
    assert rf["tensor_frac_algorithmic"] < 1 / 3 + 1e-6 and abs(rf["tensor_frac_issued_3xtf32"] - 3 * rf["tensor_frac_algorithmic"]) < 2e-3
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"] * 1.001             # host-fed can at best match the device-resident number
    assert d["config"]["dropout"] == 0.0 and "general tensor-core backward" in d["scorer_backward"]
    assert d["contraction_free_scorer_backward"]["ms_per_step"] < d["ms_per_step"]      # the variant is reported beside, not as, the headline
    par = d["parity"]
    assert par["ok"] and par["worst_over_ranks"] <= par["tolerance"] == 1e-4 and par["rows_sampled"] == 256 and len(par["layers"]) == 2
    ss = d["strong_scaling"]
    assert ss["scaling"] == "strong" and ss["config"]["workload"].startswith("rmat: 2000000 nodes") and ss["parity"]["ok"]


def test_committed_multi_gpu_line():
    d1, d8 = _load_line("r02_bench_ddi_n1.json"), _load_line("r02_bench_ddi_n8.json")
    _check_common(d8)
    assert d8["n_gpus"] == 8 and d8["scaling"] == "weak" and d8["parity"]["ok"] and "peer memory" in d8["per_gpu"]
    assert d8["value"] > 5 * d1["value"]                      # weak scaling of the DDI shape: > 5x the single-GPU rate
    s1, s8 = d1["strong_scaling"], d8["strong_scaling"]
    assert s8["n_gpus"] == 8 and s8["config"] == s1["config"] and s8["parity"]["ok"]
    assert s1["ms_per_step"] / s8["ms_per_step"] > 4.0        # the 100 M-edge graph: strong scaling 1 -> 8 GPUs
