"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/msha_b200.h declares; the drop-in modules keep the reference's parameter names / init order and
refuse CPU tensors (no CPU fallback).  No compute call is made without a GPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

import msha_gnn_b200 as mg
from msha_gnn_b200 import _lib
from conftest import load_golden, params_of


@pytest.fixture(scope="module")
def lib():
    mg.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    protos = _lib.parse_header()
    names = [n for n, _, _ in protos]
    assert len(names) == len(set(names)) and len(names) >= 40
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/msha_b200.h but not exported"
    assert lib.msha_abi_version() == 1
    assert lib.msha_launch_count() == 0 or lib.msha_launch_count() > 0


def test_no_extra_exports(lib):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = {n for n, _, _ in _lib.parse_header()}
    assert {e for e in exported if e.startswith("msha_")} == declared


def test_workspace_queries_are_host_only(lib):
    assert lib.msha_scan_workspace_bytes(10) >= 256
    assert lib.msha_csr_from_coo_workspace_bytes(1000, 50, 8) > 2 * 1000 * 8
    assert lib.msha_bn_workspace_bytes(64) == 296 * 2 * 64 * 8


def test_invalid_arguments_report_errors(lib):
    rc = lib.msha_gat_fwd(None, None, 10, None, None, None, 64, 8, 0.2, None, None, None, 0, None, 0.0, 0, None, None, None)
    assert rc < 0 and b"H <= 32" in lib.msha_last_error()
    rc = lib.msha_gemm_f32(None, None, None, -1, 1, 1, 1, 1, 1, 0, 0, None, 0.0, 0, 0.2, None)
    assert rc < 0


def test_cpu_tensors_raise():
    layer = mg.GraphAttentionLayer(4, 3, 0.0)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        layer(torch.rand(5, 4), torch.ones(5, 3))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        mg.Graph.from_dense(torch.ones(3, 3))


@pytest.mark.parametrize("name,cls,seed", [("ablation3", "ablation3", 41), ("ablation2", "ablation2", 41),
                                             ("ours", "Ours", 41), ("ablation1", "ablation1", 41)])
def test_state_dict_and_seeded_init_match_reference(name, cls, seed):
    g = load_golden(name)
    p = params_of(g)
    N, M = g["adj"].shape
    Fin = p["Sfeatures"].shape[1]
    d = p["attention_0.W1"].shape[1] if "attention_0.W1" in p else p["attention.W1"].shape[1]
    # the generator (oracle/make_golden.py:case_msha_models) draws the gdp values with its own Generator;
    # they are recoverable from the last feature column
    gdp = {str(i): float(v) for i, v in enumerate(p["Sfeatures"][:, -1])}
    torch.manual_seed(seed)
    model = getattr(mg, cls)(in_features=Fin, out_features=d, n_classes=M, n_heads=2, dropout=0.0, gdp=gdp,
                             Scount=N, Rcount=M)
    sd = model.state_dict()
    assert set(sd.keys()) == set(p.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == p[k].shape, k
        if v.dtype.is_floating_point:
            np.testing.assert_allclose(v.numpy(), p[k], rtol=0, atol=1e-7, err_msg=k)   # same RNG consumption order


def test_gat_and_linkpredictor_state_dict():
    g = load_golden("gat")
    p = params_of(g)
    N, M = g["adj"].shape
    gdp = {str(i): float(v) for i, v in enumerate(g["gdp"])}
    torch.manual_seed(2)
    model = mg.GAT(n_features=M, n_classes=M, n_heads=2, dropout=0.0, gdp=gdp, N=N)
    sd = model.state_dict()
    assert set(sd.keys()) == set(p.keys())
    for k, v in sd.items():
        np.testing.assert_allclose(v.numpy(), p[k], rtol=0, atol=1e-7, err_msg=k)
    g = load_golden("linkpred_mlp3")
    torch.manual_seed(61)
    lp = mg.LinkPredictor("mlp", 16, 24, 1, 3, 0.0)
    for k, v in lp.state_dict().items():
        np.testing.assert_allclose(v.numpy(), params_of(g)[k], rtol=0, atol=1e-7, err_msg=k)
    g = load_golden("gcn")
    torch.manual_seed(71)
    gc = mg.GraphConvolution(10, 5)
    for k, v in gc.state_dict().items():
        np.testing.assert_allclose(v.numpy(), params_of(g)[k], rtol=0, atol=1e-7, err_msg=k)


def test_integration_doc_binding_matches_header():
    """The ctypes stub shown in INTEGRATION.md binds msha_gat_fwd with as many arguments as the header declares."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "INTEGRATION.md")) as f:
        doc = f.read()
    m = re.search(r"lib\.msha_gat_fwd\.argtypes = \[(.*?)\]\s*#", doc, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the msha_gat_fwd binding"
    shown = re.findall(r"ctypes\.c_\w+", m.group(1))
    args = {n: a for n, _, a in _lib.parse_header()}["msha_gat_fwd"]
    assert len(shown) == len(args), (len(shown), len(args))
    want = {"int": "c_int", "float": "c_float", "int64_t": "c_int64", "uint64_t": "c_uint64"}
    for doc_t, (c_t, name) in zip(shown, args):
        c_t = c_t.replace("const", "").strip()
        expect = "c_void_p" if c_t.endswith("*") else want[c_t]
        assert doc_t == "ctypes." + expect, (name, doc_t, c_t)
    call = re.search(r"rc = lib\.msha_gat_fwd\((.*?)\)\nif rc", doc, flags=re.S).group(1)
    depth, n_args = 0, 1
    for ch in call:
        depth += ch in "([" 
        depth -= ch in ")]"
        n_args += ch == "," and depth == 0
    assert n_args == len(args)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU / eager fallback: with the shared library absent every op entry point raises with build instructions."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmsha_b200.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.lib()
    from msha_gnn_b200 import ops
    monkeypatch.setattr(ops, "_fn_cache", {})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.call("msha_abi_version")
