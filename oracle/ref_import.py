"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference classes.

Imports the MSHA-GNN reference modules straight from ``/root/reference`` so that
``oracle/make_golden.py`` can generate golden vectors and ``tests/`` can pin the CPU
restatement (``oracle/msha_oracle.py``) against the reference itself.  Nothing here is
copied: importable files are imported; files that execute a training script / hard-coded
file IO at import time (``LLP.py``, ``Ours.py`` -> ``import train`` -> ``dataset.py:360``)
are parsed with ``ast`` and only the wanted ``ClassDef`` nodes are executed.

``/root/reference`` exists only in the build container, never on the GPU box: nothing that
runs under ``-m gpu``, ``smoke()`` or ``bench.py`` may call into this module.
"""
from __future__ import annotations

import ast
import os
import sys
import types

REF_ROOT = os.environ.get("MSHA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "GAT.py"))


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def import_module(name: str):
    """Import one of the side-effect-free reference files: GAT, Ablation, HGANE, model."""
    assert name in ("GAT", "Ablation", "HGANE", "model"), name
    _ensure_path()
    return __import__(name)


def extract_classes(filename: str, class_names, extra_globals=None):
    """Execute only the named top-level ``class`` / ``def`` statements of a reference file (verbatim source)."""
    _ensure_path()
    path = os.path.join(REF_ROOT, filename)
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        src = f.read()
    tree = ast.parse(src, filename=path)
    wanted = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in class_names]
    missing = set(class_names) - {n.name for n in wanted}
    if missing:
        raise RuntimeError(f"{filename}: classes not found: {sorted(missing)}")
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    g = {"torch": torch, "nn": nn, "F": F, "__name__": f"ref_{filename[:-3]}"}
    if extra_globals:
        g.update(extra_globals)
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), g)
    return {n: g[n] for n in class_names}


def llp_classes():
    """LinkPredictor / Teacher_LinkPredictor / MLP / GAT(input, adj) / KD_cosine from LLP.py:34-198."""
    from torch.nn.functional import cosine_similarity          # LLP.py:3
    return extract_classes(
        "LLP.py", ["KD_cosine", "MLP", "LinkPredictor", "GraphAttentionLayer", "GAT", "Teacher_LinkPredictor"],
        extra_globals={"cosine_similarity": cosine_similarity})


def sgae_classes(Scount: int):
    """GraphSAGE from SGAE.py:41-56; ``Scount`` is a module-level global there (SGAE.py:46,70)."""
    return extract_classes("SGAE.py", ["GraphSAGE"], extra_globals={"Scount": Scount})


def ours_classes():
    """OursLayer (with the ``record`` block) / Ours from Ours.py:29-167.

    ``Ours.py:94`` writes ``train.Coeff12new``; a stub ``train`` namespace receives it.
    """
    stub = types.SimpleNamespace(Coeff12new=None)
    cls = extract_classes("Ours.py", ["OursLayer", "GraphAttentionLayer", "Ours"],
                          extra_globals={"train": stub})
    cls["train_stub"] = stub
    return cls
