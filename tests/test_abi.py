"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol include/tnqs_cuda.h declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tnqs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnqs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tensornetworkquantumsimulator_b200 import lib
    L = lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/tnqs_cuda.h but not exported"
    assert sorted(lib.SYMBOLS) == syms
    assert L.tnqs_version().decode().startswith("tnqs-cuda")
    assert L.tnqs_status_string(-3).decode() == "gates of a batch overlap"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tensornetworkquantumsimulator_b200 import engine, graphs, lib
    g = graphs.named_grid((2, 2))
    with pytest.raises(lib.TnqsError) as ei:
        engine.tensornetworkstate(lambda v: "↑", g)
    assert ei.value.status == -6


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensornetworkquantumsimulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
