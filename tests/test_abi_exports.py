"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/maus_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

import adaptive_matrix_solver_b200 as pkg
from adaptive_matrix_solver_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "maus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(maus_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    if not os.path.isfile(pkg.library_path()):
        _abi.build_library()
    lib = pkg.load_library()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/maus_b200.h but not exported"
    assert set(_abi.EXPORTS) == set(names)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.MausError):
        pkg.MausEngine(0)


def test_product_package_does_not_import_oracle():
    pdir = os.path.join(ROOT, "adaptive-matrix-solver_b200")
    for fn in os.listdir(pdir):
        if fn.endswith(".py"):
            src = open(os.path.join(pdir, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn
