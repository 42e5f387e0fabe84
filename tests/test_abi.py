"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/smallk_b200.h declares; without a GPU the product refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import smallk_b200 as sk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "smallk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = sk.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for s in names:
        assert hasattr(lib, s), f"{s} declared in include/smallk_b200.h but not exported"
    assert set(sk.EXPORTS) <= set(names)


def test_result_codes_match_reference_enum():
    # common/include/nmf.hpp:17-26
    assert (sk.OK, sk.NOTINITIALIZED, sk.INITIALIZED, sk.BAD_PARAM, sk.FAILURE, sk.SIZE_TOO_LARGE,
            sk.FLATCLUST_FAILURE) == (0, -1, -2, -3, -4, -5, -6)
    assert sk.ALGORITHMS == {"MU": 0, "HALS": 1, "RANK2": 2, "BPP": 3}


def test_options_struct_layout():
    o = sk.make_options(10, 20, 3)
    assert ctypes.sizeof(sk.NmfOptions) == 8 + 11 * 4 + 4     # double + 11 ints, padded to 8
    assert (o.height, o.width, o.k, o.min_iter, o.max_iter) == (10, 20, 3, 5, 5000)
    assert abs(o.tol - 0.005) < 1e-18                          # nmf/src/command_line.cpp:173-194


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(sk.SmallkError):
        sk.Context(0)
