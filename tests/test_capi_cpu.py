"""CPU: the C-ABI library builds, loads and exports exactly what include/hga_b200.h declares; without a GPU every
compute entry point fails loudly (there is no CPU fallback). No kernels run here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib
    importlib.import_module("hybrid-genome-assembler_b200.build").build()
    import hga_b200
    return hga_b200.load_library()


def _declared():
    with open(os.path.join(ROOT, "include", "hga_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(hga_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    import hga_b200
    declared = _declared()
    assert declared == sorted(hga_b200.capi.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", hga_b200.library_path()], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for sym in declared:
        assert sym in exported, f"{sym} declared in include/hga_b200.h but not exported"
        getattr(lib, sym)
    assert b"sm_100a" in lib.hga_version()


def test_library_contains_sm100a_sass_only(lib):
    import hga_b200
    r = subprocess.run(["cuobjdump", "-lelf", hga_b200.library_path()], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", r.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(lib):
    import torch
    import hga_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked tests")
    with pytest.raises(hga_b200.HgaError) as e:
        hga_b200.Handle(np.array([1, 2, 3], dtype=np.uint64), 5)
    assert e.value.code == 1 and "no CPU path" in str(e.value)


def test_argument_errors_do_not_need_a_gpu(lib):
    import hga_b200
    with pytest.raises(hga_b200.HgaError) as e:
        hga_b200.Handle(np.array([1], dtype=np.uint64), 33)
    assert e.value.code == 2 and "Kmer size is too big" in str(e.value)      # KmerIterator.cpp:24-26


def test_null_arguments_of_the_tail_block_entry_points(lib):
    lib.hga_enrich_full.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.hga_get_tail_block.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.hga_enrich_full(None, 30, -1, 20, 40, 16, None) == 2 and b"NULL" in lib.hga_last_error()
    assert lib.hga_get_tail_block(None, None) == 2


def test_product_never_touches_the_oracle():
    """nothing under the package may import, link or execute oracle/"""
    pkg = os.path.join(ROOT, "hybrid-genome-assembler_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, f), errors="ignore") as fh:
                    text = fh.read()
                assert "hga_oracle" not in text and "oracle_lib" not in text and "libhga_oracle" not in text, f
