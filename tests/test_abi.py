"""The C-ABI library loads and exports exactly what include/meso_b200.h declares."""
import os
import re
import subprocess

import pytest

from meso_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "meso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(meso_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_table():
    assert header_symbols() == sorted(lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib.LIB_PATH), "libmeso_b200.so not built"
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (meso_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    L = lib.load()
    for name in lib.SIGNATURES:
        assert hasattr(L, name)


def test_no_torch_types_and_sm100a_only():
    # the boundary is plain C; the device code is sm_100a and nothing else
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    assert "torch" not in subprocess.check_output(["ldd", lib.LIB_PATH], text=True)


def test_product_does_not_import_oracle():
    for top in ("meso_b200", os.path.join("lammps", "USER-MESO-B200"), "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "meso_oracle" not in txt and "oracle/" not in txt, f


def test_every_abi_entry_cites_the_reference_interface_it_replaces():
    """include/meso_b200.h: each declaration (or the block comment right above its group) names a reference file"""
    src = open(os.path.join(ROOT, "include", "meso_b200.h")).read()
    cites = len(re.findall(r"(UM/[a-z_]+\.(?:cu|h)|src/[a-z_]+\.(?:cpp|h))(?::\d+)?", src))
    assert cites >= 40, cites


def test_fails_loudly_without_gpu(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    from meso_b200.engine import Meso, MesoError
    with pytest.raises(MesoError, match="no CPU fallback"):
        Meso(0)
