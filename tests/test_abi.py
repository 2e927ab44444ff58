"""CPU: the C-ABI library loads and exports every symbol include/msfm_b200.h declares; without a GPU the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

import monocularsfm_b200 as m


def test_library_exists_and_exports_header_symbols():
    lib = m.load_library()
    names = m.exported_symbols()
    assert "msfm_match_pairs" in names and "msfm_init" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/msfm_b200.h but not exported"


def test_header_is_plain_c():
    hdr = os.path.join(os.path.dirname(m.__file__), "..", "include", "msfm_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)


def test_version_string():
    lib = m.load_library()
    assert lib.msfm_version().decode().count(".") == 2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(m.MsfmError) as ei:
        m.Context(0)
    assert ei.value.code == -2      # MSFM_E_NO_DEVICE


def test_product_package_does_not_import_oracle():
    pkg = os.path.dirname(m.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                with open(os.path.join(root, f), errors="ignore") as fh:
                    src = fh.read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
