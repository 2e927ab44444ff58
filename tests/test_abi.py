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


def _header_prototypes():
    """name -> number of parameters, parsed from include/msfm_b200.h (comments stripped)."""
    import re
    with open(os.path.join(os.path.dirname(m.__file__), "..", "include", "msfm_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(msfm_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_ctypes_signatures_match_the_header():
    """Every function of the header is bound in _ffi.py with the same number of arguments (a drifted binding would
    corrupt the stack silently)."""
    lib = m.load_library()
    protos = _header_prototypes()
    assert set(protos) == set(m.exported_symbols())
    for name, n in protos.items():
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name} has no ctypes signature"
        assert len(fn.argtypes) == n, f"{name}: header has {n} parameters, binding {len(fn.argtypes)}"


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of the C structs as gcc sees them == the ctypes mirrors."""
    from monocularsfm_b200 import _ffi
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "msfm_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(msfm_match_options), sizeof(msfm_ba_problem),'
                   ' sizeof(msfm_ba_options), sizeof(msfm_ba_summary), offsetof(msfm_ba_problem, fx), offsetof(msfm_ba_problem, cams),'
                   ' offsetof(msfm_ba_summary, initial_cost));return 0;}\n')
    exe = tmp_path / "s"
    inc = os.path.join(os.path.dirname(m.__file__), "..", "include")
    subprocess.run(["gcc", "-std=c99", "-I", inc, "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_ffi.MatchOptions), ctypes.sizeof(_ffi.BAProblemC), ctypes.sizeof(_ffi.BAOptions), ctypes.sizeof(_ffi.BASummary),
            _ffi.BAProblemC.fx.offset, _ffi.BAProblemC.cams.offset, _ffi.BASummary.initial_cost.offset]
    assert got == want, (got, want)
