"""The C-ABI library loads (no GPU needed) and exports every symbol include/*.h declares."""
import ctypes as C
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(fistr_b200_\w+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_exported():
    from frontistr_b200 import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    for s in _lib.EXPORTED:
        assert s in syms


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every compute entry must fail loudly (never route through a CPU path)."""
    import torch
    from frontistr_b200 import _lib
    if torch.cuda.is_available():
        return
    L = _lib.lib()
    rc = L.fistr_b200_init(0)
    assert rc == -2
    assert b"no CPU fallback" in L.fistr_b200_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "frontistr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_plain_c_consumer_links_and_runs(tmp_path):
    """include/fistr_b200.h is valid C and a C program needs nothing but the shared library;
    the host-only entry points (hecmw_mat_con, RCB) work from C without a GPU."""
    import subprocess
    from frontistr_b200 import _lib
    _lib.lib()
    exe = str(tmp_path / "c_consumer")
    pkg = os.path.join(ROOT, "frontistr_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "c_consumer.c"), "-o", exe,
                           "-L", pkg, "-lfistr_b200", "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_consumer ok" in out.stdout


def test_fortran_interface_matches_header():
    """Every bind(C) name in integration/hecmw_b200_if.f90 is a symbol the header declares."""
    txt = open(os.path.join(ROOT, "integration", "hecmw_b200_if.f90")).read()
    names = set(re.findall(r'bind\(C,\s*name="(fistr_b200_\w+)"\)', txt))
    assert len(names) >= 18
    assert names <= set(declared_symbols())
