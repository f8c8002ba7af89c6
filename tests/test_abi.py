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


def test_fortran_interface_argument_lists_match_header():
    """Arity and value/reference passing of every bind(C) interface in integration/hecmw_b200_if.f90 against the C
    prototype: a C scalar (int, double, int64_t) must be `value` in Fortran, a C pointer must NOT be (or be a
    `type(c_ptr), value`).  Stands in for the Fortran compiler this image does not have."""
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fistr_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char \*)\s*(fistr_b200_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        if args == ["void"]:
            args = []
        protos[m.group(1)] = ["*" in a or "[" in a for a in args]          # True = pointer
    f = open(os.path.join(ROOT, "integration", "hecmw_b200_if.f90")).read()
    f = re.sub(r"!.*", "", f)
    f = re.sub(r"&\s*\n\s*", " ", f)
    checked = 0
    for m in re.finditer(r"function\s+(fistr_b200_\w+)\s*\(([^)]*)\)\s*bind\(C[^\n]*\n(.*?)end function", f, flags=re.S | re.I):
        name, dummies, body = m.group(1), [a.strip().lower() for a in m.group(2).split(",") if a.strip()], m.group(3)
        assert name in protos, name
        assert len(dummies) == len(protos[name]), (name, len(dummies), len(protos[name]))
        by_value = set()
        for line in body.splitlines():
            if "::" in line and re.search(r",\s*value", line.split("::")[0], flags=re.I):
                is_cptr = "c_ptr" in line.split("::")[0].lower()
                for v in re.split(r",(?![^()]*\))", line.split("::")[1]):
                    by_value.add((re.sub(r"\(.*", "", v).strip().lower(), is_cptr))
        for dname, is_ptr in zip(dummies, protos[name]):
            val = (dname, False) in by_value
            cptr = (dname, True) in by_value
            assert (is_ptr and (cptr or not val)) or (not is_ptr and val and not cptr), (name, dname, is_ptr, val, cptr)
        checked += 1
    assert checked >= 24


def test_fortran_e20_12e3_number_format():
    """The dump files (hecmw_matrix_dump.f90: lineFormat "(e20.12e3)") carry numbers as [-]0.dddddddddddd E+eee in 20 columns:
    leading-zero mantissa of 12 digits, three-digit exponent, correctly rounded."""
    from frontistr_b200 import _lib
    L = _lib.lib()
    buf = C.create_string_buffer(24)
    cases = {1.0: " 0.100000000000E+001", -1234.5: "-0.123450000000E+004", 0.0: " 0.000000000000E+000",
             1e-300: " 0.100000000000E-299", 999.9999999999995: " 0.100000000000E+004", 210000.0: " 0.210000000000E+006",
             -0.5: "-0.500000000000E+000", 1.0 / 3.0: " 0.333333333333E+000", 2.0 / 3.0: " 0.666666666667E+000",
             6.02214076e23: " 0.602214076000E+024", -1e-5: "-0.100000000000E-004"}
    for v, want in cases.items():
        L.fistr_b200_debug_format_e20(v, buf)
        assert buf.value.decode() == want, (v, buf.value, want)
        assert len(want) == 20 and abs(float(want) - v) <= 5e-12 * abs(v)
