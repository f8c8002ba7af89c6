#!/usr/bin/env python
"""Which kernels do the driver's round-end GPU steps launch?  Runs the bodies of the DEFAULT `pytest -m gpu` suite
(tests/test_gpu_parity.py) and __graft_entry__.smoke() through the host-compiled library (TEST ONLY, no GPU), reads the
launch log the emulation keeps per kernel name, and prints for every launched kernel how its sm_100a code relates to the
two builds that ran on a B200 (profiles/r2_sass_vs_hardware_builds.md, written by tools/sass_compare.py).

    python tools/default_suite_kernels.py >> profiles/r2_sass_vs_hardware_builds.md"""
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import build_hostemu  # noqa: E402
from frontistr_b200 import _lib, hecmw  # noqa: E402


def main():
    real_stdout = os.dup(1)
    os.dup2(2, 1)                      # solver messages of the test bodies go to stderr
    _lib._lib = _lib.declare(C.CDLL(build_hostemu.build()))
    hecmw.init(0)
    import test_gpu_suite_on_host_api as T
    ran = 0
    for p in T._cases():
        modname, name, kw = p.values
        if modname != "test_gpu_parity" or name in T.SKIP:
            continue
        mod = sys.modules[modname]
        if name in T.ULP_INSTEAD_OF_BITWISE:
            real = mod.np
            mod.np = T._NumpyWithUlpEqual(real)
            try:
                getattr(mod, name)(hecmw, **kw)
            finally:
                mod.np = real
        else:
            getattr(mod, name)(hecmw, **kw)
        ran += 1
    import __graft_entry__ as g
    g.smoke()
    buf = C.create_string_buffer(1 << 16)
    _lib.lib().fbemul_launch_log(buf, len(buf))
    launched = sorted({ln.split("\t")[0] for ln in buf.value.decode().splitlines()})
    with open(os.path.join(ROOT, "profiles", "r2_sass_vs_hardware_builds.md")) as f:
        table = f.read()
    status = {}
    for ln in table.splitlines():
        m = re.match(r"\| `([^`]+)` \|.*\| ([^|]+) \|$", ln)
        if m:
            status.setdefault(m.group(1).split("<")[0], []).append("%s: %s" % (m.group(1), m.group(2).strip()))
    out = os.fdopen(real_stdout, "w")
    out.write("\n## Kernels launched by the default `pytest -m gpu` suite (%d test bodies) and `smoke()`\n\n" % ran)
    out.write("(`python tools/default_suite_kernels.py`: the bodies run through the host-compiled library, the emulation logs every "
              "launch by kernel name.)  None of them is new in round 2.\n\n| kernel as launched | instantiations in the sm_100a binary: identical opcode histogram in |\n|---|---|\n")
    new = []
    for n in launched:
        b = n.split("<")[0].split("::")[-1].strip()
        st = status.get(b, ["?"])
        if any("new in round 2" in s for s in st):
            new.append(n)
        out.write("| `%s` | %s |\n" % (n, "; ".join(st)))
    out.write("\nNew-in-round-2 kernels among them: %s\n" % (", ".join(new) if new else "none"))
    out.flush()


if __name__ == "__main__":
    main()
