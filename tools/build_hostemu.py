"""TEST-ONLY build of tests/emul/libfistr_b200_hostemu.so: the WHOLE fistr_b200 library (C-ABI of capi.cu + api_*.inc,
its caches, option handling and launch sequences, and every kernel) compiled by g++ for the host.

  * kernels run as blocks of cooperative fibers (tests/emul/kernel_emul.hpp, kernel_emul_impl.hpp),
  * the CUDA runtime subset the library uses is a host shim (tests/emul/cuda_shim.hpp; "device memory" = host memory filled
    with a poison pattern, streams execute immediately),
  * the only SOURCE transformation is the launch syntax:   K<<<cfg>>>(args);  ->  fb::EmulCfg(cfg).launch("K", [&]() { K(args); });
    applied to copies of capi.cu / api_*.inc under tests/emul/_hostemu/ (git-ignored).

tests/test_gpu_suite_on_host_api.py loads the result IN PLACE OF the product library (by patching the loader from the
test, never from product code) and runs the bodies of the GPU tests through frontistr_b200.hecmw on the CPU.  Single
rank only (no NCCL).  It proves nothing about launch configuration limits, memory residency or speed."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "frontistr_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "emul")          # kernel_emul*.hpp, cuda_shim.hpp, nccl_shim.hpp: test-only sources
OUT_DIR = os.path.join(EMUL, "_hostemu")
OUT_SO = os.path.join(EMUL, "libfistr_b200_hostemu.so")


def rewrite_launches(text):
    """K<<<cfg>>>(args)  ->  fb::EmulCfg(cfg).launch("K", [&]() { K(args); })   (K may carry template arguments; the name is
    recorded with the launch configuration, see fbemul_launch_log)"""
    out, pos, n = [], 0, 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            break
        j = text.index(">>>", i)
        cfg = text[i + 3:j]
        # kernel name: walk back over identifier characters, '::' and a balanced <...> template argument list
        k = i
        depth = 0
        while k > pos:
            c = text[k - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            k -= 1
        name = text[k:i]
        assert name and depth == 0, text[max(0, i - 80):i + 20]
        # argument list: balanced parentheses after >>>
        a = j + 3
        while text[a] in " \t\\\n":
            a += 1
        assert text[a] == "(", text[j:j + 40]
        d, b = 0, a
        while True:
            if text[b] == "(":
                d += 1
            elif text[b] == ")":
                d -= 1
                if d == 0:
                    break
            b += 1
        args = text[a + 1:b]
        out.append(text[pos:k])
        out.append('fb::EmulCfg(%s).launch("%s", [&]() { %s(%s); })' % (cfg, " ".join(name.split()), name, args))
        pos = b + 1
        n += 1
    return "".join(out), n


def build(force=False, ubsan=None):
    """ubsan (default: env FB_EMUL_UBSAN=1): a second artefact, libfistr_b200_hostemu_ubsan.so, compiled with
    -fsanitize=undefined,alignment,bounds -fno-sanitize-recover: signed overflow in index arithmetic, misaligned vector
    accesses, out-of-range shifts and array-bound violations of every kernel abort the test that triggers them."""
    global OUT_SO
    if ubsan is None:
        ubsan = os.environ.get("FB_EMUL_UBSAN") == "1"
    OUT_SO = os.path.join(EMUL, "libfistr_b200_hostemu_ubsan.so" if ubsan else "libfistr_b200_hostemu.so")
    srcs = [f for f in os.listdir(SRC) if f.endswith((".cu", ".inc", ".cuh", ".hpp", ".cpp"))]
    newest = max(os.path.getmtime(os.path.join(SRC, f)) for f in srcs + []) if srcs else 0
    newest = max([newest] + [os.path.getmtime(os.path.join(EMUL, f)) for f in os.listdir(EMUL) if f.endswith((".hpp", ".cpp"))])
    newest = max(newest, os.path.getmtime(os.path.abspath(__file__)), os.path.getmtime(os.path.join(ROOT, "include", "fistr_b200.h")))
    if not force and os.path.exists(OUT_SO) and os.path.getmtime(OUT_SO) >= newest:
        return OUT_SO
    os.makedirs(os.path.join(OUT_DIR, "inc", "cub", "device"), exist_ok=True)
    with open(os.path.join(OUT_DIR, "inc", "cuda_runtime.h"), "w") as f:
        f.write('#include "cuda_shim.hpp"\n')
    for hdr in ("device_scan.cuh", "device_radix_sort.cuh"):
        with open(os.path.join(OUT_DIR, "inc", "cub", "device", hdr), "w") as f:
            f.write('#include "cuda_shim.hpp"\n')
    total = 0
    for name in ["capi.cu"] + [f for f in srcs if f.endswith(".inc")]:
        with open(os.path.join(SRC, name)) as f:
            text, n = rewrite_launches(f.read())
        total += n
        if name == "capi.cu":
            text = text.replace('#include "context.hpp"', '#include "context.hpp"\n#include "kernel_emul_impl.hpp"   // TEST-ONLY host execution', 1)
        with open(os.path.join(OUT_DIR, name if name.endswith(".inc") else "capi_hostemu.cpp"), "w") as f:
            f.write("// GENERATED by tools/build_hostemu.py from frontistr_b200/csrc/%s — TEST ONLY, do not edit\n" % name + text)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    cdef, clink = g.metis_flags()
    san = ["-fsanitize=undefined,alignment,bounds,float-cast-overflow", "-fno-sanitize-recover=all", "-fno-sanitize=vptr"] if ubsan else []
    cmd = ["g++", "-O1"] + san + [ "-march=x86-64-v3", "-ffp-contract=fast", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DFB_EMUL_KERNELS=1", "-DFB_EMUL_WARP=1",
           "-I", os.path.join(OUT_DIR, "inc"), "-I", OUT_DIR, "-I", SRC, "-I", EMUL, "-I", os.path.join(ROOT, "include"),
           "-o", OUT_SO, os.path.join(OUT_DIR, "capi_hostemu.cpp")] + cdef + clink + ["-ldl"]
    subprocess.check_call(cmd)
    return OUT_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
