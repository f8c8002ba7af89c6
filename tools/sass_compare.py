#!/usr/bin/env python
"""Static comparison of the sm_100a kernels of HEAD with the kernels of the two builds that RAN ON HARDWARE in round 1
(no GPU needed: git worktree + nvcc + cuobjdump):

    7f35b48  the build behind profiles/r1_bench_n1.json, r1_launches.csv.gz and r1_spmv_sell2_full.* (bench + ncu, N = 1)
    57ec45e  the build of the last `pytest -m gpu` run on a B200 (31 passed; profiles/r1_pytest_gpu.md)

For every kernel of HEAD: registers / stack bytes / SASS instruction count in the three builds, and which hardware-run
build has the IDENTICAL opcode histogram (same instructions in the same numbers: the same code up to register
allocation and addresses).  Usage:  python tools/sass_compare.py > profiles/r2_sass_vs_hardware_builds.md"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
BUILDS = [("7f35b48", "measured: bench + ncu, N = 1"), ("57ec45e", "last `pytest -m gpu` on hardware")]


def load(so):
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
    names = re.findall(r"Function (\S+):", usage)
    nums = re.findall(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    code, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            code[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            op = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", op)
            code[cur].append(op.split()[0] if op else "")
    out = {}
    for mg, d, n in zip(names, dem, nums):
        d = re.sub(r"^void\s+", "", d)
        m = re.match(r"([\w:]+<.*?>)\(", d)
        nm = (m.group(1) if m else d.split("(")[0]).replace("fb::", "").replace("(int)", "").replace("(bool)", "")
        hist = {}
        for o in code.get(mg, []):
            hist[o] = hist.get(o, 0) + 1
        out[nm] = dict(reg=int(n[0]), stack=int(n[1]), n=len(code.get(mg, [])), hist=hist)
    return out


def key(n):
    args = re.findall(r"<(.*)>", n)
    return n.split("<")[0], ([x.strip() for x in args[0].split(",")] if args else [])


def main():
    builds = []
    with tempfile.TemporaryDirectory() as tmp:
        for rev, what in BUILDS:
            wt = os.path.join(tmp, rev)
            subprocess.run(["git", "-C", ROOT, "worktree", "add", "-f", wt, rev], check=True, capture_output=True)
            so = os.path.join(tmp, rev + ".so")
            subprocess.run(NVCC + ["-o", so, os.path.join(wt, "frontistr_b200", "csrc", "capi.cu"), "-ldl"], check=True, capture_output=True)
            builds.append(("%s (%s)" % (rev, what), load(so)))
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], capture_output=True)
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        builds.append(("HEAD", load(g.build_product())))
    head = builds[-1][1]
    print("| kernel (HEAD) | " + " | ".join(b[0] for b in builds) + " | identical opcode histogram in |")
    print("|---|---|---|---|---|")
    for n in sorted(head):
        if n.startswith("cub::"):
            continue
        b, a = key(n)
        cells, same = [], []
        for label, bd in builds:
            match = None
            for o in bd:
                ob, oa = key(o)
                # HEAD added trailing template arguments (PEER of the SpMV / fused update, ICINV of stiff_kernel): an old
                # instantiation corresponds to every HEAD instantiation that extends its argument list
                if ob == b and oa == a[:len(oa)]:
                    match = o
                    break
            if match is None:
                cells.append("-")
                continue
            r = bd[match]
            cells.append("%d / %d / %d" % (r["reg"], r["stack"], r["n"]))
            if label != "HEAD" and r["hist"] == head[n]["hist"]:
                same.append(label.split()[0])
        verdict = ", ".join(same) or ("new in round 2" if cells[0] == "-" and cells[1] == "-" else "changed")
        print("| `%s` | %s | %s |" % (n, " | ".join(cells), verdict))


if __name__ == "__main__":
    main()
