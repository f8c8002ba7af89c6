"""Static checks of the PRODUCT library as nvcc compiled it for sm_100a (cuobjdump; no GPU needed): properties of the hot
kernels that neither the host emulation nor a parity test can see, and that were lost once already without anybody
noticing - after round 1 the SpMV and the fused CG update silently acquired a 280 / 240-byte per-thread stack frame (the
peer-memory tables captured by value), i.e. 30 local-memory stores per thread in kernels that are judged against the
HBM roofline.  The test pins: no stack frame / local memory in the kernels of the CG loop, 16-byte non-allocating loads
in the SpMV, and no peer-memory code in the instantiations that run on one GPU."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("cu++filt") is None,
                                reason="CUDA binary utilities not on PATH")

# kernels launched once or more per CG iteration (DIAG and SSOR paths, both NCCL schedules, peer memory) + the gather
HOT = ("spmv_sell2_kernel", "spmv_sell1_kernel", "cg_update_p_kernel", "cg_fused_update_kernel", "cg_update_xr_kernel",
       "cg_update_xrz_kernel", "cg_precond_dot_kernel", "dot_kernel", "ssor_sweep_kernel", "ssor_sweep_bwd_dot_kernel",
       "cg_update_p_zh_kernel", "cg_fused_update_zh_kernel", "gather_sum_kernel", "vec_add_kernel", "vec_copy_zero_tail_kernel")


@pytest.fixture(scope="module")
def product():
    import __graft_entry__ as g
    so = g.build_product()
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True, check=True).stdout
    names = re.findall(r"Function (\S+):", usage)
    nums = re.findall(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", usage)
    assert names and len(names) == len(nums)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    res = {}
    for mangled, d, (reg, stack, shared, local) in zip(names, dem, nums):
        d = re.sub(r"^void\s+", "", d)
        m = re.match(r"([\w:]+<.*?>)\(", d)      # template instantiation: keep the argument list, e.g. k<(int)256, (bool)0>
        res[mangled] = dict(name=m.group(1) if m else d.split("(")[0], reg=int(reg), stack=int(stack), shared=int(shared), local=int(local))
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    code = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            code[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            code[cur].append(line)
    return res, code


def _base(name):
    return name.split("<")[0].split("::")[-1]


def test_no_kernel_exceeds_the_register_file_at_its_launch_bound(product):
    res, _ = product
    assert len(res) > 60
    for r in res.values():
        assert r["reg"] <= 255
        assert r["shared"] <= 48 * 1024          # static shared memory (dynamic is checked per launch in test_gpu_suite_on_host_api)


def test_hot_kernels_have_no_stack_frame_and_no_local_memory(product):
    res, code = product
    seen = set()
    for mangled, r in res.items():
        b = _base(r["name"])
        if b not in HOT:
            continue
        seen.add(b)
        assert r["stack"] == 0 and r["local"] == 0, (r["name"], r)
        ops = " ".join(code[mangled])
        assert " STL" not in ops and " LDL" not in ops, r["name"]
    assert seen == set(HOT), set(HOT) - seen


def test_spmv_streams_the_matrix_with_16_byte_non_allocating_loads(product):
    res, code = product
    n = 0
    for mangled, r in res.items():
        if _base(r["name"]) != "spmv_sell2_kernel":
            continue
        n += 1
        ops = "\n".join(code[mangled])
        assert len(re.findall(r"LDG\.E\.NA\.128", ops)) >= 10, r["name"]      # 9 value pairs + 1 column pair per block column
        assert r["reg"] <= 64, r                                                 # 4 blocks of 256 threads per SM at the least
    assert n == 3


def test_single_gpu_instantiations_carry_no_peer_memory_code(product):
    """PEER = false (one GPU, NCCL): no system-scope access, no %globaltimer spin.  PEER = true has both."""
    res, code = product
    found = {}
    for mangled, r in res.items():
        b = _base(r["name"])
        if b in ("spmv_sell2_kernel", "cg_fused_update_kernel"):
            peer = r["name"].endswith("(bool)1>")
            ops = "\n".join(code[mangled])
            has_sys = ".SYS" in ops
            has_timer = "GLOBALTIMER" in ops
            found[(b, peer)] = (has_sys, has_timer)
            if peer:
                assert has_sys and has_timer, r["name"]
            else:
                assert not has_sys and not has_timer, r["name"]
    assert ("spmv_sell2_kernel", True) in found and ("spmv_sell2_kernel", False) in found
    assert ("cg_fused_update_kernel", True) in found and ("cg_fused_update_kernel", False) in found
