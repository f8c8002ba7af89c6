"""TEST ONLY (run under torchrun by tests/test_dist_host_api.py): the multi-rank flow of bench.py on the CPU — every
rank loads the host-compiled library (tools/build_hostemu.py; NCCL / CUDA IPC = shared-memory stand-ins), torch.distributed
runs on gloo, torch.cuda and nvidia-smi are stubbed.  Checks the flow, the transport selection and the JSON line of rank 0;
every number in that line is meaningless here."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hex", dest="n", type=int, default=8)
    ap.add_argument("--comm", default="auto")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench
    import build_hostemu
    from frontistr_b200 import _lib, hecmw
    _lib._lib = _lib.declare(ctypes.CDLL(build_hostemu.build()))
    torch.cuda.set_device = lambda *_: None
    torch.cuda.synchronize = lambda *_: None
    real_init, real_tensor = dist.init_process_group, torch.tensor
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo")
    torch.tensor = lambda *x, **kw: real_tensor(*x, **{k: v for k, v in kw.items() if k != "device"})

    class NoSampler:
        def __init__(self, *_):
            pass

        def start(self):
            return self

        def stop(self):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["stub"]}
    bench.ClockSampler = NoSampler
    hecmw.fp64_peak = lambda: 1.0
    world = int(os.environ["WORLD_SIZE"])
    args = argparse.Namespace(gpus=world, steps=1, warmup=1, impl="b200", n=a.n, tol=1.0e-8, precond=3, maxit=20000, ref_iters=0,
                              no_cpu_baseline=True, no_e2e=False, comm=a.comm)
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_b200(args)
    if int(os.environ["RANK"]) == 0:
        d = json.loads([ln for ln in buf.getvalue().splitlines() if ln.startswith("{")][-1])
        want = a.comm if a.comm in ("p2p", "nccl-zhalo") else "nccl"   # auto: the transport whose kernels have run on hardware
        gave_up = os.environ.get("FB_EMUL_P2P_TIMEOUT") is not None and want == "p2p"
        if gave_up:      # injected time-outs: the run must have fallen back to NCCL collectively and still produced its line
            want = "nccl"
        ok = (d["n_gpus"] == world and d["config"]["comm"] == want and d["value"] > 0 and d["e2e"]["value"] and d["resid_true"] < 1e-8
              and d["scaling"] == "strong" and (("comm_note" in d["config"]) == gave_up))
        print("%s host_bench_check world=%d comm=%s->%s iterations=%d e2e_h2d=%d" % ("PASS" if ok else "FAIL", world, a.comm,
              d["config"]["comm"], d["iterations_per_step"], d["e2e"]["h2d_bytes_per_step"]), flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
