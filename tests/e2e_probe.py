"""B200 box: steps per second of the pipelined host-data path (not a pytest test).  python tests/e2e_probe.py [steps]

Prints the copy-path switches in force (BSMR_HOST_COPY_KERNEL_MAX, BSMR_HOST_PIPE_DUPLEX) and us per nips K=128 step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    pkg = entry.load_package()
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    M, N, ro, ci = pkg.synth.nips_like()
    K = 128
    A, B = pkg.synth.make_ab(M, N, K)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
    hPs = [torch.zeros(len(ci)).pin_memory() for _ in range(2)]
    res = []
    for rep in range(4):
        for i in range(8):
            plan.sddmm_host_submit(K, hA, hB, hPs[i % 2])
        plan.sddmm_host_wait()
        t0 = time.perf_counter()
        for i in range(steps):
            plan.sddmm_host_submit(K, hA, hB, hPs[i % 2])
        plan.sddmm_host_wait()
        res.append((time.perf_counter() - t0) / steps * 1e6)
    print("copy-kernel max %s duplex %s: us per step %s" % (os.environ.get("BSMR_HOST_COPY_KERNEL_MAX", "default"),
          os.environ.get("BSMR_HOST_PIPE_DUPLEX", "no"), " ".join("%.0f" % r for r in res)), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
