"""B200 box: per-CTA time stamps of the wide kernel (not a pytest test).  python tests/wide_trace.py [K] [workload] [cold]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

NAMES = ["prologue", "regs", "A stored", "B0", "B1", "B2", "B3", "prod done", "mma first", "mma last", "acc0 full", "epi0 done",
         "acc1 full", "epi1 done", "epi done", "exit"]


def main():
    import torch
    pkg = entry.load_package()
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    wl = sys.argv[2] if len(sys.argv) > 2 else "nips"
    s = pkg.synth
    M, N, ro, ci = {"nips": s.nips_like, "mask90": lambda: s.dlmc_mask(0.90)}[wl]()
    A, B = s.make_ab(M, N, K)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(len(ci), device="cuda")
    for _ in range(5):
        plan.sddmm(K, dA, dB, dP, flags=pkg.SDDMM_THREE_KERNEL)
    trace = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
    lib = pkg.lib()
    lib.bsmr_debug_set_wide_trace.argtypes = [C.c_void_p]
    lib.bsmr_debug_set_wide_trace(trace.data_ptr())
    if "cold" in sys.argv:      # L2 flushed before the traced launch, like a timed step of bench.py
        torch.empty(512 << 20, dtype=torch.uint8, device="cuda").fill_(1)
    ms = plan.sddmm(K, dA, dB, dP, flags=pkg.SDDMM_THREE_KERNEL)
    torch.cuda.synchronize()
    raw = trace.cpu().numpy().reshape(148, 32)
    t = raw[:, :16].astype(np.float64)
    used = t[:, 0] > 0
    t = t[used]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    rel[t == 0] = np.nan
    print("K=%d %s: %d CTAs, event time %.1f us" % (K, wl, used.sum(), ms * 1e3))
    for i, n in enumerate(NAMES):
        col = rel[:, i]
        print("%-10s min %6.2f  med %6.2f  max %6.2f us" % (n, np.nanmin(col), np.nanmedian(col), np.nanmax(col)))


if __name__ == "__main__":
    main()
