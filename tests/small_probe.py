"""B200 box (not a pytest test): the latency-bound end of the table.

    python tests/small_probe.py            # runs itself once per library under bsmr-sddmm_b200/lib/_variants + the release build

Per library: nips at K = 32 / 128 / 256 through the wide kernel with either epilogue form (bsmr_plan_set_wide_epilogue),
and the CSR-order residual kernel on nips K = 32 / 128 and the 98 % mask at K = 64; hot (50 back to back) and cold
(L2 flushed, median of 9) in microseconds."""
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {"lib": os.path.basename(os.environ.get("BSMR_B200_LIB", "release"))}

    def cold(fn):
        ts = []
        for _ in range(9):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return round(1e3 * float(np.median(ts)), 2)

    M, N, ro, ci = pkg.synth.nips_like()
    for K in (32, 128, 256):
        A, B = pkg.synth.make_ab(M, N, K)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        dP = torch.zeros(len(ci), device="cuda")
        for form, tag in ((1, "list"), (2, "mask")):
            plan = pkg.Plan(ctx, M, N, ro, ci)
            try:
                plan.set_wide_epilogue(form)
            except AttributeError:
                plan.close()
                continue
            plan.reorder(0.3, 0.3)
            f = pkg.SDDMM_THREE_KERNEL
            plan.sddmm(K, dA, dB, dP, flags=f)
            out["nips_K%d_wide_%s_hot" % (K, tag)] = round(1e3 * plan.sddmm(K, dA, dB, dP, iterations=50, flags=f), 2)
            out["nips_K%d_wide_%s_cold" % (K, tag)] = cold(lambda: plan.sddmm(K, dA, dB, dP, flags=f, timed=False))
            plan.close()
        if K != 256:
            plan = pkg.Plan(ctx, M, N, ro, ci)
            f = pkg.SDDMM_NO_REORDER
            plan.sddmm(K, dA, dB, dP, flags=f)
            out["nips_K%d_csr_hot" % K] = round(1e3 * plan.sddmm(K, dA, dB, dP, iterations=50, flags=f), 2)
            out["nips_K%d_csr_cold" % K] = cold(lambda: plan.sddmm(K, dA, dB, dP, flags=f, timed=False))
            plan.close()
    M, N, ro, ci = pkg.synth.dlmc_mask(0.98)
    A, B = pkg.synth.make_ab(M, N, 64)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(len(ci), device="cuda")
    plan = pkg.Plan(ctx, M, N, ro, ci)
    f = pkg.SDDMM_NO_REORDER
    plan.sddmm(64, dA, dB, dP, flags=f)
    out["mask98_K64_csr_hot"] = round(1e3 * plan.sddmm(64, dA, dB, dP, iterations=50, flags=f), 2)
    out["mask98_K64_csr_cold"] = cold(lambda: plan.sddmm(64, dA, dB, dP, flags=f, timed=False))
    print("RESULT " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        return
    libs = [None] + sorted(l for l in glob.glob(os.path.join(ROOT, "bsmr-sddmm_b200", "lib", "_variants", "*.so")) if "r01" not in l)
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["BSMR_B200_LIB"] = lib
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env, timeout=600)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        print(res[-1][7:] if res else json.dumps({"lib": lib, "rc": p.returncode, "tail": (p.stdout + p.stderr)[-600:]}), flush=True)


if __name__ == "__main__":
    main()
