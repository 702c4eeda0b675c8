"""B200 box (not a pytest test): A/B of the wide kernel on the DLMC masks -- identity against BSMR row order, this build
against the libraries under bsmr-sddmm_b200/lib/_variants (e.g. the previous round's build).

    python tests/mask_ab_probe.py            # runs itself once per library
"""
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
        return float(np.median(ts))

    for name, gen, K in (("mask90", lambda: pkg.synth.dlmc_mask(0.90), 64), ("mask70", lambda: pkg.synth.dlmc_mask(0.70), 64),
                         ("nips", lambda: pkg.synth.nips_like(), 128)):
        M, N, ro, ci = gen()
        A, B = pkg.synth.make_ab(M, N, K)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        dP = torch.zeros(len(ci), device="cuda")
        for order, flags in (("identity", pkg.ROW_IDENTITY), ("bsmr", pkg.ROW_REFERENCE_COMPAT)):
            plan = pkg.Plan(ctx, M, N, ro, ci)
            plan.reorder(0.3, 0.3, flags=flags)
            f = pkg.SDDMM_THREE_KERNEL
            plan.sddmm(K, dA, dB, dP, flags=f)
            out["%s_%s_hot_us" % (name, order)] = 1e3 * plan.sddmm(K, dA, dB, dP, iterations=50, flags=f)
            out["%s_%s_cold_us" % (name, order)] = 1e3 * cold(lambda: plan.sddmm(K, dA, dB, dP, flags=f, timed=False))
            out["%s_%s_csr_hot_us" % (name, order)] = 1e3 * plan.sddmm(K, dA, dB, dP, iterations=50, flags=pkg.SDDMM_NO_REORDER)
            plan.close()
    print("RESULT " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        return
    libs = [None] + sorted(glob.glob(os.path.join(ROOT, "bsmr-sddmm_b200", "lib", "_variants", "*r01*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["BSMR_B200_LIB"] = lib
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env, timeout=600)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        print(res[-1][7:] if res else json.dumps({"lib": lib, "rc": p.returncode, "tail": (p.stdout + p.stderr)[-600:]}), flush=True)


if __name__ == "__main__":
    main()
