"""B200 box (not a pytest test): the residual kernel's tuning knobs, one build of the library per variant.

    python tests/variant_probe.py            # runs itself once per library under bsmr-sddmm_b200/lib/_variants + the release build
    BSMR_B200_LIB=... python tests/variant_probe.py child

Variants are built with `make -C bsmr-sddmm_b200/csrc VARIANT=<name> EXTRA="-DBSMR_RES_OCC=4 ..."`.  For every library:
R-MAT 2^20 rows / 3e7 nnz at K = 128 and 256 and the nips-shaped matrix at K = 128, CSR-order residual kernel, fp32 and
fp16 B, L2 flushed before every pass, median of 7.
"""
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import torch
    import __graft_entry__ as entry
    from graph8m_probe import rmat_device
    pkg = entry.load_package()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {"lib": os.path.basename(os.environ.get("BSMR_B200_LIB", "release"))}

    def timed(fn):
        ts = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    n, ro, ci, rows = rmat_device(torch, 20, 30_000_000, seed=20)
    del rows
    plan = pkg.Plan(ctx, n, n, ro, ci, on_device=True)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    for K in (128, 256):
        dA = torch.rand((n, K), device="cuda") * 2
        dB = torch.rand((n, K), device="cuda") * 2
        dBh = dB.half()
        dP = torch.zeros(30_000_000, device="cuda")
        out["graph20_K%d_f32_ms" % K] = timed(lambda: plan.sddmm(K, dA, dB, dP, flags=pkg.SDDMM_NO_REORDER, timed=False))
        out["graph20_K%d_f16b_ms" % K] = timed(lambda: plan.sddmm_f16b(K, dA, dBh, dP, flags=pkg.SDDMM_NO_REORDER, timed=False))
        out["graph20_K%d_f16b_reordered_list_ms" % K] = timed(lambda: plan.sddmm_f16b(K, dA, dBh, dP, timed=False))
        del dA, dB, dBh, dP
    plan.close()
    M, N, ro, ci = pkg.synth.nips_like()
    plan = pkg.Plan(ctx, M, N, ro, ci)
    A, B = pkg.synth.make_ab(M, N, 128)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(len(ci), device="cuda")
    out["nips_K128_f32_ms"] = timed(lambda: plan.sddmm(128, dA, dB, dP, flags=pkg.SDDMM_NO_REORDER, timed=False))
    out["nips_K128_f16b_ms"] = timed(lambda: plan.sddmm_f16b(128, dA, dB.half(), dP, flags=pkg.SDDMM_NO_REORDER, timed=False))
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
        print(res[-1][7:] if res else json.dumps({"lib": lib, "rc": p.returncode, "tail": (p.stdout + p.stderr)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
