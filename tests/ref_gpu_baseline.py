"""B200 box only: the reference's own GPU path (oracle/_ref, rebuilt for sm_100) and cuSPARSE SDDMM on
one workload, each in its own subprocess so that a crash in the reference cannot take the caller down.

    python tests/ref_gpu_baseline.py [nips|blocks] [K ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import sys, json, numpy as np
sys.path.insert(0, %(root)r)
import __graft_entry__ as entry
from oracle.bindings import Ref, Oracle
pkg = entry.load_package()
what, K, mode = %(what)r, %(K)d, %(mode)r
if what == "nips":
    M, N, ro, ci = pkg.synth.nips_like()
else:
    M, N, ro, ci = pkg.synth.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64)
A, B = pkg.synth.make_ab(M, N, K)
ref = Ref()
out = {"what": what, "K": K, "mode": mode, "nnz": int(len(ci))}
if mode == "cusparse":
    P, ms = ref.cusparse_sddmm(M, N, K, ro, ci, A, B, iters=20)
    out["ms"] = ms
else:
    r = ref.bsmr_sddmm_gpu(M, N, K, ro, ci, A, B, 0.3, 0.3, 16, iters=10)
    P = r["P"]
    out.update(ms=r["sddmm_ms"], row_ms=r["row_ms"], col_ms=r["col_ms"], clusters=r["num_clusters"])
want = Oracle().sddmm_cpu(M, N, K, A, B, ro, ci)
out["mismatches"] = Oracle().check_data(want, P)
out["gflops"] = 2.0 * len(ci) * K / (out["ms"] * 1e-3) / 1e9 if out["ms"] > 0 else 0.0
print("RESULT " + json.dumps(out))
"""


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "blocks"
    Ks = [int(x) for x in sys.argv[2:]] or [32, 128]
    for K in Ks:
        for mode in ("cusparse", "bsmr_ref"):
            code = CHILD % dict(root=ROOT, what=what, K=K, mode=mode)
            try:
                p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
                res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if res:
                    print(res[-1][7:], flush=True)
                else:
                    tail = (p.stdout + p.stderr).strip().splitlines()[-3:]
                    print(json.dumps({"what": what, "K": K, "mode": mode, "rc": p.returncode, "tail": tail}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"what": what, "K": K, "mode": mode, "error": "timeout"}), flush=True)


if __name__ == "__main__":
    main()
