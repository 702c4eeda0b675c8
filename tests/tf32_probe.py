"""B200 box: signed error of the wide kernel's TF32 products under three operand treatments (not a pytest test).

    python tests/tf32_probe.py            # spawns one child per mode (the switches are read once per process)

modes: converter warps (cvt.rna in shared memory), no conversion with an fp32 tensor map (tcgen05 truncates), no
conversion with a CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 tensor map (does the TMA unit round on the way in?)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(K):
    import torch
    import __graft_entry__ as entry
    pkg = entry.load_package()
    M, N, ro, ci = pkg.synth.nips_like()
    A, B = pkg.synth.make_ab(M, N, K)
    ctx = pkg.Context(0)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.reorder(0.3, 0.3, flags=pkg.ROW_IDENTITY)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(len(ci), device="cuda")
    plan.sddmm(K, dA, dB, dP)
    ms = plan.sddmm(K, dA, dB, dP, iterations=50)
    got = dP.cpu().numpy().astype(np.float64)
    rows = np.repeat(np.arange(M), np.diff(ro.astype(np.int64)))
    want = np.einsum("ij,ij->i", A.astype(np.float64)[rows], B.astype(np.float64)[ci])
    rel = (got - want) / np.maximum(np.abs(want), 1e-3)
    print("RESULT " + json.dumps({"K": K, "wide_tiles": plan.info()["num_wide_tiles"], "hot_ms": ms, "mean_rel": float(rel.mean()),
                                  "max_abs_rel": float(np.abs(rel).max()), "rms_rel": float(np.sqrt((rel ** 2).mean()))}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        return
    modes = [("default (tf32 maps)", {}),
             ("fp32 maps (tensor core truncates)", {"BSMR_WIDE_FP32_MAPS": "1"})]
    for K in (128, 32):
        for name, env in modes:
            e = dict(os.environ)
            e.update(env)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(K)], env=e, capture_output=True, text=True, timeout=300)
            res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            print(name, res[-1][7:] if res else (p.stdout + p.stderr)[-400:], flush=True)


if __name__ == "__main__":
    main()
