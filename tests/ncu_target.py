"""Small fixed workload for ncu captures (B200 box): python tests/ncu_target.py <workload> <K> <delta> [identity] [nowide] [csr] [f16b] [batch8]

Runs the reorder once and a handful of SDDMM passes, so that `ncu -k regex:... -s N -c M` can pick launches.
csr = the CSR-order residual kernel alone, f16b = bsmr_sddmm_f16b (B stored as fp16), batch8 = one bsmr_sddmm_batch of 8 elements;
graph20d / graph23d = the R-MAT graphs generated on the GPU (tests/graph8m_probe.py: rmat_device)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    pkg = entry.load_package()
    name, K, delta = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    identity = "identity" in sys.argv[4:]
    nowide = "nowide" in sys.argv[4:]          # force the reference's split: dense-block + residual kernels only
    s = pkg.synth
    gen = {"nips": lambda: s.nips_like(), "graph17": lambda: s.rmat(17, 3_000_000, 17), "graph20": lambda: s.rmat(20, 30_000_000, 20),
           "blocks16k": lambda: s.block_structured(16000, 16000, seed=5, groups=200, cols_per_group=96, noise=0.001),
           "mask90": lambda: s.dlmc_mask(0.90), "mask70": lambda: s.dlmc_mask(0.70)}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    if name in ("graph20d", "graph23d"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from graph8m_probe import rmat_device
        scale, edges = (20, 30_000_000) if name == "graph20d" else (23, 250_000_000)
        M, ro, ci, rows = rmat_device(torch, scale, edges, seed=scale)
        N = M
        del rows
        plan = pkg.Plan(ctx, M, N, ro, ci, on_device=True)
        nnz = edges
        dA, dB = torch.rand((M, K), device="cuda") * 2, torch.rand((N, K), device="cuda") * 2
    else:
        M, N, ro, ci = gen[name]()
        A, B = s.make_ab(M, N, K)
        plan = pkg.Plan(ctx, M, N, ro, ci)
        nnz = len(ci)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    plan.row_reorder(0.3, flags=pkg.ROW_IDENTITY if identity else pkg.ROW_REFERENCE_COMPAT)
    plan.col_reorder(delta)
    dP = torch.zeros(nnz, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    if "batch8" in sys.argv[4:]:
        dA8, dB8, dP8 = dA.repeat(8, 1, 1).contiguous(), dB.repeat(8, 1, 1).contiguous(), torch.zeros((8, nnz), device="cuda")
        for _ in range(3):
            flush.fill_(1)
            plan.sddmm_batch(8, K, dA8, dB8, dP8, timed=False)
        torch.cuda.synchronize()
        print(plan.info())
        return
    dBh = dB.half() if "f16b" in sys.argv[4:] else None
    for _ in range(6):
        flush.fill_(1)
        if dBh is not None:
            plan.sddmm_f16b(K, dA, dBh, dP, flags=pkg.SDDMM_NO_REORDER if "csr" in sys.argv[4:] else pkg.SDDMM_DEFAULT, timed=False)
        elif "csr" in sys.argv[4:]:
            plan.sddmm(K, dA, dB, dP, flags=pkg.SDDMM_NO_REORDER, timed=False)
        else:
            plan.sddmm_profile(K, dA, dB, dP, flags=pkg.SDDMM_NO_WIDE if nowide else pkg.SDDMM_DEFAULT)
    torch.cuda.synchronize()
    print(plan.info())


if __name__ == "__main__":
    main()
