"""Small fixed workload for ncu captures (B200 box): python tests/ncu_target.py <workload> <K> <delta> [identity] [nowide]

Runs the reorder once and a handful of SDDMM passes, so that `ncu -k regex:... -s N -c M` can pick launches."""
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
           "mask90": lambda: s.dlmc_mask(0.90)}[name]
    M, N, ro, ci = gen()
    A, B = s.make_ab(M, N, K)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(0, stream.cuda_stream)
    plan = pkg.Plan(ctx, M, N, ro, ci)
    plan.row_reorder(0.3, flags=pkg.ROW_IDENTITY if identity else pkg.ROW_REFERENCE_COMPAT)
    plan.col_reorder(delta)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(len(ci), device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(6):
        flush.fill_(1)
        plan.sddmm_profile(K, dA, dB, dP, flags=pkg.SDDMM_NO_WIDE if nowide else pkg.SDDMM_DEFAULT)
    torch.cuda.synchronize()
    print(plan.info())


if __name__ == "__main__":
    main()
