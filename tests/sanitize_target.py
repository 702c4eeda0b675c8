"""Small pass over every kernel (B200 box, not a pytest test): python tests/sanitize_target.py

Meant as the target of `compute-sanitizer --tool memcheck` (closed on this pool's boxes in round 1, so it has only been
run plain): one reorder (row clustering, column reorder, wide format with the bar lowered so that small matrices have
wide groups) and one SDDMM per execution plan and K on small matrices; prints the mismatch count against the oracle
(0 expected).  Both forms of the wide epilogue are forced in turn (bsmr_plan_set_wide_epilogue)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402


def main():
    import torch
    pkg = entry.load_package()
    oracle = Oracle()
    ctx = pkg.Context(0)
    s = pkg.synth
    cases = [("blocks", *s.block_structured(600, 900, seed=3, groups=6, cols_per_group=48)),
             ("mask", *s.dlmc_mask(0.90, n=512, seed=4)),
             ("uniform", *s.random_uniform(300, 2000, 24000, seed=5))]
    bad = 0
    for name, M, N, ro, ci in cases:
        for K in (32, 128, 256, 20):
            A, B = s.make_ab(M, N, K)
            want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci)
            plan = pkg.Plan(ctx, M, N, ro, ci)
            plan.set_wide_ratio(1.0)
            dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
            for form in (pkg.WIDE_EPILOGUE_LIST, pkg.WIDE_EPILOGUE_MASK):
                plan.set_wide_epilogue(form)
                plan.reorder(0.3, 0.3, block_size=16)
                for flags in (pkg.SDDMM_THREE_KERNEL, pkg.SDDMM_NO_WIDE, pkg.SDDMM_NO_REORDER):
                    dP = torch.zeros(len(ci), device="cuda")
                    plan.sddmm(K, dA, dB, dP, flags=flags)
                    torch.cuda.synchronize()
                    bad += oracle.check_data(want, dP.cpu().numpy())
            print(name, K, plan.info()["num_wide_groups"], "wide groups, mismatches so far", bad, flush=True)
            plan.close()
    print("total mismatches", bad)


if __name__ == "__main__":
    main()
