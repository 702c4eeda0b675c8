"""Child process for the GPU parity tests: runs the reference's own GPU pipeline (oracle/_ref) on one
seeded case and writes the results to an .npz.  Kept out of the test process because the reference
changes device-wide limits (cudaDeviceSetLimit) and its K > 32 kernels fault on sm_100.

    python tests/ref_gpu_child.py <case> <K> <alpha> <delta> <block_size> <out.npz> [rows-only]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from cases import named_case  # noqa: E402
from oracle.bindings import Ref  # noqa: E402


def main():
    case, K, alpha, delta, bs, out = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    rows_only = len(sys.argv) > 7
    pkg = entry.load_package()
    name, M, N, ro, ci = named_case(pkg, case)
    ref = Ref()
    if rows_only:
        perm, clusters, ms = ref.row_reordering_gpu(M, N, ro, ci, alpha, bs)
        np.savez(out, reordered_rows=perm, num_clusters=clusters, row_ms=ms)
        return
    A, B = pkg.synth.make_ab(M, N, K)
    r = ref.bsmr_sddmm_gpu(M, N, K, ro, ci, A, B, alpha, delta, bs, iters=2)
    np.savez(out, **r)


if __name__ == "__main__":
    main()
