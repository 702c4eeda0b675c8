"""Golden permutations of R-MAT graphs from the CPU oracle (sparse restatement, `oracle_row_reordering_indexed`), for sizes the
oracle needs minutes to hours for (2^15 rows: 8 min, 2^16 rows: 34 min, 2^17 rows: 125 min on one core) -- too slow for the test suite, so computed once here
and committed; the GPU tests compare both clustering kernels with them.

    python tests/golden/make_graph_golden.py 15        ->  tests/golden/graph15_perm_oracle.npz
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from cases import named_case  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402


def main():
    scale = int(sys.argv[1])
    alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    pkg = entry.load_package()
    oracle = Oracle()
    _, M, N, ro, ci = named_case(pkg, "graph%d" % scale)
    bs = oracle.calculate_block_size(M, N, int(178.35 * 2 ** 30))          # a B200's free memory: only the rows x rows term uses it
    t0 = time.time()
    perm, compat, true = oracle.row_reordering_indexed(M, N, ro, ci, alpha, bs)
    out = os.path.join(HERE, "graph%d_perm_oracle.npz" % scale)
    np.savez_compressed(out, perm=perm.astype(np.uint32), num_clusters=compat, num_clusters_true=true, block_size=bs, alpha=alpha,
                        M=M, N=N, nnz=len(ci))
    print("graph%d: block size %d, %d clusters, %.0f s -> %s" % (scale, bs, true, time.time() - t0, out))


if __name__ == "__main__":
    main()
