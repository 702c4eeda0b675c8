"""Generates tests/golden/*.npz from the reference's own code (oracle/_ref/libbsmr_ref.so).

CPU parts (sddmm_cpu, colReordering_cpu) run anywhere the library was built; with --gpu (on the
B200 box, via gpurun) the reference's bsa_rowReordering_gpu adds the row permutation.  Inputs are
the seeded synthetic matrices of bsmr-sddmm_b200/synth.py, so only outputs are stored.

    python tests/golden/make_golden.py [--gpu] [--out DIR]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle.bindings import Ref  # noqa: E402

CASES = {
    "blocks_300x520": lambda s: s.block_structured(300, 520, seed=7),
    "blocks_1000x2000": lambda s: s.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64),
    "wide_33x9000": lambda s: s.random_uniform(33, 9000, 3000, seed=3),
    "uniform_200x333": lambda s: s.random_uniform(200, 333, 4000, seed=2),
}


def main():
    gpu = "--gpu" in sys.argv
    out = os.path.dirname(os.path.abspath(__file__))
    if "--out" in sys.argv:
        out = sys.argv[sys.argv.index("--out") + 1]
        os.makedirs(out, exist_ok=True)
    pkg = entry.load_package()
    ref = Ref()
    K, alpha, block_size = 64, 0.3, 16
    for name, make in CASES.items():
        M, N, ro, ci = make(pkg.synth)
        A, B = pkg.synth.make_ab(M, N, K)
        d = dict(M=M, N=N, K=K, row_offsets=ro, col_indices=ci, alpha=alpha, block_size=block_size)
        d["P_ref_cpu"] = ref.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=1)
        if gpu:
            perm, clusters, _ = ref.row_reordering_gpu(M, N, ro, ci, alpha, block_size)
            d["perm_ref_gpu"] = perm
            d["clusters_ref_gpu"] = clusters
            rows = perm
        else:
            rows = np.nonzero(np.diff(ro.astype(np.int64)))[0].astype(np.uint32)
        d["rows"] = rows
        for delta in (0.1, 0.3):
            c = ref.col_reordering_cpu(M, N, ro, ci, rows, delta)
            for k in ["dense_cols", "dense_col_offsets", "sparse_cols", "sparse_col_offsets", "sparse_value_offsets"]:
                d["d%02d_%s" % (int(delta * 10), k)] = c[k]
        np.savez_compressed(os.path.join(out, name + ".npz"), **d)
        print("wrote", name, "gpu" if gpu else "cpu-only")
    if gpu:
        # BASELINE configs[0]/[1] at full size: nb = 777 column blocks -> 7 warps in the reference's clustering CTA, the
        # lossy block reduction.  Only the outputs are stored; the input is synth.nips_like().
        M, N, ro, ci = pkg.synth.nips_like()
        perm, clusters, _ = ref.row_reordering_gpu(M, N, ro, ci, alpha, block_size)
        np.savez_compressed(os.path.join(out, "nips_perm_ref_gpu.npz"), perm_ref_gpu=perm, num_clusters=clusters, alpha=alpha,
                            block_size=block_size, input_checksum=np.uint64(int(ci.astype(np.uint64).sum()) * 1000003 + int(ro.astype(np.uint64).sum())))
        print("wrote nips_perm_ref_gpu")


if __name__ == "__main__":
    main()
