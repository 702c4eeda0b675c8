"""Small seeded parity cases shared by the CPU and GPU tests."""
import numpy as np


def small_cases(pkg):
    """(name, M, N, row_offsets, col_indices) parity cases the oracle finishes in well under a second."""
    s = pkg.synth
    cases = []
    cases.append(("uniform_64x96", *s.random_uniform(64, 96, 900, seed=1)))
    cases.append(("uniform_200x333_ragged", *s.random_uniform(200, 333, 4000, seed=2)))
    cases.append(("blocks_300x520", *s.block_structured(300, 520, seed=7)))
    cases.append(("blocks_1000x2000", *s.block_structured(1000, 2000, seed=11, groups=12, cols_per_group=64)))
    cases.append(("single_row", 1, 40, np.array([0, 3], dtype=np.uint32), np.array([1, 17, 39], dtype=np.uint32)))
    cases.append(("wide_33x9000", *s.random_uniform(33, 9000, 3000, seed=3)))      # 7 reference warps: lossy tree
    cases.append(("tall_2500x64", *s.random_uniform(2500, 64, 9000, seed=4)))
    return cases


def named_case(pkg, name):
    """One case by name: the small parity cases above, or a BASELINE.json workload at full size
    (nips = configs[0]/[1]; mask70/90/98 = configs[2]; graph<scale> = R-MAT 2^scale rows with 28.6 edges per row)."""
    s = pkg.synth
    if name == "nips":
        return (name, *s.nips_like())
    if name.startswith("mask"):
        return (name, *s.dlmc_mask(int(name[4:]) / 100.0))
    if name.startswith("graph"):
        scale = int(name[5:])
        return (name, *s.rmat(scale, int(30.0e6 / (1 << 20) * (1 << scale)), scale))
    for c in small_cases(pkg):
        if c[0] == name:
            return c
    raise KeyError(name)
