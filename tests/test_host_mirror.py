"""CPU: the host mirror of the reference's C++ API (bsmr-sddmm_b200/host/*.hpp: Matrix / CSR loaders, makeData,
Options, sddmm_cpu, checkData) driven through tests/host_mirror_harness.cpp -- plain g++, no CUDA, no C-ABI library
-- against the oracle and, where it was built, the unmodified reference (oracle/_ref)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "bsmr-sddmm_b200", "host")


@pytest.fixture(scope="session")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host_mirror") / "harness")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-Wall", "-Werror", "-I" + HOST_DIR,
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_mirror_harness.cpp"),
                           "-o", exe])

    def run(*args, threads=1):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        return subprocess.run([exe, *map(str, args)], env=env, capture_output=True, text=True, timeout=120)
    return run


def load(harness, path, tmp_path):
    out = str(tmp_path / (os.path.basename(path) + ".bin"))
    r = harness("load", path, out)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    if np.frombuffer(raw[:4], dtype=np.int32)[0] == 0:
        return None
    M, N, nnz = (int(x) for x in np.frombuffer(raw[4:16], dtype=np.uint32))
    body = np.frombuffer(raw[16:], dtype=np.uint32)
    assert len(body) == M + 1 + 2 * nnz
    return M, N, body[:M + 1], body[M + 1:M + 1 + nnz], body[M + 1 + nnz:].view(np.float32)


def same_csr(a, b):
    return a[0] == b[0] and a[1] == b[1] and all(np.array_equal(x, y) for x, y in zip(a[2:], b[2:]))


def test_mtx_loader_matches_oracle(pkg, oracle, harness, tmp_path):
    M, N, ro, ci = pkg.synth.random_uniform(50, 70, 600, seed=8)
    vals = np.random.default_rng(1).random(len(ci)).astype(np.float32)
    p = str(tmp_path / "m.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci, values=vals, shuffle_seed=3)      # shuffled lines: row-stable sort only
    got, want = load(harness, p, tmp_path), oracle.load_mtx(p)
    assert got is not None and want is not None and same_csr(got, want)
    assert np.array_equal(np.sort(got[4]), np.sort(vals))


def test_mtx_loader_matches_reference(pkg, ref, harness, tmp_path):
    M, N, ro, ci = pkg.synth.block_structured(120, 200, seed=3)
    p = str(tmp_path / "b.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci, shuffle_seed=9)
    assert same_csr(load(harness, p, tmp_path), ref.load_matrix_file(p))


def test_mtx_loader_rejections(oracle, harness, tmp_path):
    cases = {
        "dup.mtx": "%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n1 1 2\n",
        "range.mtx": "%%MatrixMarket\n3 3 2\n1 1 1\n4 2 1\n",
        "few.mtx": "%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n",
        "many.mtx": "%%MatrixMarket\n3 3 2\n1 1 1\n2 2 1\n3 3 1\n",
        "one.mtx": "%%MatrixMarket\n3 3 1\n1 1 1\n",
        "ok_novalue.mtx": "%%MatrixMarket\n% comment\n3 4 3\n3 1\n1 4\n\n2 2\n",
    }
    for name, text in cases.items():
        p = str(tmp_path / name)
        open(p, "w").write(text)
        got, want = load(harness, p, tmp_path), oracle.load_mtx(p)
        assert (got is None) == (want is None), name
        if got is not None:
            assert same_csr(got, want), name
    assert load(harness, str(tmp_path / "missing.mtx"), tmp_path) is None
    open(str(tmp_path / "m.bin2"), "w").write("1 1 1\n")
    assert load(harness, str(tmp_path / "m.bin2"), tmp_path) is None        # unsupported suffix


def test_mtx_loader_token_forms(oracle, harness, tmp_path, request):
    """Separators, line endings and number spellings the in-place parser must read exactly like std::stoi / std::stod:
    tabs, CRLF, exponents, signs, an explicit '+' on an index (std::stoi path), a value out of double range (-> 0)."""
    text = ("%%MatrixMarket matrix coordinate real general\r\n% c\r\n4 5 7\r\n"
            "1\t2\t1.5e0\r\n2 3 -2.25\r\n+3 4 7\r\n4  1   1e999\r\n1 5 .5\r\n2 1 3.\r\n4 5\r\n")
    p = str(tmp_path / "forms.mtx")
    open(p, "w", newline="").write(text)
    got, want = load(harness, p, tmp_path), oracle.load_mtx(p)
    assert got is not None and want is not None and same_csr(got, want)
    assert list(got[2]) == [0, 2, 4, 5, 7] and list(got[3]) == [1, 4, 2, 0, 3, 0, 4]
    assert list(got[4]) == [1.5, 0.5, -2.25, 3.0, 7.0, 0.0, 0.0]
    from oracle.bindings import Ref, REF_SO
    if os.path.exists(REF_SO):
        assert same_csr(got, Ref().load_matrix_file(p))


def test_smtx_and_edge_list_loaders(ref, harness, tmp_path):
    """The other two loaders of initializeFromMatrixFile (src/Matrix.cpp:296-371, 482-585) against the reference."""
    smtx = str(tmp_path / "mask.smtx")
    open(smtx, "w").write("4, 6, 7\n0 2 2 5 7\n1 4 0 3 5 2 3\n")
    edges = str(tmp_path / "graph.txt")
    open(edges, "w").write("# Directed graph\n# Nodes: 5 Edges: 6\n# FromNodeId\tToNodeId\n"
                           "10\t20\n10\t30\n20\t30\n40\t10\n30\t50\n50\t40\n")
    for p in (smtx, edges):
        got, want = load(harness, p, tmp_path), ref.load_matrix_file(p)
        assert got is not None and want is not None, p
        assert same_csr(got, want), p
    dup = str(tmp_path / "dup.smtx")
    open(dup, "w").write("2, 4, 3\n0 2 3\n1 1 2\n")
    assert load(harness, dup, tmp_path) is None and ref.load_matrix_file(dup) is None


def test_smtx_and_edge_list_loaders_known_answers(harness, tmp_path):
    """The same two files with the answers written out (runs where oracle/_ref is absent)."""
    smtx = str(tmp_path / "mask.smtx")
    open(smtx, "w").write("4, 6, 7\n0 2 2 5 7\n1 4 0 3 5 2 3\n")
    M, N, ro, ci, v = load(harness, smtx, tmp_path)
    assert (M, N) == (4, 6) and list(ro) == [0, 2, 2, 5, 7] and list(ci) == [1, 4, 0, 3, 5, 2, 3] and np.all(v == 1)
    edges = str(tmp_path / "graph.txt")
    open(edges, "w").write("# Directed graph\n# Nodes: 5 Edges: 6\n# FromNodeId\tToNodeId\n"
                           "10\t20\n10\t30\n20\t30\n40\t10\n30\t50\n50\t40\n")
    M, N, ro, ci, v = load(harness, edges, tmp_path)     # ids renumbered in order of first appearance: 10 20 30 40 50
    assert (M, N) == (5, 5) and list(ro) == [0, 2, 3, 4, 5, 6] and list(ci) == [1, 2, 2, 4, 0, 3]


def test_make_data_is_the_mt19937_stream(oracle, harness, tmp_path):
    out = str(tmp_path / "a.bin")
    assert harness("makedata", 50, 100, out).returncode == 0
    a = np.fromfile(out, dtype=np.float32)
    assert np.array_equal(a, oracle.make_data(5000))
    assert abs(float(a[0]) - 1.629447) < 1e-6            # mt19937(5489) * 2


def test_options_flags_and_positional(harness):
    def parse(*argv):
        r = harness("options", *argv)
        assert r.returncode == 0
        return dict(line.split("=", 1) for line in r.stdout.splitlines() if "=" in line), r.stderr
    o, _ = parse("-f", "dataset/nips.mtx", "-k", "128", "-a", "0.5", "-d", "0.25", "-t", "1", "-l", "logs/")
    assert o == {"inputFile": "dataset/nips.mtx", "K": "128", "alpha": "0.500000", "delta": "0.250000", "testMode": "1",
                 "logDirectory": "logs/"}
    o, _ = parse()                                       # defaults (include/Options.hpp:49-56)
    assert (o["K"], o["alpha"], o["delta"], o["testMode"]) == ("32", "0.300000", "0.300000", "0")
    o, _ = parse("m.mtx", "64")                          # positional fallback
    assert (o["inputFile"], o["K"]) == ("m.mtx", "64")
    o, err = parse("-k", "abc", "-f", "x.mtx")            # std::stoi failure is caught and reported, K keeps its default
    assert o["K"] == "32" and o["inputFile"] == "x.mtx" and "Invalid argument" in err
    o, err = parse("-F", "y.mtx", "-K")                   # upper-case aliases; a flag without a value is reported
    assert o["inputFile"] == "y.mtx" and "requires an argument" in err


def test_sddmm_cpu_and_check_data(pkg, oracle, harness, tmp_path):
    M, N, ro, ci = pkg.synth.block_structured(90, 150, seed=4)
    p = str(tmp_path / "s.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci)
    for K in (32, 40):
        out = str(tmp_path / ("p%d.bin" % K))
        assert harness("cpu", p, K, out).returncode == 0
        A = oracle.make_data(M * K).reshape(M, K)
        B = oracle.make_data(N * K).reshape(N, K)        # K x N column-major, ld = K: the same default-seeded stream
        want = oracle.sddmm_cpu(M, N, K, A, B, ro, ci, num_threads=1)
        got = np.fromfile(out, dtype=np.float32)
        assert np.array_equal(got, want), K
    a = got.copy()
    b = got * (1 + np.random.default_rng(0).normal(0, 7e-4, size=got.shape)).astype(np.float32)
    a.tofile(str(tmp_path / "a.bin"))
    b.astype(np.float32).tofile(str(tmp_path / "b.bin"))
    r = harness("check", str(tmp_path / "a.bin"), str(tmp_path / "b.bin"))
    errors = int(r.stdout.strip().splitlines()[-1].split("=")[1])
    assert errors == oracle.check_data(a, b.astype(np.float32)) and 0 < errors < len(a)


# ---- the CLI (bsmr-sddmm_b200/host/main.cpp): same flags, exit codes and log keys as the reference's binary ----
CLI = os.path.join(HOST_DIR, "BSMR-sddmm-validate")      # built with -DVALIDATE: check_rphm + checkSddmm after the SDDMM
LOG_KEYS = ["[File : ", "[K : ", "[NNZ : ", "[NumRowPanel : ", "[bsmr_alpha : ", "[bsmr_delta : ", "[bsmr_numClusters : ",
            "[bsmr_numDenseBlock : ", "[bsmr_rowReordering : ", "[bsmr_colReordering : ", "[bsmr_gflops : ", "[bsmr_sddmm : "]


def run_cli(*args):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=300)


def test_cli_rejects_bad_input_and_has_no_cpu_path(pkg, tmp_path):
    import torch
    assert os.path.exists(CLI), "run __graft_entry__.build()"
    r = run_cli("-f", str(tmp_path / "missing.mtx"), "-k", 32)
    assert r.returncode == 255 and "matrix S initialize failed" in r.stderr          # main returns -1 (src/main.cu:20-23)
    if torch.cuda.is_available():
        return
    M, N, ro, ci = pkg.synth.block_structured(64, 96, seed=2)
    p = str(tmp_path / "s.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci)
    r = run_cli("-f", p, "-k", 32)
    assert r.returncode == 255 and "no CPU fallback" in r.stderr and "[bsmr_gflops" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("K,alpha,delta", [(32, 0.3, 0.3), (128, 0.3, 0.3), (64, 0.5, 0.0), (256, 0.1, 1.1)])
def test_cli_validate_run(pkg, tmp_path, K, alpha, delta):
    """The reference's own end-to-end self-check (VALIDATE build, src/sddmm.cu:35-38) through the host mirror on the B200:
    .mtx loader -> makeData -> BSMR -> RPHM -> sddmm_gpu -> check_rphm + checkSddmm (sddmm_cpu, 1e-3 tolerance)."""
    M, N, ro, ci = pkg.synth.block_structured(300, 520, seed=7)
    p = str(tmp_path / "s.mtx")
    pkg.synth.write_mtx(p, M, N, ro, ci, shuffle_seed=1)
    r = run_cli("-f", p, "-k", K, "-a", alpha, "-d", delta)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Pass! Result validates successfully." in r.stdout and "No Pass" not in r.stdout
    assert "incorrect" not in r.stderr
    for key in LOG_KEYS:
        assert key in r.stdout, key
    assert "[K : %d]" % K in r.stdout and "[NNZ : %d]" % len(ci) in r.stdout and "[M : 300], [N : 520]" in r.stdout
