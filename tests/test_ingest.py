"""Graph ingest: Matrix Market text -> symmetric CSR without self-loops.
CPU: the numpy restatement (oracle/oracle_py.py: mtx_to_csr) against the UNMODIFIED reference's
own reader + symmetrize + self-loop removal (main.cxx:243-245, compiled into oracle/_ref).
GPU: nlp_ingest_mtx (parse, pair emit, sort, unique, CSR build on the device) against both, and a
prediction on the ingested graph against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O   # noqa: E402


def _mtx_cases():
    rng = np.random.default_rng(7)
    cases = {}
    # general banner, weights, comments, self-loops, repeated and mirrored entries, no newline at the
    # end, an isolated last vertex (rows > largest id used)
    cases["general_small"] = (b"%%MatrixMarket matrix coordinate real general\n% a comment\n%another\n"
                              b"9 8 11\n1 2 0.5\n2 1 1.5\n3 3 2\n  4 7 1e-3\n7 4 1\n5 6 1\n5 6 2\n8 1 1\n2 8 3\n6 6 1\n1 3 7")
    # a blank line and a comment inside the body are skipped here; the reference's OpenMP reader turns a
    # blank line into the edge (0, 0) (strtoull of nothing), so this case is not compared with it
    cases["blank_line"] = b"%%MatrixMarket matrix coordinate real general\n5 5 4\n1 2 1\n\n% note\n2 3 1\n   \n4 4 1\n5 1 1\n"
    # symmetric banner (lower triangle), pattern (no weights), CRLF line ends
    cases["symmetric_pattern_crlf"] = (b"%%MatrixMarket matrix coordinate pattern symmetric\r\n6 6 5\r\n2 1\r\n3 1\r\n4 4\r\n6 5\r\n5 2\r\n")
    # a random graph big enough for several text tiles (8 KB each) and lines straddling tile borders
    n, m = 3000, 40000
    u = rng.integers(1, n + 1, m); v = rng.integers(1, n + 1, m)
    body = "".join("%d %d %d\n" % (a, b, (a * 7 + b) % 5 + 1) for a, b in zip(u, v))
    cases["general_random"] = ("%%%%MatrixMarket matrix coordinate integer general\n%% generated\n%d %d %d\n" % (n, n - 5, m)).encode() + body.encode()
    # tabs and leading blanks
    cases["blanks_tabs"] = b"%%MatrixMarket matrix coordinate real general\n4 4 3\n\t1\t2\t1.0\n   2   3   1.0   \n4 1 2\n"
    # header only
    cases["empty_body"] = b"%%MatrixMarket matrix coordinate real general\n5 5 0\n"
    return cases


CASES = _mtx_cases()
NO_REFERENCE = {"blank_line"}


def _symmetric_banner(text):
    return b"symmetric" in text.split(b"\n", 1)[0]


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_ingest_matches_reference_reader(name, tmp_path):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnlpref_batch.so")):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    if name in NO_REFERENCE:
        pytest.skip("documented deviation (blank lines)")
    text = CASES[name]
    path = tmp_path / (name + ".mtx")
    path.write_bytes(text)
    roff, rkeys = O.ref_read_mtx(str(path), symmetric=False, drop_self_loops=True)
    off, keys = O.mtx_to_csr(text, symmetrize=True, drop_self_loops=True)
    assert np.array_equal(off, roff) and np.array_equal(keys, rkeys)
    # main.cxx's "symmetric" argument skips symmetrizeOmp; self-loops kept
    roff, rkeys = O.ref_read_mtx(str(path), symmetric=True, drop_self_loops=False)
    off, keys = O.mtx_to_csr(text, symmetrize=False, drop_self_loops=False)
    assert np.array_equal(off, roff) and np.array_equal(keys, rkeys)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_ingest_matches_reference_reader(name, tmp_path):
    import nlp_b200 as N
    text = CASES[name]
    p = N.Predictor(0)
    try:
        for symmetrize, drop in ((True, True), (False, False), (False, True)):
            S, M = p.ingest_mtx(text, symmetrize=symmetrize, drop_self_loops=drop)
            off, keys = p.fetch_graph()
            woff, wkeys = O.mtx_to_csr(text, symmetrize=symmetrize, drop_self_loops=drop)
            assert S == len(woff) - 1 and M == len(wkeys), (name, S, M, len(woff) - 1, len(wkeys))
            assert np.array_equal(off, woff) and np.array_equal(keys, wkeys), name
            if name not in NO_REFERENCE and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnlpref_batch.so")):
                path = tmp_path / (name + ".mtx")
                path.write_bytes(text)
                roff, rkeys = O.ref_read_mtx(str(path), symmetric=not symmetrize, drop_self_loops=drop)
                assert np.array_equal(off, roff) and np.array_equal(keys, rkeys), name
        # the ingested graph is the resident graph: predict on it
        if name == "general_random":
            S, M = p.ingest_mtx(text)
            off, keys = p.fetch_graph()
            import parity
            for m, D in (("JC", 0), ("AA", 16), ("CN", 4)):
                err, r, st = parity.check_case(p, O, off, keys, m, D, 2000, tag="ingested")
                assert err is None, err
    finally:
        p.close()


@pytest.mark.gpu
def test_device_ingest_rejects_bad_text():
    import nlp_b200 as N
    p = N.Predictor(0)
    try:
        with pytest.raises(Exception):
            p.ingest_mtx(b"%%MatrixMarket matrix array real general\n3 3\n1\n2\n3\n")          # not a coordinate file
        with pytest.raises(Exception):
            p.ingest_mtx(b"%%MatrixMarket matrix coordinate real general\n3 3 2\n1 2 1\n2 9 1\n")   # id 9 > 3
        S, M = p.ingest_mtx(b"%%MatrixMarket matrix coordinate real general\n3 3 1\n1 2 1\n")  # the handle still works
        assert (S, M) == (4, 2)
    finally:
        p.close()


def test_closed_form_of_the_reference_merge():
    """csrc/ingest.cuh computes the copies of an entry from per-row statistics instead of running the
    reference's merge: x united with y, plus a second copy of every common entry above the smallest
    y that is not in x -- if that y is below the largest x.  Checked against the literal port."""
    import random
    rnd = random.Random(3)
    for _ in range(20000):
        n = rnd.randint(1, 12)
        x = sorted(rnd.sample(range(1, n + 1), rnd.randint(0, n)))
        y = sorted(rnd.sample(range(1, n + 1), rnd.randint(0, n)))
        sx, sy = set(x), set(y)
        only_y = sorted(sy - sx)
        want = sorted(sx | sy)
        if x and only_y and only_y[0] < max(x):
            want = sorted(want + [e for e in sx & sy if e > only_y[0]])
        assert O._set_union_last(x, y) == want, (x, y)


def test_header_parse_on_the_host(tmp_path):
    """csrc/mtx_header.hpp (the host half of nlp_ingest_mtx) compiled and run without a GPU."""
    import subprocess
    exe = tmp_path / "mtx_header_check"
    src = os.path.join(ROOT, "tests", "host", "mtx_header_check.cxx")
    inc = os.path.join(ROOT, "neighborhood-link-prediction-openmp_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", inc, src, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
