"""The C++ mirror of the reference API (include/predict_b200.hxx: the 18 predictLinks* templates
and the two structs of inc/predict.hxx:33-102, 502-831) on top of the C ABI.

CPU: the shim compiles against a graph class that offers only the reference's four accessors,
links to the CUDA library, and -- with no GPU -- fails loudly instead of falling back.
GPU: every entry point, for D in {0, 2, 16, 1024}, against the oracle, bit-exact.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tests", "host")
BUILD = os.path.join(HOST, "_build")
EXE = os.path.join(BUILD, "shim_check")


def build_shim_check(nlp):
    lib = nlp.build.build()
    src = os.path.join(HOST, "shim_check.cxx")
    hdr = os.path.join(ROOT, "include", "predict_b200.hxx")
    os.makedirs(BUILD, exist_ok=True)
    if (not os.path.exists(EXE)) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in (src, hdr, lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", os.path.dirname(lib), "-lnlp_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def write_graph(path, off, keys):
    with open(path, "wb") as f:
        f.write(struct.pack("<QQ", len(off) - 1, len(keys)))
        f.write(np.ascontiguousarray(off, np.uint64).tobytes())
        f.write(np.ascontiguousarray(keys, np.uint32).tobytes())


def read_cases(path):
    buf = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(buf):
        m, omp, D, n = struct.unpack_from("<IIIQ", buf, pos)
        pos += 20
        rec = np.frombuffer(buf, dtype=np.dtype([("u", "<u4"), ("v", "<u4"), ("s", "<f4")]), count=n, offset=pos)
        pos += 12 * n
        out.append((m, omp, D, rec["u"].copy(), rec["v"].copy(), rec["s"].copy()))
    return out


def test_shim_compiles_and_fails_loudly_without_gpu(nlp, tmp_path):
    import torch
    exe = build_shim_check(nlp)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU box")
    off, keys = parity.kat_graph()
    write_graph(tmp_path / "g.bin", off, keys)
    p = subprocess.run([exe, str(tmp_path / "g.bin"), str(tmp_path / "o.bin"), "-1"], capture_output=True, text=True)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "no CUDA device" in p.stderr and "no CPU path" in p.stderr, p.stderr


def test_shim_helpers_compile(nlp, tmp_path):
    """Batch generation / evaluation helpers of the C++ mirror, instantiated for a host graph class
    and for DeviceGraph (compile only: running them needs a B200)."""
    src = os.path.join(HOST, "shim_compile_only.cxx")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fopenmp", "-Wall", "-Werror", "-c", "-I", os.path.join(ROOT, "include"),
                           src, "-o", str(tmp_path / "shim_compile_only.o")])


@pytest.mark.gpu
@pytest.mark.parametrize("K", [-1, 500])
def test_shim_matches_oracle(nlp, oracle, tmp_path, K):
    exe = build_shim_check(nlp)
    off, keys = nlp.graphs.to_numpy(*nlp.graphs.rmat(10, 8, 21))
    write_graph(tmp_path / "g.bin", off, keys)
    p = subprocess.run([exe, str(tmp_path / "g.bin"), str(tmp_path / "o.bin"), str(K)], capture_output=True, text=True)
    assert p.returncode == 0, (p.stdout, p.stderr)
    cases = read_cases(tmp_path / "o.bin")
    assert len(cases) == 4 * 18 + 3
    kk = nlp.UNBOUNDED if K < 0 else K
    for i, (m, omp, D, u, v, s) in enumerate(cases):
        k = kk
        if i == 4 * 18:
            k = nlp.UNBOUNDED          # default options
        if i == 4 * 18 + 2:
            k = 100
        want = oracle.oracle_predict(off, keys, m, D, max_edges=k)[:3]
        assert parity.compare((u, v, s), want, "shim m=%d omp=%d D=%d" % (m, omp, D)) is None
