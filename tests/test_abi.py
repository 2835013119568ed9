"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol the header
declares; without a GPU it refuses loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nlp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nlp_[a-z_0-9]+)\s*\(", text)))


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_library_exports_every_declared_symbol(nlp):
    lib = nlp.binding.load_library()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libnlp_b200.so does not export %s" % s
    assert set(syms) == set(nlp.binding.EXPORTS)
    assert b"sm_100a" in lib.nlp_version()


def test_struct_layout_matches_header(nlp, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "nlp_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu\\n", sizeof(nlp_options), sizeof(nlp_result), sizeof(nlp_evaluation));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(nlp.binding.Options) and b == C.sizeof(nlp.binding.Result)
    assert c == C.sizeof(nlp.binding.Evaluation)


def test_no_cpu_fallback(nlp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nlp.NlpError) as e:
        nlp.Predictor(0)
    assert e.value.code == 2 and "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neighborhood-link-prediction-openmp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hxx", ".cxx")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "nlpref" not in text, f
