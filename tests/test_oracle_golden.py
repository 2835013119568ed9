"""CPU: the C oracle (oracle/nlp_oracle.c) against the golden vectors generated from the
unmodified reference, and the survey's hand-recorded known-answer bit patterns."""
import numpy as np
import pytest

import golden_util as G
import parity


@pytest.mark.parametrize("name", G.fixture_names())
def test_oracle_matches_golden(oracle, name):
    z = G.load(name)
    off, keys = z["offsets"], z["keys"]
    for m in G.MEASURES:
        for D in G.DEGREES:
            u, v, s, _ = oracle.oracle_predict(off, keys, m, D)
            err = G.check_against(z, m, D, u, v, s)
            assert err is None, "%s %s" % (name, err)


def _bits(x):
    return np.array([x], np.uint32).view(np.float32)[0]


# SURVEY.md section 8c, reference sequential output in canonical order (float bit patterns)
KAT = {
    ("CN", 0): [((1, 2), 0x40000000), ((3, 4), 0x40000000), ((1, 7), 0x3f800000)],
    ("JC", 0): [((1, 2), 0x3f2aaaab), ((3, 4), 0x3f2aaaab), ((1, 7), 0x3eaaaaab), ((3, 5), 0x3eaaaaab), ((5, 7), 0x3eaaaaab)],
    ("JC", 2): [((5, 7), 0x3eaaaaab), ((1, 2), 0x3e800000)],
    ("SI", 0): [((1, 2), 0x3ecccccd), ((3, 4), 0x3ecccccd)],
    ("SC", 0): [((1, 2), 0x3f5105ec), ((3, 4), 0x3f5105ec), ((1, 7), 0x3f000000)],
    ("SC", 2): [((5, 7), 0x3f000000), ((1, 2), 0x3ed105ec)],
    ("HP", 0): [((1, 2), 0x3f800000), ((3, 4), 0x3f800000), ((1, 7), 0x3f000000)],
    ("HD", 0): [((1, 2), 0x3f2aaaab), ((3, 4), 0x3f2aaaab), ((1, 7), 0x3f000000)],
    ("LHN", 0): [((1, 2), 0x3eaaaaab), ((3, 4), 0x3eaaaaab), ((1, 7), 0x3e800000)],
    ("LHN", 2): [((5, 7), 0x3e800000), ((1, 2), 0x3e2aaaab)],
    ("AA", 0): [((1, 2), 0x4016967a), ((3, 4), 0x4016967a), ((2, 6), 0x3fb8aa3b), ((4, 6), 0x3fb8aa3b), ((5, 7), 0x3fb8aa3b), ((1, 7), 0x3f690570)],
    ("AA", 2): [((1, 2), 0x3fb8aa3b)],
    ("RA", 0): [((1, 2), 0x3f555555), ((3, 4), 0x3f555555), ((2, 6), 0x3f000000)],
    ("RA", 2): [((1, 2), 0x3f000000)],
}


@pytest.mark.parametrize("case", sorted(KAT))
def test_oracle_known_answers(oracle, case):
    off, keys = parity.kat_graph()
    u, v, s, _ = oracle.oracle_predict(off, keys, case[0], case[1])
    for i, ((eu, ev), bits) in enumerate(KAT[case]):
        assert (int(u[i]), int(v[i])) == (eu, ev), (case, i)
        assert int(s.view(np.uint32)[i]) == bits, (case, i, hex(int(s.view(np.uint32)[i])))


def test_oracle_tie_rule_top3(oracle):
    """Top-3 Jaccard D=0: the third slot is a 3-way tie at 1/3; canonical order keeps (1,7)."""
    off, keys = parity.kat_graph()
    u, v, s, _ = oracle.oracle_predict(off, keys, "JC", 0, max_edges=3)
    assert list(zip(u.tolist(), v.tolist())) == [(1, 2), (3, 4), (1, 7)]


def test_oracle_edge_cases(oracle):
    # empty graph, max_edges == 0 (inc/predict.hxx:429), min_score filter, isolated vertices
    off = np.zeros(1, np.uint64)
    u, v, s, _ = oracle.oracle_predict(off, np.empty(0, np.uint32), "JC", 4)
    assert len(u) == 0
    off, keys = parity.kat_graph()
    u, v, s, _ = oracle.oracle_predict(off, keys, "JC", 0, max_edges=0)
    assert len(u) == 0
    u, v, s, st = oracle.oracle_predict(off, keys, "JC", 0, min_score=0.3)
    assert len(u) == 5 and st["kept"] == 5 and (s > 0.3).all()
    off2 = np.concatenate([off, np.full(5, off[-1], np.uint64)])       # 5 isolated vertices appended
    u2, v2, s2, _ = oracle.oracle_predict(off2, keys, "JC", 0)
    u1, v1, s1, _ = oracle.oracle_predict(off, keys, "JC", 0)
    assert np.array_equal(u1, u2) and np.array_equal(v1, v2) and np.array_equal(s1, s2)


def test_oracle_thread_count_independent(oracle):
    import nlp_b200 as N
    off, keys = N.graphs.to_numpy(*N.graphs.rmat(9, 8, 3))
    a = oracle.oracle_predict(off, keys, "AA", 0, max_edges=500, threads=1)
    b = oracle.oracle_predict(off, keys, "AA", 0, max_edges=500, threads=4)
    assert parity.compare(a[:3], b[:3]) is None and a[3] == b[3]
