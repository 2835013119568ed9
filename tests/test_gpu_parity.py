"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the golden
vectors of the reference -- bit-exact for counts, scores and the predicted edge set."""
import numpy as np
import pytest

import golden_util as G
import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pred(nlp):
    p = nlp.Predictor(0)
    yield p
    p.close()


def _graph(nlp, name):
    g = nlp.graphs
    table = {
        "rmat12": lambda: g.rmat(12, 16, 31),
        "rmat14p": lambda: g.rmat(14, 16, 32, permute=True),
        "road60": lambda: g.road_lattice(60, 0.6, 33),
        "pp4k": lambda: g.planted_partition(4000, 50, 10, 2, 34),
        "pp4k_multiset": lambda: g.duplicate_some_entries(*g.planted_partition(4000, 50, 10, 2, 35), every=4),
        "pp4k_symdup": lambda: g.duplicate_symmetric(*g.planted_partition(4000, 50, 10, 2, 35), every=4, copies=3),
        "web20k": lambda: g.web_crawl(20000, 10, window=500, seed=36),
    }
    return g.to_numpy(*table[name]())


@pytest.mark.parametrize("name", G.fixture_names())
def test_gpu_matches_reference_golden(pred, name):
    z = G.load(name)
    pred.set_graph(z["offsets"], z["keys"])
    for m in G.MEASURES:
        for D in G.DEGREES:
            r = pred.predict(m, D)
            u, v, s = pred.fetch(r["count"])
            err = G.check_against(z, m, D, u, v, s)
            assert err is None, "%s %s" % (name, err)


SOURCE_PATH, PAIR_PATH, PAIR_SORT_PATH = 1, 2, 3     # nlp_path: source-centric kernels, bucket path, global-sort pair path


@pytest.mark.parametrize("name", ["rmat12", "rmat14p", "road60", "pp4k", "pp4k_multiset", "pp4k_symdup", "web20k"])
@pytest.mark.parametrize("D", [0, 2, 4, 32, 1024])
def test_gpu_matches_oracle(pred, oracle, nlp, name, D):
    """Both scoring paths (source-centric kernels; LHub pair path) against the oracle."""
    off, keys = _graph(nlp, name)
    pred.set_graph(off, keys)
    K = max(3, len(keys) // 20)
    try:
        for path in (SOURCE_PATH, PAIR_PATH, PAIR_SORT_PATH) if D else (SOURCE_PATH,):
            pred.set_path(path)
            for m in nlp.MEASURES:
                for k in (K, nlp.UNBOUNDED) if (D != 0 or name != "rmat14p") else (K,):
                    err, r, st = parity.check_case(pred, oracle, off, keys, m, D, k, tag="%s path%d" % (name, path))
                    assert err is None, err
                    # asymmetric multiset rows are not admissible for the pair path: it must fall back
                    want = SOURCE_PATH if (path == SOURCE_PATH or name == "pp4k_multiset") else path
                    assert r["path"] == want, (name, D, m, r["path"])
    finally:
        pred.set_path(0)


def test_tie_rule_and_known_answer(pred):
    off, keys = parity.kat_graph()
    pred.set_graph(off, keys)
    r = pred.predict("JC", 0, max_edges=3)
    u, v, s = pred.fetch(r["count"])
    assert list(zip(u.tolist(), v.tolist())) == [(1, 2), (3, 4), (1, 7)]
    assert s.view(np.uint32).tolist() == [0x3f2aaaab, 0x3f2aaaab, 0x3eaaaaab]
    r = pred.predict("AA", 0)
    u, v, s = pred.fetch(r["count"])
    assert s.view(np.uint32).tolist()[:3] == [0x4016967a, 0x4016967a, 0x3fb8aa3b]


def test_edge_cases(pred, oracle, nlp):
    # empty graph / no edges / max_edges = 0 / min_score / max_factor2 / arbitrary D
    pred.set_graph(np.zeros(1, np.uint64), np.empty(0, np.uint32))
    assert pred.predict("JC", 4)["count"] == 0
    pred.set_graph(np.zeros(11, np.uint64), np.empty(0, np.uint32))
    assert pred.predict("CN", 0)["count"] == 0
    off, keys = parity.kat_graph()
    pred.set_graph(off, keys)
    assert pred.predict("JC", 0, max_edges=0)["count"] == 0
    off, keys = _graph(nlp, "rmat12")
    pred.set_graph(off, keys)
    for m, D, kw in (("JC", 0, dict(min_score=0.25)), ("CN", 7, dict(min_score=2.0)), ("HP", 0, dict(max_factor2=2)),
                     ("AA", 3, dict(max_factor2=1)), ("SC", 5, dict(min_score=-1.0)), ("RA", 100000, {})):
        err, r, st = parity.check_case(pred, oracle, off, keys, m, D, 5000, tag="edge", **kw)
        assert err is None, err
    with pytest.raises(nlp.NlpError):
        pred.predict(11, 4)


def test_pruned_buffer_passes(pred, oracle, nlp):
    """Force the candidate buffer to be far smaller than the candidate set: sources are admitted
    while there is room, the buffer is cut to the best K (threshold), deferred sources follow."""
    off, keys = _graph(nlp, "rmat14p")
    pred.set_graph(off, keys)
    try:
        K = 2000
        S = len(off) - 1
        pred.set_scratch_limit((K + S + 4096 + 50000) * 24 * 10 // 8 + (64 << 20))
        for m, D in (("CN", 0), ("JC", 0), ("AA", 0), ("JC", 1024)):
            err, r, st = parity.check_case(pred, oracle, off, keys, m, D, K, tag="pruned")
            assert err is None, err
            if D == 0:
                assert r["passes"] > 1, r
    finally:
        pred.set_scratch_limit(0)


def test_partitions_merge_to_single_gpu_result(pred, oracle, nlp):
    """Two ranks emulated one after the other on one GPU: local top-K of each source partition,
    concatenated and merged with nlp_merge, equals the unpartitioned result."""
    import torch
    off, keys = _graph(nlp, "rmat12")
    pred.set_graph(off, keys)
    K = 3000
    for m, D in (("JC", 0), ("AA", 4), ("CN", 16)):
        parts = []
        for rank in range(3):
            pred.set_partition(rank, 3)
            r = pred.predict(m, D, max_edges=K)
            parts.append(pred.fetch(r["count"]))
        pred.set_partition(0, 1)
        u = torch.from_numpy(np.concatenate([p[0] for p in parts]).view(np.int32)).cuda()
        v = torch.from_numpy(np.concatenate([p[1] for p in parts]).view(np.int32)).cuda()
        s = torch.from_numpy(np.concatenate([p[2] for p in parts])).cuda()
        pred.merge(u.data_ptr(), v.data_ptr(), s.data_ptr(), u.numel(), K)
        got = pred.fetch(min(K, u.numel()))
        want = oracle.oracle_predict(off, keys, m, D, max_edges=K)[:3]
        assert parity.compare(got, want, "merge %s D=%d" % (m, D)) is None


def test_properties_at_scale(pred, nlp):
    """BASELINE-sized properties that need no oracle: sorted canonical order, u < v, no existing
    edge predicted, idempotence, top-K is a prefix of top-2K, counters consistent."""
    import torch
    g = nlp.graphs
    o, k = g.rmat(20, 16, 43, permute=True, device="cuda")
    o, k, rl, rh = g.remove_edges(o, k, 0.1, 1043)
    K = int(rl.numel())
    S = o.numel() - 1
    pred.set_graph_pointers(o.data_ptr(), k.data_ptr(), S, device=True, keep=(o, k))
    for m in ("JC", "AA"):
        r = pred.predict(m, 16, max_edges=K)
        assert r["count"] == K and r["kept"] >= K
        u, v, s = pred.fetch(K)
        assert (u < v).all()
        key = np.lexsort((v, u, -s.astype(np.float64)))
        assert (key == np.arange(K)).all(), "result not in canonical order"
        # no predicted pair is an existing edge
        comp = torch.from_numpy(u.astype(np.int64) * (S + 1) + v.astype(np.int64)).cuda()
        deg = (o[1:] - o[:-1])
        src = torch.repeat_interleave(torch.arange(S, device="cuda"), deg)
        ecomp = src * (S + 1) + k.to(torch.int64)
        assert not torch.isin(comp, ecomp).any()
        r2 = pred.predict(m, 16, max_edges=K)
        u2, v2, s2 = pred.fetch(K)
        assert parity.compare((u2, v2, s2), (u, v, s), "idempotence") is None
        r3 = pred.predict(m, 16, max_edges=K // 2)
        u3, v3, s3 = pred.fetch(K // 2)
        assert parity.compare((u3, v3, s3), (u[:K // 2], v[:K // 2], s[:K // 2]), "prefix") is None
        assert r["first_hop"] == int(k.numel()) and r["wedges"] >= r["candidates"] >= r["kept"]


def test_paths_agree_at_scale(nlp, monkeypatch):
    """Independent code paths must give the same bits at a size no CPU oracle finishes quickly:
    LHub through the source-centric kernels, the bucket path and the global-sort pair path (R-MAT 20, D = 16), and IHub
    common neighbours / Jaccard with word counters and with half-word counters in k_range
    (R-MAT 18); the wedge counter equals sum of deg^2 (W(0))."""
    import torch
    g = nlp.graphs
    o, k = g.rmat(20, 16, 43, permute=True, device="cuda")
    o, k, rl, rh = g.remove_edges(o, k, 0.1, 1043)
    K = int(rl.numel())
    p = nlp.Predictor(0)
    try:
        p.set_graph_pointers(o.data_ptr(), k.data_ptr(), o.numel() - 1, device=True, keep=(o, k))
        for m in ("JC", "AA", "LHN"):
            res = []
            for path in (SOURCE_PATH, PAIR_PATH, PAIR_SORT_PATH):
                p.set_path(path)
                r = p.predict(m, 16, max_edges=K)
                assert r["path"] == path
                res.append((r, p.fetch(r["count"])))
            for other in (1, 2):
                assert parity.compare(res[0][1], res[other][1], "paths %s (source vs %d)" % (m, other + 1)) is None
                for c in ("wedges", "candidates", "kept", "eligible_first_hop"):
                    assert res[0][0][c] == res[other][0][c], (m, c, other)
    finally:
        p.close()
    o, k = g.rmat(18, 16, 42, device="cuda")
    o, k, rl, rh = g.remove_edges(o, k, 0.01, 1042)
    K = int(rl.numel())
    deg = (o[1:] - o[:-1])
    want_wedges = int((deg * deg).sum())
    out = {}
    for half in ("1", "0"):
        monkeypatch.setenv("NLP_B200_RANGE_HALF", half)
        p = nlp.Predictor(0)
        try:
            p.set_graph_pointers(o.data_ptr(), k.data_ptr(), o.numel() - 1, device=True, keep=(o, k))
            for m in ("CN", "JC"):
                r = p.predict(m, 0, max_edges=K)
                assert r["wedges"] == want_wedges and r["bin_sources"][6] > 0
                out[(half, m)] = (r, p.fetch(r["count"]))
        finally:
            p.close()
    for m in ("CN", "JC"):
        assert parity.compare(out[("1", m)][1], out[("0", m)][1], "half vs word counters %s" % m) is None
        assert out[("1", m)][0]["candidates"] == out[("0", m)][0]["candidates"]


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_fetch_async_double_buffered(pred, nlp):
    """nlp_fetch_async: results staged on the GPU and transferred on a second stream while the
    next predictions run; three transfers in a row (more than the two staging slots)."""
    import torch
    off, keys = _graph(nlp, "rmat12")
    pred.set_graph(off, keys)
    cases = [("JC", 4, 4000), ("AA", 0, 3000), ("CN", 16, 5000)]
    want, bufs = [], []
    for m, D, K in cases:
        r = pred.predict(m, D, max_edges=K)
        want.append(pred.fetch(r["count"]))
    for m, D, K in cases:
        r = pred.predict(m, D, max_edges=K)
        n = r["count"]
        b = (torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
             torch.empty(n, dtype=torch.float32).pin_memory())
        pred.fetch_async(b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), n)
        bufs.append(b)
    pred.fetch_wait()
    for (m, D, K), w, b in zip(cases, want, bufs):
        got = (b[0].numpy().view(np.uint32), b[1].numpy().view(np.uint32), b[2].numpy())
        assert parity.compare(got, w, "fetch_async %s D=%d" % (m, D)) is None


def _hub_graph(nlp, n=250_000, hubs=6, hub_deg=3000, bg_deg=5, seed=71):
    """A few hubs (ids spread over the whole range) in a sparse random background: their 2-hop
    neighbourhoods are far too large for the shared-memory hash tables, and n is large enough for
    several k_range windows (52 K word counters, 104 K half-word counters per window)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    a = [torch.randint(1, n + 1, (n * bg_deg // 2,), generator=g)]
    b = [torch.randint(1, n + 1, (n * bg_deg // 2,), generator=g)]
    ids = torch.linspace(1, n - 5, hubs).to(torch.int64)
    for h in ids.tolist():
        a.append(torch.full((hub_deg,), h, dtype=torch.int64))
        b.append(torch.randint(1, n + 1, (hub_deg,), generator=g))
    return nlp.graphs.to_numpy(*nlp.graphs.csr_from_pairs(torch.cat(a), torch.cat(b), n))


@pytest.mark.parametrize("half", ["1", "0", "multiset", "nofence", "noquarter"])
def test_hub_heavy_sources_windowed_counters(nlp, oracle, monkeypatch, half):
    """k_range (hub-heavy sources on windowed shared-memory counters): several windows per source,
    word, half-word and byte counters (NLP_B200_RANGE_HALF / _QUARTER), long rows through their fence
    tables and through cursor records (NLP_B200_RANGE_FENCE), IHub and LHub (first-hop lists cut
    into pieces), against the oracle."""
    monkeypatch.setenv("NLP_B200_RANGE_HALF", "0" if half == "0" else "1")
    monkeypatch.setenv("NLP_B200_RANGE_FENCE", "0" if half == "nofence" else "1")
    monkeypatch.setenv("NLP_B200_RANGE_QUARTER", "0" if half == "noquarter" else "1")
    off, keys = _hub_graph(nlp)
    if half == "multiset":     # rows with repeated entries: counts exceed the simple-graph bound, the
        import torch           # half-word limit must follow the largest multiplicity (3 here)
        o2, k2 = nlp.graphs.duplicate_symmetric(torch.from_numpy(off.astype(np.int64)), torch.from_numpy(keys.astype(np.int32)),
                                                every=3, copies=3)
        off, keys = nlp.graphs.to_numpy(o2, k2)
    p = nlp.Predictor(0)
    try:
        p.set_graph(off, keys)
        p.set_path(SOURCE_PATH)
        for m, D, K in (("CN", 0, 20000), ("JC", 0, 60000), ("HP", 1024, 20000), ("LHN", 5, 20000),
                        ("AA", 0, 20000), ("RA", 0, 60000), ("AA", 1024, 20000), ("RA", 5, 20000)):
            err, r, st = parity.check_case(p, oracle, off, keys, m, D, K, tag="hubs half=%s" % half)
            assert err is None, err
            flt = m in ("AA", "RA")
            if D != 5 and not flt:
                assert r["bin_sources"][6] > 0, r["bin_sources"]
            if flt and half == "multiset":
                # repeated entries of a row would collide inside a window of k_range_flt: dense tables instead
                assert r["bin_sources"][6] == 0, r["bin_sources"]
    finally:
        p.close()


def test_float_measures_on_warp_windows(oracle, nlp):
    """k_range_flt: hub-heavy sources of Adamic-Adar / resource allocation on per-warp windows of
    float accumulators, rows taken one after the other (the reference's accumulation order) -- IHub
    and LHub with a high threshold, also with the pruned candidate buffer, against the oracle; the
    dense-table path (NLP_B200_RANGE_FLT=0 is the same code with that bin empty) gives the same bits."""
    g = nlp.graphs
    off, keys = g.to_numpy(*g.rmat(14, 16, 32))            # ids as generated: hubs at the low ids, high-degree neighbours
    S = len(off) - 1
    p = nlp.Predictor(0)
    try:
        p.set_graph(off, keys)
        p.set_path(SOURCE_PATH)
        for K, limit in ((5000, 0), (2000, (2000 + S + 4096 + 50000) * 24 * 10 // 8 + (64 << 20))):
            p.set_scratch_limit(limit)
            for m, D in (("AA", 0), ("RA", 0), ("AA", 1024), ("RA", 256)):
                err, r, st = parity.check_case(p, oracle, off, keys, m, D, K, tag="flt windows K=%d" % K)
                assert err is None, err
                assert r["bin_sources"][6] > 0, r["bin_sources"]
    finally:
        p.close()


def test_reuse_across_measures(pred, oracle, nlp):
    """nlp_set_reuse: the distinct pairs of a threshold with their counts after the exclusion feed
    every later count measure at that threshold (main.cxx:212-220 order: measure-major, thresholds
    inside; the seven count measures share their counts); results stay bit-exact and the later
    predictions skip gather, sort and exclusion.  Adjacent pairs survive with count 0 (min_score < 0)."""
    off, keys = _graph(nlp, "pp4k_symdup")
    pred.set_graph(off, keys)
    K = len(keys) // 10
    try:
        pred.set_reuse(True)
        seen = set()
        for m in nlp.MEASURES:
            for D in (16, 24):       # 68 and 3086 sources with wedges (D = 2 has none on this graph: nothing to keep)
                err, r, st = parity.check_case(pred, oracle, off, keys, m, D, K, tag="reuse")
                assert r["pair_records"] > 0
                assert err is None, err
                assert r["path"] == PAIR_PATH
                flt = m in ("AA", "RA")
                if D in seen and not flt:     # served from the store: no bucket kernel, only back-to-back event records
                    assert r["bin_sources"][7] == 1 and r["phase_ms"][1] < 0.02, (m, D, r["bin_sources"], r["phase_ms"])
                else:
                    assert r["bin_sources"][7] == 0
                if not flt:
                    seen.add(D)
        err, r, st = parity.check_case(pred, oracle, off, keys, "SC", 16, K, min_score=-1.0, tag="reuse min_score<0")
        assert err is None and r["bin_sources"][7] == 1, (err, r["bin_sources"])
        # a new graph empties the store
        off2, keys2 = _graph(nlp, "rmat12")
        pred.set_graph(off2, keys2)
        err, r, st = parity.check_case(pred, oracle, off2, keys2, "JC", 16, 3000, tag="reuse-newgraph")
        assert err is None, err
        assert r["bin_sources"][7] == 0
        err, r, st = parity.check_case(pred, oracle, off2, keys2, "HD", 16, 3000, tag="reuse-newgraph")
        assert err is None and r["bin_sources"][7] == 1
    finally:
        pred.set_reuse(False)


def test_repeat_semantics(pred, oracle, nlp):
    """repeat > 1 (inc/predict.hxx:426-430): the scoring phase runs `repeat` times, scoringTime is
    the MEAN over the repeats, the merge runs once, time = scoringTime + merge; the result is the
    same as with repeat = 1 -- on both LHub paths and on the IHub kernels."""
    off, keys = _graph(nlp, "rmat12")
    pred.set_graph(off, keys)
    try:
        for m, D, path in (("JC", 4, PAIR_PATH), ("AA", 16, PAIR_PATH), ("JC", 4, SOURCE_PATH), ("CN", 0, SOURCE_PATH), ("RA", 0, SOURCE_PATH)):
            pred.set_path(path)
            r1 = pred.predict(m, D, max_edges=4000, repeat=1)
            a = pred.fetch(r1["count"])
            r3 = pred.predict(m, D, max_edges=4000, repeat=3)
            b = pred.fetch(r3["count"])
            assert parity.compare(b, a, "repeat=3 vs 1 %s D=%d" % (m, D)) is None
            want = oracle.oracle_predict(off, keys, m, D, max_edges=4000)[:3]
            assert parity.compare(b, want, "repeat=3 %s D=%d" % (m, D)) is None
            for c in ("wedges", "candidates", "kept", "first_hop", "eligible_first_hop"):
                assert r3[c] == r1[c], (m, D, c, r3[c], r1[c])      # counters are per repeat, not summed
            assert r3["scoring_ms"] > 0 and r3["time_ms"] >= r3["scoring_ms"]
            assert abs(r3["time_ms"] - (r3["scoring_ms"] + r3["select_ms"])) < 1e-3
    finally:
        pred.set_path(0)


def test_lhub_float_long_rows_pruned_buffer(oracle, nlp):
    """LHub float measures (ordered accumulation), first-hop rows longer than 2048 entries (cut into
    chunks by the frontier) and the pruned candidate buffer (passes > 1), all at once, on the
    source-centric kernels; and the same requests on the default path."""
    g = nlp.graphs
    off, keys = g.to_numpy(*g.rmat(15, 16, 37))
    deg = np.diff(off.astype(np.int64))
    assert deg.max() > 2048, deg.max()
    K = 3000
    S = len(off) - 1
    p = nlp.Predictor(0)
    try:
        p.set_graph(off, keys)
        for path in (SOURCE_PATH, 0):
            p.set_path(path)
            p.set_scratch_limit((K + S + 4096 + 60000) * 24 * 10 // 8 + (64 << 20) if path == SOURCE_PATH else 0)
            for m, D in (("AA", 64), ("RA", 1024), ("AA", 1024), ("JC", 1024)):
                err, r, st = parity.check_case(p, oracle, off, keys, m, D, K, tag="lhub-float-long-pruned path%d" % path)
                assert err is None, err
                if path == SOURCE_PATH and D == 1024:
                    assert r["passes"] > 1, r
    finally:
        p.close()


def test_bad_graphs_are_rejected(pred, nlp):
    """nlp_set_graph validates the CSR on the device (include/nlp_b200.h: NLP_ERR_ARG): offsets not
    non-decreasing, a key not below span, a row not sorted -- the kernels index with the keys and
    bisect the rows, so such input must not get past nlp_set_graph.  A good graph works afterwards."""
    off, keys = parity.kat_graph()
    bad_off = off.copy(); bad_off[4] = off[3] - 1
    bad_key = keys.copy(); bad_key[5] = 8
    unsorted = keys.copy(); unsorted[0], unsorted[1] = keys[1], keys[0]
    for o, k, what in ((bad_off, keys, "offsets"), (off, bad_key, "span"), (off, unsorted, "sorted")):
        with pytest.raises(nlp.NlpError) as e:
            pred.set_graph(o, k)
        assert e.value.code == 1 and what in str(e.value), str(e.value)
        with pytest.raises(nlp.NlpError):
            pred.predict("JC", 4)
    pred.set_graph(off, keys)
    assert pred.predict("JC", 0, max_edges=3)["count"] == 3
