"""Deterministic synthetic workloads (no network, so no SuiteSparse downloads).

Every generator is a pure function of its seed built on a counter-based integer hash
(splitmix64), written with torch integer ops so the same code gives the *same graph* on CPU
and on GPU (bench.py builds the big ones on the device, tests build small ones on the host).

All graphs come out the way the reference's loader leaves them for the predictor
(main.cxx:243-245): symmetric, no self-loops, sorted + deduplicated rows, vertex ids 1-based
with vertex 0 empty (mtx.hxx:240: span = n + 1), as CSR ``offsets`` (int64) / ``keys`` (int32).
"""
import math

import torch

_M64 = (1 << 64) - 1


def _s64(c):
    """Python int -> the signed 64-bit value with the same bit pattern."""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def mix64(x):
    """splitmix64 finaliser on an int64 tensor (wrapping arithmetic, logical shifts)."""
    x = x + _s64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def hash_u64(idx, seed, stream=0):
    """Uniform 64 random bits per element of ``idx`` (int64 tensor), as int64 bit patterns."""
    return mix64(mix64(idx + _s64(seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95)))


def hash_unit(idx, seed, stream=0):
    """Uniform doubles in [0, 1)."""
    return _lsr(hash_u64(idx, seed, stream), 11).to(torch.float64) * (1.0 / (1 << 53))


# ---------------------------------------------------------------------------------------------
def csr_from_pairs(a, b, n):
    """Undirected pairs (a[i], b[i]), 1 <= a,b <= n  ->  symmetric, loop-free, deduplicated CSR.

    Returns (offsets int64[S+1], keys int32[M]) with S = n + 1.
    """
    dev = a.device
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    und = torch.unique(lo * (n + 1) + hi)                       # one key per undirected edge
    lo, hi = und // (n + 1), und % (n + 1)
    del und
    src = torch.cat([lo, hi]); dst = torch.cat([hi, lo])
    del lo, hi
    comp, _ = torch.sort(src * (n + 1) + dst)
    del src, dst
    src = comp // (n + 1)
    keys = (comp % (n + 1)).to(torch.int32)
    del comp
    deg = torch.bincount(src, minlength=n + 1)
    offsets = torch.zeros(n + 2, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=offsets[1:])
    return offsets, keys


_CUB_LIMIT = (1 << 31) - 1024


def sorted_unique(x, limit=_CUB_LIMIT):
    """torch.unique(x) for 1-D int64 tensors of ANY length: torch's CUDA sort refuses more than
    2^31 - 1 elements, so longer inputs are split into value ranges (splitters = quantiles of a
    sample), each range is made unique on its own and the pieces are concatenated in range order."""
    n = x.numel()
    if n <= limit:
        return torch.unique(x)
    parts = 2 * ((n + limit - 1) // limit) + 2
    step = max(1, n // (1 << 20))
    sample, _ = torch.sort(x[::step])
    cut = [int(sample[(i * sample.numel()) // parts]) for i in range(1, parts)]
    cut = sorted(set(cut) | set(c + 1 for c in cut))            # a heavily repeated value gets a range of its own
    bounds = [None] + cut + [None]
    out = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        m = torch.ones(n, dtype=torch.bool, device=x.device) if lo is None else x >= lo
        if hi is not None:
            m &= x < hi
        piece = x[m]
        del m
        if piece.numel() > limit and int(piece.min()) == int(piece.max()):
            out.append(piece[:1].clone())                      # one value, repeated: nothing to split
        else:
            out.append(sorted_unique(piece, limit) if piece.numel() > limit else torch.unique(piece))
        del piece
    return torch.cat(out)


def csr_from_undirected_keys(und, n, chunk=1 << 27):
    """Sorted unique undirected keys lo * (n + 1) + hi (lo < hi)  ->  the same CSR csr_from_pairs
    builds, without ever sorting the 2|E| directed entries: row u = its lower neighbours (found
    by ONE more sort, by (hi, lo)) followed by its upper neighbours (already in order), written
    to their final positions chunk by chunk.  For graphs of billions of entries (BASELINE
    configs[3]), where csr_from_pairs' temporaries do not fit."""
    dev = und.device
    E = und.numel()
    n1 = n + 1
    nup = torch.zeros(n1, dtype=torch.int64, device=dev)     # upper neighbours per vertex (as lo)
    nlow = torch.zeros(n1, dtype=torch.int64, device=dev)    # lower neighbours per vertex (as hi)
    key2 = torch.empty(E, dtype=torch.int64, device=dev)
    for s0 in range(0, E, chunk):
        c = und[s0:s0 + chunk]
        lo, hi = c // n1, c % n1
        nup += torch.bincount(lo, minlength=n1)
        nlow += torch.bincount(hi, minlength=n1)
        key2[s0:s0 + chunk] = hi * n1 + lo
        del lo, hi
    deg = nup + nlow
    offsets = torch.zeros(n1 + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=offsets[1:])
    up_off = torch.cumsum(nup, 0) - nup                        # first index of vertex u's group in und
    low_off = torch.cumsum(nlow, 0) - nlow                     # ... in key2 sorted
    keys = torch.empty(2 * E, dtype=torch.int32, device=dev)
    for s0 in range(0, E, chunk):                              # upper neighbours: und is sorted by (lo, hi)
        c = und[s0:s0 + chunk]
        lo, hi = c // n1, c % n1
        i = torch.arange(s0, s0 + c.numel(), device=dev, dtype=torch.int64)
        keys[offsets[lo] + nlow[lo] + i - up_off[lo]] = hi.to(torch.int32)
        del lo, hi, i
    del und
    key2 = sorted_unique(key2)                                 # already unique: this is the sort by (hi, lo)
    for s0 in range(0, E, chunk):
        c = key2[s0:s0 + chunk]
        hi, lo = c // n1, c % n1
        j = torch.arange(s0, s0 + c.numel(), device=dev, dtype=torch.int64)
        keys[offsets[hi] + j - low_off[hi]] = lo.to(torch.int32)
        del lo, hi, j
    return offsets, keys


def undirected_pairs(offsets, keys):
    """(lo, hi) with lo < hi, one row per undirected edge of a symmetric CSR."""
    S = offsets.numel() - 1
    deg = offsets[1:] - offsets[:-1]
    src = torch.repeat_interleave(torch.arange(S, device=offsets.device, dtype=torch.int64), deg)
    dst = keys.to(torch.int64)
    m = src < dst
    return src[m], dst[m]


def remove_edges(offsets, keys, fraction, seed):
    """Remove ~``fraction`` of the undirected edges (hash-selected, deterministic).

    Stands in for generateEdgeDeletions + applyBatchUpdate (batch.hxx:99-112, 239-247; that
    sampler is host harness, out of the hot path).  Returns (offsets', keys', removed_lo,
    removed_hi); the number of removed edges is the prediction count K (main.cxx:50).
    """
    n = offsets.numel() - 2
    lo, hi = undirected_pairs(offsets, keys)
    r = hash_unit(lo * (n + 1) + hi, seed, 7)
    gone = r < fraction
    o2, k2 = csr_from_pairs(lo[~gone], hi[~gone], n)
    return o2, k2, lo[gone], hi[gone]


# ---------------------------------------------------------------------------------------------
def rmat(scale, edge_factor=16, seed=42, abc=(0.57, 0.19, 0.19), permute=False, device="cpu", chunk=1 << 24):
    """R-MAT (Chakrabarti et al.) with 2**scale vertices and edge_factor * 2**scale edge draws."""
    n = 1 << scale
    m = edge_factor * n
    a, b, c = abc
    ta, tab, tabc = int(a * 65536), int((a + b) * 65536), int((a + b + c) * 65536)
    us, vs = [], []
    for start in range(0, m, chunk):
        cnt = min(chunk, m - start)
        idx = torch.arange(start, start + cnt, device=device, dtype=torch.int64)
        u = torch.zeros(cnt, dtype=torch.int64, device=device)
        v = torch.zeros(cnt, dtype=torch.int64, device=device)
        for lvl in range(scale):
            if lvl % 4 == 0:
                bits = hash_u64(idx, seed, 1 + lvl // 4)
            r = _lsr(bits, 16 * (lvl % 4)) & 0xFFFF
            ubit = (r >= tab).to(torch.int64)                       # quadrants c, d: lower half
            vbit = (((r >= ta) & (r < tab)) | (r >= tabc)).to(torch.int64)   # quadrants b, d
            u = (u << 1) | ubit
            v = (v << 1) | vbit
        us.append(u); vs.append(v)
    u = torch.cat(us); v = torch.cat(vs)
    if permute:
        perm = torch.argsort(hash_u64(torch.arange(n, device=device, dtype=torch.int64), seed, 99))
        u, v = perm[u], perm[v]
    return csr_from_pairs(u + 1, v + 1, n)


def road_lattice(side, keep=0.6, seed=44, device="cpu"):
    """2-D grid, every grid edge kept with probability ``keep`` (avg degree ~ 4*keep)."""
    n = side * side
    idx = torch.arange(n, device=device, dtype=torch.int64)
    x = idx % side
    right = (x < side - 1) & (hash_unit(idx, seed, 1) < keep)
    down = (idx < n - side) & (hash_unit(idx, seed, 2) < keep)
    a = torch.cat([idx[right], idx[down]]) + 1
    b = torch.cat([idx[right] + 1, idx[down] + side]) + 1
    return csr_from_pairs(a, b, n)


def web_crawl(n, avg_out=19, alpha=2.1, window=10000, local=0.9, max_out=None, seed=45, device="cpu",
              chunk=1 << 24, leaves=False):
    """Web-crawl-shaped graph: power-law out-degrees, ``local`` of the links inside a +-window id
    range (host locality), the rest to power-law-popular targets (heavy in-degree hubs).
    ``leaves``: two of every five ids are LEAF pages -- 1 to 4 out-links, never a link target -- the
    low-degree vertices a real crawl is full of (and the only intermediates an LHub threshold of 4
    admits: without them a graph of average degree ~90 has no vertex of degree <= 4 at all)."""
    ids = torch.arange(n, device=device, dtype=torch.int64)
    r = hash_unit(ids, seed, 1)
    dmin = max(1.0, avg_out * (alpha - 2.0) / (alpha - 1.0))
    out = torch.floor(dmin * (1.0 - r) ** (-1.0 / (alpha - 1.0))).to(torch.int64)
    out = torch.clamp(out, max=max_out if max_out else max(64, n // 50))
    if leaves:
        leaf = ((ids % 5) == 1) | ((ids % 5) == 3)
        out = torch.where(leaf, 1 + (hash_u64(ids, seed, 4) & 3), out)
        del leaf
    cs = torch.cumsum(out, 0)
    m = int(cs[-1])
    starts = cs - out
    n1 = n + 1
    und = torch.empty(m, dtype=torch.int64, device=device)         # undirected keys lo * (n + 1) + hi, 0 = self-loop
    for s0 in range(0, m, chunk):
        cnt = min(chunk, m - s0)
        e = torch.arange(s0, s0 + cnt, device=device, dtype=torch.int64)
        src = torch.searchsorted(cs, e, right=True)
        r1, r2 = hash_unit(e, seed, 2), hash_unit(e, seed, 3)
        off = torch.floor((r2 * 2.0 - 1.0) * window).to(torch.int64)
        near = torch.clamp(src + off, 0, n - 1)
        pop = torch.floor(n * r2 ** 4.0).to(torch.int64)             # popular targets: low "rank"
        pop = (pop * 0x9E3779B1 + 12345) % n                          # scatter the hubs over ids
        dst = torch.where(r1 < local, near, pop)
        if leaves:                                                    # leaf pages are never linked to
            res = dst % 5
            dst = dst - ((res == 1) | (res == 3)).to(torch.int64)
            del res
        a, b = src + 1, dst + 1
        key = torch.minimum(a, b) * n1 + torch.maximum(a, b)
        und[s0:s0 + cnt] = torch.where(a != b, key, torch.zeros_like(key))
        del e, src, r1, r2, off, near, pop, dst, a, b, key
    del starts, cs, out
    und = sorted_unique(und)
    if und.numel() and int(und[0]) == 0:
        und = und[1:]
    return csr_from_undirected_keys(und, n)


def planted_partition(n, communities, deg_in=12, deg_out=2, seed=47, device="cpu"):
    """Clustered graph (vertex i in community i % communities): link prediction has a
    non-trivial F1 here, unlike on R-MAT."""
    ids = torch.arange(n, device=device, dtype=torch.int64)
    size = n // communities
    us, vs = [], []
    for k in range(deg_in):
        j = (hash_u64(ids, seed, 10 + k) & 0x7FFFFFFF) % max(size, 1)
        us.append(ids); vs.append(torch.clamp((ids % communities) + j * communities, max=n - 1))
    for k in range(deg_out):
        us.append(ids); vs.append((hash_u64(ids, seed, 50 + k) & 0x7FFFFFFFFFFF) % n)
    return csr_from_pairs(torch.cat(us) + 1, torch.cat(vs) + 1, n)


def duplicate_some_entries(offsets, keys, every=7):
    """Multiset rows: repeat every ``every``-th adjacency entry, like the
    duplicates the reference's symmetrizeOmp leaves behind.  Rows stay sorted; may be asymmetric."""
    M = keys.numel()
    idx = torch.arange(M, device=keys.device, dtype=torch.int64)
    rep = torch.ones(M, dtype=torch.int64, device=keys.device)
    rep[idx % every == 0] = 2
    S = offsets.numel() - 1
    deg = offsets[1:] - offsets[:-1]
    src = torch.repeat_interleave(torch.arange(S, device=keys.device, dtype=torch.int64), deg)
    keys2 = torch.repeat_interleave(keys, rep)
    src2 = torch.repeat_interleave(src, rep)
    deg2 = torch.bincount(src2, minlength=S)
    off2 = torch.zeros(S + 1, dtype=torch.int64, device=keys.device)
    torch.cumsum(deg2, 0, out=off2[1:])
    return off2, keys2


def duplicate_symmetric(offsets, keys, every=5, copies=2):
    """Symmetric multiset rows: every ``every``-th undirected edge (by a hash of its endpoints) is
    stored ``copies`` times in BOTH rows, so entry multiplicities stay symmetric (the shape the
    LHub pair path accepts) while rows hold duplicates.  Rows stay sorted."""
    S = offsets.numel() - 1
    deg = offsets[1:] - offsets[:-1]
    src = torch.repeat_interleave(torch.arange(S, device=keys.device, dtype=torch.int64), deg)
    dst = keys.to(torch.int64)
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    pick = _lsr(hash_u64(lo * (S + 1) + hi, 77), 20) % every == 0
    rep = torch.ones_like(src)
    rep[pick] = copies
    keys2 = torch.repeat_interleave(keys, rep)
    src2 = torch.repeat_interleave(src, rep)
    deg2 = torch.bincount(src2, minlength=S)
    off2 = torch.zeros(S + 1, dtype=torch.int64, device=keys.device)
    torch.cumsum(deg2, 0, out=off2[1:])
    return off2, keys2


def apply_deletions(offsets, keys, del_u, del_v):
    """Remove the directed entries (del_u[i], del_v[i]) from a CSR, ONE stored copy per request, the
    way the reference's applyBatchUpdateOmpU does (inc/batch.hxx:239-247: removeEdge + update; a
    duplicated entry of a multiset row survives its own removal, _algorithm.hxx:132-139).  The
    pairs must be unique (tidyBatchUpdateU, inc/batch.hxx:200-208); pairs that are not stored are
    ignored.  torch reference of ``nlp_apply_deletions`` (used by tests and to build workloads)."""
    S = offsets.numel() - 1
    dev = keys.device
    deg = offsets[1:] - offsets[:-1]
    src = torch.repeat_interleave(torch.arange(S, device=dev, dtype=torch.int64), deg)
    comp = src * S + keys.to(torch.int64)                     # ascending: rows ascending, rows sorted
    want = del_u.to(device=dev, dtype=torch.int64) * S + del_v.to(device=dev, dtype=torch.int64)
    pos = torch.searchsorted(comp, want)                       # first stored copy of every request
    pos = torch.clamp(pos, max=max(comp.numel() - 1, 0))
    hit = pos[comp[pos] == want] if comp.numel() else pos[:0]
    keep = torch.ones(comp.numel(), dtype=torch.bool, device=dev)
    keep[hit] = False
    del comp
    keys2 = keys[keep]
    deg2 = torch.bincount(src[keep], minlength=S)
    off2 = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg2, 0, out=off2[1:])
    return off2, keys2


def to_numpy(offsets, keys):
    import numpy as np
    return (offsets.cpu().numpy().astype(np.uint64), keys.cpu().numpy().astype(np.uint32))


def write_mtx(path, offsets, keys):
    """MatrixMarket ``pattern symmetric`` file, one line per undirected edge (u > v, 1-based ids as
    stored), the on-disk format the reference's loader reads (mtx.hxx:39-54, 151-188; run the
    reference with argv[2] = 1 so it does not symmetrize again)."""
    import numpy as np
    off, k = to_numpy(offsets, keys)
    S = off.shape[0] - 1
    src = np.repeat(np.arange(S, dtype=np.int64), np.diff(off).astype(np.int64))
    dst = k.astype(np.int64)
    keep = src > dst
    rows, cols = src[keep], dst[keep]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n")
        f.write("%d %d %d\n" % (S - 1, S - 1, rows.shape[0]))
        np.savetxt(f, np.stack([rows, cols], 1), fmt="%d %d")
    return int(rows.shape[0])


def describe(offsets, keys):
    deg = offsets[1:] - offsets[:-1]
    return {"span": int(offsets.numel() - 1), "entries": int(keys.numel()), "max_degree": int(deg.max()),
            "avg_degree": float(keys.numel()) / max(1, int((deg > 0).sum()))}
