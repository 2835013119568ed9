"""Data-parallel formulation of the reference's random edge removal (inc/batch.hxx:99-112), in
numpy -- the algorithm a GPU kernel would run, checked here against the sequential oracle
(oracle/batch_oracle.c) and the compiled reference.  TEST INFRASTRUCTURE: prototype and parity
anchor for SURVEY.md section 8f-3, not part of the product path.

The reference draws its batch from ONE sequential random stream, and how much of the stream a
deletion consumes depends on the stream itself (a draw that lands on an isolated vertex is retried,
up to five times, and skips its second draw).  Two observations make it parallel all the same:

1. std::default_random_engine is the Lehmer generator x <- 16807 x mod (2^31 - 1), so the state at
   word k is seed * 16807^k mod (2^31 - 1): every position of the stream can be computed on its own
   (one modular exponentiation), and with it the double that generate_canonical<double, 53> makes
   from words 2p and 2p + 1 ("slot" p).
2. What a deletion does is a function of the slot it STARTS at: which vertex / entry it picks (if
   any) and at which slot the next deletion starts.  Compute that function for every slot
   (independent work), then the starts of the batch are 0, next(0), next(next(0)), ...: the orbit
   of slot 0, which pointer doubling finds in O(log batchSize) rounds.
"""
import numpy as np

M31 = np.uint64(2147483647)
A = np.uint64(16807)


def _powmod(exp):
    """16807**exp mod (2^31 - 1), elementwise (exp: uint64 array)."""
    result = np.ones_like(exp, dtype=np.uint64)
    base = np.full_like(exp, A, dtype=np.uint64)
    e = exp.copy()
    for _ in range(40):
        odd = (e & np.uint64(1)).astype(bool)
        result[odd] = (result[odd] * base[odd]) % M31
        base = (base * base) % M31
        e >>= np.uint64(1)
        if not e.any():
            break
    return result


def slot_doubles(seed, nslots):
    """canonical doubles of slots 0 .. nslots-1 of default_random_engine(seed): slot p is made
    from engine outputs 2p+1 and 2p+2 (the engine advances before it returns)."""
    s0 = np.uint64(seed % 2147483647) or np.uint64(1)
    p = np.arange(nslots, dtype=np.uint64)
    x1 = (s0 * _powmod(2 * p + 1)) % M31
    x2 = (x1 * A) % M31
    prod = (x2 - np.uint64(1)).astype(np.float64) * 2147483646.0      # rounded product, then the add
    total = (x1 - np.uint64(1)).astype(np.float64) + prod
    d = total / 4611686009837453312.0                                  # double(2147483646^2) = 2^62 - 2^33
    return np.where(d >= 1.0, np.nextafter(1.0, 0.0), d)


def edge_deletions_parallel(offsets, keys, seed, batch_size):
    """Same result as oracle_edge_deletions / the reference, computed slot-parallel."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    keys = np.asarray(keys, dtype=np.uint32)
    span = offsets.shape[0] - 1
    if batch_size == 0 or span <= 1:
        return np.empty(0, np.uint32), np.empty(0, np.uint32), 0
    nslots = 6 * batch_size + 8                       # a deletion uses at most 4 failures + 2 = 6 slots
    d = slot_doubles(seed, nslots + 8)
    deg = (offsets[1:] - offsets[:-1]).astype(np.int64)
    # what an ATTEMPT starting at slot p does
    u_at = (1.0 + float(span - 1) * d).astype(np.uint32)               # K(i + n*dis(rnd)), i = 1, n = span-1
    ok_at = deg[np.minimum(u_at, span - 1)] > 0
    # what a DELETION starting at slot p does: first successful attempt among p .. p+4
    P = nslots
    idx = np.arange(P)
    win = np.stack([ok_at[idx + k] for k in range(5)], axis=1)         # [P, 5]
    any_ok = win.any(axis=1)
    first = np.argmax(win, axis=1)                                     # index of the first success
    hit = idx + first                                                  # slot of the successful vertex draw
    nxt = np.where(any_ok, hit + 2, idx + 5)                           # start slot of the next deletion
    # orbit of slot 0 by pointer doubling: start[l] = nxt^l(0)
    starts = np.zeros(batch_size, dtype=np.int64)
    jump = np.minimum(nxt, P - 1)                                      # nxt^(2^r), clamped (clamped slots are never used)
    known = 1                                                          # starts[0 .. known) are final
    while known < batch_size:
        take = min(known, batch_size - known)
        starts[known:known + take] = jump[starts[:take]]               # starts[l + known] = nxt^known(starts[l])
        jump = jump[jump]
        known += take
    su = starts
    good = any_ok[su]
    hs = hit[su][good]
    u = u_at[hs]
    vi = (d[hs + 1] * deg[u].astype(np.float64)).astype(np.uint32)     # K(dis(rnd) * deg(u))
    v = keys[offsets[u].astype(np.int64) + vi.astype(np.int64)]
    both = np.concatenate([(u.astype(np.uint64) << np.uint64(32)) | v.astype(np.uint64),
                           (v.astype(np.uint64) << np.uint64(32)) | u.astype(np.uint64)])
    both = np.unique(both)                                             # sortEdgesByIdU + uniqueEdgesU
    last = int(nxt[su[-1]])                                            # slots consumed by the whole batch
    return (both >> np.uint64(32)).astype(np.uint32), (both & np.uint64(0xffffffff)).astype(np.uint32), 2 * last
