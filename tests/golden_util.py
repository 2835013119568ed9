"""Read tests/golden/*.npz (made by tests/golden/make_golden.py from the compiled reference)."""
import glob
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEGREES = [0, 2, 4, 16]
MEASURES = ["CN", "JC", "SI", "SC", "HP", "HD", "LHN", "AA", "RA"]


def fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def digest(u, v, s):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(u, np.uint32).tobytes())
    h.update(np.ascontiguousarray(v, np.uint32).tobytes())
    h.update(np.ascontiguousarray(s, np.float32).view(np.uint32).tobytes())
    return h.hexdigest()


def check_against(z, measure, D, u, v, s):
    """Compare a full canonical candidate list with the stored golden vector; None if equal."""
    tag = "%s_%d" % (measure, D)
    cnt = int(z[tag + "_count"][0])
    if len(u) != cnt:
        return "%s: count %d, golden %d" % (tag, len(u), cnt)
    gu, gv, gs = z[tag + "_u"], z[tag + "_v"], z[tag + "_s"]
    n = len(gu)
    if not (np.array_equal(u[:n], gu) and np.array_equal(v[:n], gv) and
            np.array_equal(np.ascontiguousarray(s[:n]).view(np.uint32), gs)):
        return "%s: leading %d rows differ from golden" % (tag, n)
    if digest(u, v, s) != str(z[tag + "_sha"][0]):
        return "%s: sha256 of the full list differs from golden" % tag
    return None
