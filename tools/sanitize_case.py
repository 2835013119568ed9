"""Small predictions through every LHub / IHub path, checked against the golden vectors; meant to
be run under compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import golden_util as G
    import nlp_b200 as N
    p = N.Predictor(0)
    bad = 0
    for name in G.fixture_names():
        z = G.load(name)
        p.set_graph(z["offsets"], z["keys"])
        for path in (2, 3, 1):
            p.set_path(path)
            for m in ("JC", "AA", "CN", "SC"):
                for D in G.DEGREES:
                    r = p.predict(m, D)
                    u, v, s = p.fetch(r["count"])
                    err = G.check_against(z, m, D, u, v, s)
                    bad += err is not None
                    print(name, "path", path, "->", r["path"], m, D, r["count"], "OK" if err is None else err, flush=True)
    p.close()
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
