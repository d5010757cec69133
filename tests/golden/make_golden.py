"""Generate tests/golden/kat.json from the UNMODIFIED reference leaf code.

Run in the authoring container (needs /root/reference; builds oracle/_ref/librtk_ref.so via
oracle/Makefile).  Every vector is produced by rtk_trace_ray of the unmodified rtk.c driven through
flat single-leaf blobs; floats are stored as hex bit patterns.

Two result sets are stored per known-answer vector:
  "expect"   scene padded to a multiple of four triangles with far-away dummies, so that the
             reference's group-coupled fp64 promotion (rtk.c:305-336: one exact zero in a group of
             four promotes all four lanes, and zero padding always does) is triggered only by the
             triangle's own edge values -- the canonical own-lane semantics.
  "literal"  the bare scene (padding lanes present): what rtk.c literally returns; differs from
             "expect" by <= 3e-7 relative in t/u/v on K12/K13.
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

INF = 3.402823e+38


def fhex(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


T1 = [[(0, 0, 1), (1, 0, 1), (0, 1, 1)]]
T2 = T1 + [[(1, 0, 1), (1, 1, 1), (0, 1, 1)]]
T3 = T1 + T1
T4 = [[(0.1, 0.2, 3.3), (2.7, -0.4, 2.9), (0.3, 1.9, 4.1)]]
DUMMY = [[(100 + 3 * k, 100, 100), (101 + 3 * k, 100.25, 100), (100 + 3 * k, 101, 100.5)] for k in range(3)]

KATS = [
    ("K1", T1, (0.25, 0.25, 0), (0, 0, 1), 0, INF),
    ("K2", T1, (0.125, 0.25, 0), (0, 0, 1), 0, INF),
    ("K3", T1, (0.25, 0.25, 2), (0, 0, -1), 0, INF),
    ("K4", T1, (0.5, 0, 0), (0, 0, 1), 0, INF),
    ("K5", T1, (0, 0, 0), (0, 0, 1), 0, INF),
    ("K6", T2, (0.5, 0.5, 0), (0, 0, 1), 0, INF),
    ("K7", T3, (0.25, 0.25, 0), (0, 0, 1), 0, INF),
    ("K8", T1, (0.5, -2.0 ** -20, 0), (0, 0, 1), 0, INF),
    ("K9", T1, (0.25, 0.25, 0), (0, 0, 1), 0, 1.0),
    ("K10", T1, (0.25, 0.25, 0), (0, 0, 1), 1.0, INF),
    ("K11", T1, (0.25, 0.25, 0), (0, 0, 4), 0, INF),
    ("K12", T4, (0.3, 0.1, -1), (0.2, 0.15, 1), 0, INF),
    ("K13", T4, (-3, 0.5, 3.4), (1, 0.05, 0.02), 0, INF),
    ("K14", T1, (-1, 0.25, 1), (1, 0, 0), 0, INF),
]


def one_ray(o, d, mn, mx):
    r = np.zeros(1, dtype=orc.RAY_DTYPE)
    r["o"], r["d"], r["min_t"], r["max_t"] = o, d, mn, mx
    return r


def rec(h):
    if h["prim"] == orc.MISS:
        return {"hit": False}
    return {"hit": True, "prim": int(h["prim"]), "t": fhex(h["t"]), "u": fhex(h["u"]), "v": fhex(h["v"])}


def main():
    assert orc.have_reference(), "needs oracle/_ref/librtk_ref.so (make -C oracle with /root/reference present)"
    out = {"source": "unmodified /root/reference/rtk.c via oracle/_ref/librtk_ref.so (flat single-leaf blobs)",
           "kats": [], "random": None}
    for name, tris, o, d, mn, mx in KATS:
        bare = np.asarray(tris, dtype=np.float32)
        pad = (-len(bare)) % 4
        padded = np.concatenate([bare, np.asarray(DUMMY[:pad], dtype=np.float32)]) if pad else bare
        ray = one_ray(o, d, mn, mx)
        lit = orc.trace_flat_reference(bare, ray)[0]
        exp = orc.trace_flat_reference(padded, ray)[0]
        mine = orc.trace_brute(padded, ray)[0]
        assert exp.tobytes() == mine.tobytes(), (name, exp, mine)
        out["kats"].append({"name": name, "tris": [[fhex(c) for v in t for c in v] for t in padded.tolist()],
                            "num_real_tris": len(bare),
                            "ray": {"o": [fhex(x) for x in o], "d": [fhex(x) for x in d], "min_t": fhex(mn), "max_t": fhex(mx)},
                            "expect": rec(exp), "literal": rec(lit)})
    # a random soup: 64 triangles (full groups of four in every 60-chunk), 256 rays
    rng = np.random.default_rng(20261018)
    c = rng.random((64, 1, 3)).astype(np.float32)
    tris = (c + 0.2 * (rng.random((64, 3, 3)).astype(np.float32) * 2 - 1)).astype(np.float32)
    rays = np.zeros(256, dtype=orc.RAY_DTYPE)
    rays["o"] = (rng.random((256, 3)) * 2 - 0.5).astype(np.float32)
    rays["d"] = rng.normal(size=(256, 3)).astype(np.float32)
    rays["max_t"] = INF
    ref = orc.trace_flat_reference(tris, rays)
    mine = orc.trace_brute(tris, rays)
    assert ref.tobytes() == mine.tobytes()
    out["random"] = {"tris": tris.view(np.uint32).reshape(-1).tolist(),
                     "rays": rays.view(np.uint32).reshape(-1).tolist(),
                     "hits": ref.view(np.uint32).reshape(-1).tolist(),
                     "num_hits": int((ref["prim"] != orc.MISS).sum())}
    with open(os.path.join(os.path.dirname(__file__), "kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote kat.json:", len(out["kats"]), "KATs,", out["random"]["num_hits"], "random hits")
    for k in out["kats"]:
        print(k["name"], k["expect"], "" if k["expect"] == k["literal"] else "literal=%s" % k["literal"])


if __name__ == "__main__":
    main()
