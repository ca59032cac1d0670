"""Load the committed golden fixtures (tests/golden/*.npz) and build ORACLE blocks for them.

Mirrors what the reference's jig does (tests/jigs.lua:55-256): instantiate with `args`,
get_rate() patched to 2.0 (:69), differentiate on the input vector types, initialize.
"""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
JIG_RATE = 2.0


def load_spec(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    man = json.loads(str(z["manifest"]))

    def resolve(a):
        if isinstance(a, dict) and "array" in a:
            return z[a["array"]]
        if isinstance(a, list):
            return [resolve(x) for x in a]
        return a

    vectors = []
    for i, v in enumerate(man["vectors"]):
        vectors.append({
            "desc": v["desc"],
            "args": [resolve(a) for a in v["args"]],
            "inputs": [z["v%d_in%d" % (i, j)] for j in range(v["n_in"])],
            "outputs": [z["v%d_out%d" % (i, j)] for j in range(v["n_out"])],
        })
    eps = man["epsilon"]
    # translator/tuner epsilons are Lua expressions "... and 5e-3 or 1e-5": the non-liquid value applies
    epsilon = float(eps) if eps[0].isdigit() else float(eps.rsplit("or", 1)[1])
    return man["block"], vectors, epsilon


def all_block_specs():
    out = []
    for f in sorted(os.listdir(GOLDEN_DIR)):
        if f.endswith("_spec.npz"):
            out.append(f[:-4])
    return out


def epsilon_ok(got, want, eps):
    """tests/jigs.lua:18-26: absolute |x-y| < eps on every component; lengths must match."""
    got = np.asarray(got)
    want = np.asarray(want)
    if got.shape != want.shape:
        return False, "length %s != %s" % (got.shape, want.shape)
    if got.size == 0:
        return True, ""
    if np.iscomplexobj(want):
        d = max(np.max(np.abs(got.real - want.real)), np.max(np.abs(got.imag - want.imag)))
    else:
        d = np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))
    return bool(d < eps), "max abs err %.3g (eps %.1g)" % (d, eps)
