#!/usr/bin/env python3
"""Golden vector for the WHOLE examples/rtlsdr_wbfm_mono.lua chain, produced by the reference itself: the stock reference's
pure-Lua process() branches, its TunerBlock composite and its CompositeBlock run loop are executed from /root/reference in
this repo's test interpreter (tests/lua_interp.py, tests/lua_reference_env.py: float32-faithful sample cells), fed three
ragged vectors of synthetic FM I/Q; what the reference's sink received is stored next to the input.

    python tests/golden/make_chain_golden.py            # writes tests/golden/wbfm_chain_ref_executed.npz

(The reference's own spec vectors stop at single blocks -- SURVEY.md 8c; LuaJIT is not installed here, hence the interpreter.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    os.environ["LUARADIO_DISABLE_CUDA"] = "1"
    from tests import lua_reference_env as E
    from tests.lua_interp import to_lua
    from tests.test_lua_reference import REF_CHAIN
    it, types = E.make_env(lib=None, cuda=False)
    rate, n = 1102500.0, 1650
    rng = np.random.default_rng(11)
    t = np.arange(n)
    audio = np.sin(2 * np.pi * 3000.0 / rate * t)
    x = (0.9 * np.exp(1j * (2 * np.pi * 250e3 / rate * t + 2 * np.pi * 75e3 / rate * np.cumsum(audio)))
         + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    run = it.run(REF_CHAIN)[0]
    splits = [0, 700, 701, n]
    parts = [x[a:b] for a, b in zip(splits[:-1], splits[1:])]
    got = it.call(run, [to_lua([it.f32.vector_from_numpy(p) for p in parts])])[0]
    y = np.concatenate([it.f32.to_numpy(v) for v in got.array()])
    out = os.path.join(ROOT, "tests", "golden", "wbfm_chain_ref_executed.npz")
    np.savez_compressed(out, x=x, y=y, rate=np.float64(rate), splits=np.array(splits))
    print("wrote", out, "input", x.shape, "output", y.shape)


if __name__ == "__main__":
    main()
