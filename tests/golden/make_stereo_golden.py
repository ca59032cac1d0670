#!/usr/bin/env python3
"""Golden vector for the reference's WBFMStereoDemodulator composite (composites/wbfmstereodemodulator.lua:22-64), produced by
the reference itself: the stock composite, the pure-Lua process() branches of its 14 blocks (discriminator, Hilbert, delay,
complex band-pass, PLL, multiply-conjugate, low-pass x2, complex-to-real x2, add, subtract, de-emphasis x2) and the
CompositeBlock run loop are executed from /root/reference in this repo's test interpreter on float32-faithful sample cells
(tests/lua_reference_env.py), fed three ragged vectors of an FM stereo multiplex (pilot, L+R, L-R on 38 kHz).  About 90 s.

    python tests/golden/make_stereo_golden.py           # writes tests/golden/wbfm_stereo_ref_executed.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

LUA = """
    local radio = require('radio')
    local block = require('radio.core.block')
    local types = require('radio.types')
    local Source = block.factory("ArraySource")
    function Source:instantiate(rate, vectors)
        self.rate, self.vectors, self.k = rate, vectors, 0
        self:add_type_signature({}, {block.Output("out", types.ComplexFloat32)})
    end
    function Source:get_rate() return self.rate end
    function Source:process() self.k = self.k + 1 return self.vectors[self.k] end
    local Sink = block.factory("CollectSink")
    function Sink:instantiate() self:add_type_signature({block.Input("in", types.Float32)}, {}) end
    function Sink:initialize() self.got = {} end
    function Sink:process(x)
        local copy = types.Float32.vector(x.length)
        for i = 0, x.length - 1 do copy.data[i] = x.data[i] end
        self.got[#self.got + 1] = copy
    end
    return function (vectors)
        local l, r = Sink(), Sink()
        local top = radio.CompositeBlock()
        local demod = radio.WBFMStereoDemodulator()
        top:connect(Source(220500, vectors), demod)
        top:connect(demod, 'left', l, 'in')
        top:connect(demod, 'right', r, 'in')
        top:start(false)
        return l.got, r.got
    end
"""


def stereo_mpx(n, rate, rng):
    """An FM-modulated stereo multiplex: L+R, 19 kHz pilot, (L-R) on 38 kHz DSB-SC."""
    t = np.arange(n) / rate
    left, right = 0.5 * np.sin(2 * np.pi * 700 * t), 0.4 * np.sin(2 * np.pi * 2300 * t)
    mpx = 0.45 * (left + right) + 0.1 * np.sin(2 * np.pi * 19e3 * t) + 0.45 * (left - right) * np.sin(2 * np.pi * 38e3 * t)
    phase = 2 * np.pi * 75e3 * np.cumsum(mpx) / rate
    noise = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    return (np.exp(1j * phase) + 0.001 * noise).astype(np.complex64)


def main():
    os.environ["LUARADIO_DISABLE_CUDA"] = "1"
    from tests import lua_reference_env as E
    from tests.lua_interp import to_lua
    it, types = E.make_env(lib=None, cuda=False)
    rate, n = 220500.0, 4500
    x = stereo_mpx(n, rate, np.random.default_rng(5))
    run = it.run(LUA)[0]
    splits = [0, n // 3, n // 3 + 1, n]
    parts = [x[a:b] for a, b in zip(splits[:-1], splits[1:])]
    left, right = it.call(run, [to_lua([it.f32.vector_from_numpy(p) for p in parts])])
    y_l = np.concatenate([it.f32.to_numpy(v) for v in left.array()])
    y_r = np.concatenate([it.f32.to_numpy(v) for v in right.array()])
    out = os.path.join(ROOT, "tests", "golden", "wbfm_stereo_ref_executed.npz")
    np.savez_compressed(out, x=x, left=y_l, right=y_r, rate=np.float64(rate), splits=np.array(splits))
    print("wrote", out, x.shape, y_l.shape, y_r.shape)


if __name__ == "__main__":
    main()
