#!/usr/bin/env python3
"""Golden vectors for the reference's demodulator / resampler COMPOSITES that its own spec vectors do not cover, produced by
the reference itself: each composite (and the pure-Lua process() branches of the blocks inside it) is executed from
/root/reference in this repo's test interpreter on float32-faithful sample cells (tests/lua_reference_env.py) through the
reference's CompositeBlock run loop, fed three ragged vectors.

    python tests/golden/make_composite_goldens.py       # writes tests/golden/composites_ref_executed.npz   (a few minutes)
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
    function Source:instantiate(rate, vectors, data_type)
        self.rate, self.vectors, self.k = rate, vectors, 0
        self:add_type_signature({}, {block.Output("out", data_type)})
    end
    function Source:get_rate() return self.rate end
    function Source:process() self.k = self.k + 1 return self.vectors[self.k] end
    local Sink = block.factory("CollectSink")
    function Sink:instantiate(data_type) self.dt = data_type self:add_type_signature({block.Input("in", data_type)}, {}) end
    function Sink:initialize() self.got = {} end
    function Sink:process(x)
        local copy = self.dt.vector(x.length)
        for i = 0, x.length - 1 do copy.data[i] = x.data[i] end
        self.got[#self.got + 1] = copy
    end
    local make = {
        nbfm = function () return radio.NBFMDemodulator(5e3, 4e3) end,
        am_envelope = function () return radio.AMEnvelopeDemodulator(5e3) end,
        ssb_usb = function () return radio.SSBDemodulator('usb', 3e3) end,
        ssb_lsb = function () return radio.SSBDemodulator('lsb', 3e3) end,
        am_synchronous = function () return radio.AMSynchronousDemodulator(10e3, 5e3) end,
        decimator_c = function () return radio.DecimatorBlock(4, {num_taps = 64}) end,
        interpolator_c = function () return radio.InterpolatorBlock(3, {num_taps = 48}) end,
        resampler_c = function () return radio.RationalResamplerBlock(3, 2, {num_taps = 48}) end,
        interpolator_r = function () return radio.InterpolatorBlock(2, {num_taps = 32}) end,
    }
    return function (name, rate, vectors, complex_in, complex_out)
        local tin = complex_in and types.ComplexFloat32 or types.Float32
        local tout = complex_out and types.ComplexFloat32 or types.Float32
        local sink = Sink(tout)
        local top = radio.CompositeBlock()
        top:connect(Source(rate, vectors, tin), make[name](), sink)
        top:start(false)
        return sink.got
    end
"""

# name -> (rate, complex input, complex output)
CASES = {
    "nbfm": (48000.0, True, False), "am_envelope": (48000.0, True, False), "ssb_usb": (48000.0, True, False), "ssb_lsb": (48000.0, True, False),
    "am_synchronous": (48000.0, True, False), "decimator_c": (48000.0, True, True), "interpolator_c": (48000.0, True, True),
    "resampler_c": (48000.0, True, True), "interpolator_r": (48000.0, False, False),
}


def signal(name, n, rate, rng):
    t = np.arange(n) / rate
    if name == "am_synchronous":                    # AM on a 10 kHz IF: the PLL has a carrier to lock to
        env = 1.0 + 0.5 * np.sin(2 * np.pi * 700 * t)
        return (0.6 * env * np.exp(2j * np.pi * 10010.0 * t + 0.3j) + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    if name == "interpolator_r":
        return (0.5 * np.sin(2 * np.pi * 900 * t) + 0.2 * rng.uniform(-1, 1, n)).astype(np.float32)
    ph = 2 * np.pi * 3e3 * np.cumsum(np.sin(2 * np.pi * 600 * t)) / rate
    env = 1.0 + 0.4 * np.sin(2 * np.pi * 450 * t)
    return (0.7 * env * np.exp(1j * (2 * np.pi * 1500.0 * t + ph)) + 0.05 * (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))).astype(np.complex64)


def main():
    os.environ["LUARADIO_DISABLE_CUDA"] = "1"
    from tests import lua_reference_env as E
    from tests.lua_interp import to_lua
    it, types = E.make_env(lib=None, cuda=False)
    run = it.run(LUA)[0]
    out = {}
    for k, (name, (rate, cin, cout)) in enumerate(CASES.items()):
        n = 1500 if name != "am_synchronous" else 2400
        x = signal(name, n, rate, np.random.default_rng(100 + k))
        splits = [0, n // 3, n // 3 + 1, n]
        parts = [x[a:b] for a, b in zip(splits[:-1], splits[1:])]
        got = it.call(run, [name, rate, to_lua([it.f32.vector_from_numpy(p) for p in parts]), cin, cout])[0]
        y = np.concatenate([it.f32.to_numpy(v) for v in got.array()])
        out[name + "_x"], out[name + "_y"], out[name + "_rate"], out[name + "_splits"] = x, y, np.float64(rate), np.array(splits)
        print(name, x.shape, "->", y.shape, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "composites_ref_executed.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
