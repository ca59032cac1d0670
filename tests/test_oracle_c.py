"""Pin the C restatement (oracle/lr_oracle.c, the CPU-baseline port) against the numpy oracle."""
import math

import numpy as np

from oracle import cbuild
from oracle import lr_oracle as O


def chain_params(rate=1102500.0):
    t1 = O.f32_taps(O.firwin_lowpass(128, (200e3 / 2) / (rate / 2)))
    r1 = rate / 5
    t2 = O.f32_taps(O.firwin_lowpass(128, 15e3 / (r1 / 2)))
    b, a = O.fm_deemphasis_taps(75e-6, r1)
    omega = 2 * math.pi * (-250e3 / rate)
    return t1, t2, b, a, omega


def run_c_chain(lib, x, n0, threads, halo=25 * 1024):
    t1, t2, b, a, omega = chain_params()
    out = np.zeros(len(x) // 25 + 8, np.float32)
    n = lib.lro_wbfm_chain(x.ctypes.data, len(x), n0, t1.ctypes.data, 128, 5, omega, 2 * math.pi * 1.25,
                           t2.ctypes.data, 128, b.ctypes.data, a.ctypes.data, 5, out.ctypes.data, threads, halo)
    return out[:n]


def test_c_chain_matches_numpy_oracle():
    lib = cbuild.load()
    n = 300000
    x = O.synth_fm_iq(0, n)
    ref = O.wbfm_mono_chain().process(x)
    for threads in (1, 4):
        got = run_c_chain(lib, x, 0, threads)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 2e-5, (threads, np.max(np.abs(got - ref)))


def test_c_fir_matches_numpy_oracle():
    lib = cbuild.load()
    rng = np.random.default_rng(2)
    n = 100000
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    taps = O.f32_taps(O.firwin_lowpass(128, 15e3 / 110250.0))
    y = np.zeros(n, np.complex64)
    lib.lro_fir_crcf_parallel(x.ctypes.data, n, taps.ctypes.data, 128, y.ctypes.data, 4)
    ref = O.FIRFilter(taps, True).process(x)
    assert np.max(np.abs(y - ref)) < 1e-5
