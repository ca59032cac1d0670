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


def run_c_chain2(lib, x, n0, threads, use_fft, halo=4000):
    t1, t2, b, a, omega = chain_params()
    out = np.zeros(len(x) // 25 + 8, np.float32)
    n = lib.lro_wbfm_chain2(x.ctypes.data, len(x), n0, t1.ctypes.data, 128, 5, omega, 2 * math.pi * 1.25,
                            t2.ctypes.data, 128, b.ctypes.data, a.ctypes.data, 5, out.ctypes.data, threads, halo, use_fft)
    return out[:n]


def test_c_overlap_save_fir_matches_numpy_firfilterfft():
    """The C overlap-save path (firfilter.lua:320-398 with the in-house FFT) against the numpy FIRFilterFFT, call by call:
    same block sizes N = 2^floor(log2 8M), L = N - M + 1, same number of samples emitted per call."""
    lib = cbuild.load()
    rng = np.random.default_rng(5)
    for M in (17, 33, 128):
        taps = O.f32_taps(O.firwin_lowpass(M, 0.2))
        n = 40000
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        f = lib.lro_firfft_new(taps.ctypes.data, M, 0, 0)
        ref = O.FIRFilterFFT(taps, True)
        i = 0
        for k in (1, 5, 8192, 3000, 20000, 7, 8795):
            seg = np.ascontiguousarray(x[i:i + k])
            i += k
            y = np.zeros(len(seg) + 8 * M + 16, np.complex64)
            no = lib.lro_firfft_process_c(f, seg.ctypes.data, len(seg), y.ctypes.data)
            r = ref.process(seg)
            assert no == len(r), (M, k, no, len(r))
            if no:
                assert np.max(np.abs(y[:no] - r)) < 2e-6
        lib.lro_firfft_free(f)
        # real input (two blocks per transform) + flush == the dot-product filter over the whole stream
        xr = rng.uniform(-1, 1, n).astype(np.float32)
        f2 = lib.lro_firfft_new(taps.ctypes.data, M, 0, 1)
        y = np.zeros(n + 16 * M + 16, np.float32)
        no = 0
        for a, b in ((0, 12345), (12345, 12346), (12346, n)):
            seg = np.ascontiguousarray(xr[a:b])
            no += lib.lro_firfft_process_r(f2, seg.ctypes.data, len(seg), y[no:].ctypes.data)
        no += lib.lro_firfft_flush_r(f2, y[no:].ctypes.data)
        assert no == n
        assert np.max(np.abs(y[:n] - O.FIRFilter(taps, False).process(xr))) < 2e-6
        lib.lro_firfft_free(f2)


def test_c_chain_fft_path_matches_numpy_oracle():
    lib = cbuild.load()
    n = 300000
    x = O.synth_fm_iq(0, n)
    ref = O.wbfm_mono_chain().process(x)
    for threads in (1, 4):
        got = run_c_chain2(lib, x, 0, threads, 1)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) < 2e-5, (threads, np.max(np.abs(got - ref)))
    # the FIR alone, chunk-parallel, both algorithms agree
    taps = O.f32_taps(O.firwin_lowpass(128, 15e3 / 110250.0))
    y = np.zeros(n, np.complex64)
    lib.lro_fir_crcf_fft_parallel(x.ctypes.data, n, taps.ctypes.data, 128, y.ctypes.data, 4)
    assert np.max(np.abs(y - O.FIRFilter(taps, True).process(x))) < 1e-5


def test_c_synth_and_core_count():
    lib = cbuild.load()
    xs = np.zeros(50000, np.complex64)
    for n0 in (0, (1 << 33) + 12345):
        lib.lro_synth_fm_iq(xs.ctypes.data, n0, len(xs), 1, 1102500.0, 250e3, 75e3, 0.5, 0.01, 3)
        assert np.max(np.abs(xs - O.synth_fm_iq(n0, len(xs)))) < 2e-6
    assert 1 <= lib.lro_usable_cpus() <= lib.lro_max_threads()
