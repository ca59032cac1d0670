"""Parity at BASELINE.json's FULL sizes (configs[1..3]), where the oracle cannot run the whole stream:

  * windowed oracle: every path here has finite memory (FIR history, a single-pole IIR whose 4000-sample tail is
    < 1e-100), and the synthetic sources are counter-based (any window of the global stream can be regenerated on
    the CPU), so the oracle run COLD from HALO samples before a window must equal the GPU's full-stream output
    inside that window -- checked at the start, deep inside (sample index > 2^27) and at the very end;
  * size-independent properties: output length, linearity f(x1 + a x2) = f(x1) + a f(x2), DC gain = sum(taps),
    chunked == whole (streaming state carried across calls).

Inputs are generated on the device (lrb200_synth_*), device buffers through lrb200_graph_execute_device /
lrb200_block_execute with DEVICE-mode blocks; only the compared windows cross PCIe.
Tolerance: 1e-5 * max(1, ||ref||_inf) (north_star)."""
import ctypes
import math

import numpy as np
import pytest

from luaradio_b200 import _lib
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu

HALO = 4000
D = _lib.LRB200_DEVICE


def dev_alloc(lib, nbytes):
    p = lib.lrb200_malloc(nbytes)
    assert p, _lib.last_error()
    return p


def fetch(lib, dptr, offset_bytes, count, dtype):
    out = np.empty(count, dtype)
    _lib.check(lib.lrb200_memcpy_d2h(out.ctypes.data, ctypes.c_void_p(dptr + offset_bytes), out.nbytes), "d2h")
    _lib.check(lib.lrb200_sync(), "sync")
    return out


def close(got, ref, what, rel=1e-5):
    assert got.shape == ref.shape, what
    scale = max(1.0, float(np.max(np.abs(ref))))
    err = float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128))))
    assert err <= rel * scale, "%s: max abs err %.3g > %.3g" % (what, err, rel * scale)


def windows(n, width, align):
    """(start, count) windows at the head, past 2^27, and at the tail of an n-sample stream."""
    w = (width // align) * align
    mids = [s for s in ((1 << 27) + 12345, n // 2 + 777) if s + w < n]
    starts = [0] + [(s // align) * align for s in mids] + [((n - w) // align) * align]
    return [(s, min(w, n - s)) for s in sorted(set(starts))]


def test_wbfm_chain_256Mi_windows_and_chunking():
    """configs[2]: the full rtlsdr_wbfm_mono flow graph on 256 Mi samples of synthetic FM IQ at 1.1025 MS/s."""
    import bench
    lib = _lib.require_device()
    n = 268435450
    x = dev_alloc(lib, n * 8)
    _lib.check(lib.lrb200_synth_fm_iq(ctypes.c_void_p(x), 0, n, 1, bench.RATE, 250e3, 75e3, 0.5, 0.01), "synth")
    g = bench.build_chain_graph(lib, _lib)
    n_max = lib.lrb200_graph_max_output(g, n)
    y = dev_alloc(lib, (n_max + 16) * 4)
    no = ctypes.c_size_t(0)
    _lib.check(lib.lrb200_graph_execute_device(g, ctypes.c_void_p(x), n, ctypes.c_void_p(y), ctypes.byref(no)), "execute")
    _lib.check(lib.lrb200_sync(), "sync")
    assert no.value == n // 25
    for s, cnt in windows(n, 200000, 25):
        lead = HALO if s > 0 else 0
        ch = O.wbfm_mono_chain()
        ch.blocks[0].blocks[0].n0 = s - lead                      # translator phase of the global index
        ref = ch.process(O.synth_fm_iq(s - lead, cnt + lead))[lead // 25:]
        got = fetch(lib, y, (s // 25) * 4, len(ref), np.float32)
        assert np.all(np.isfinite(got))
        close(got, ref, "window @%d" % s)
    # chunked (three ragged calls, state carried) == whole, compared on the device output
    y2 = dev_alloc(lib, (n_max + 16) * 4)
    _lib.check(lib.lrb200_graph_reset(g), "reset")
    cuts, done = [0, 100000007, 100000007 + 33333331, n], 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        _lib.check(lib.lrb200_graph_execute_device(g, ctypes.c_void_p(x + a * 8), b - a, ctypes.c_void_p(y2 + done * 4),
                                                   ctypes.byref(no)), "execute")
        done += no.value
    _lib.check(lib.lrb200_sync(), "sync")
    assert done == n // 25
    for s, cnt in windows(n, 2000000, 25):
        a = fetch(lib, y, (s // 25) * 4, cnt // 25, np.float32)
        b = fetch(lib, y2, (s // 25) * 4, cnt // 25, np.float32)
        assert float(np.max(np.abs(a - b))) <= 2e-6, "chunked vs whole @%d" % s
    lib.lrb200_graph_destroy(g)
    for p in (x, y, y2):
        lib.lrb200_free(p)


def test_translate_fir_decimate_64Mi_windows_and_linearity():
    """configs[1]: FrequencyTranslator + 128-tap complex FIR + decimate-by-5 on 64 Mi samples of white IQ."""
    lib = _lib.require_device()
    n, rate, offset, Dn = 1 << 26, 2.0e6, -310e3, 5
    taps = O.f32_taps(O.firwin_complex_bandpass(128, [0.05, 0.3])).astype(np.complex64)

    def graph():
        g = _lib.check_handle(lib.lrb200_graph_create(), "graph")
        for h in (lib.lrb200_rotator_create(offset / rate, D), lib.lrb200_fir_create_cccf(taps.ctypes.data, 128, 1, D),
                  lib.lrb200_downsample_create(Dn, 8, D)):
            _lib.check(lib.lrb200_graph_append(g, _lib.check_handle(h, "block")), "append")
        _lib.check(lib.lrb200_graph_commit(g, 1), "commit")
        return g

    def run(g, dx, dy):
        no = ctypes.c_size_t(0)
        _lib.check(lib.lrb200_graph_reset(g), "reset")
        _lib.check(lib.lrb200_graph_execute_device(g, ctypes.c_void_p(dx), n, ctypes.c_void_p(dy), ctypes.byref(no)), "execute")
        _lib.check(lib.lrb200_sync(), "sync")
        return no.value

    g = graph()
    n_max = lib.lrb200_graph_max_output(g, n)
    x1, x2 = dev_alloc(lib, n * 8), dev_alloc(lib, n * 8)
    y1, y2, y3 = (dev_alloc(lib, (n_max + 16) * 8) for _ in range(3))
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x1), 0, n, 1), "synth")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x2), 0, n, 2), "synth")
    n_out = run(g, x1, y1)
    assert n_out == (n + Dn - 1) // Dn
    for s, cnt in windows(n, 100000, Dn):
        lead = HALO if s > 0 else 0
        tr = O.FrequencyTranslator(offset, rate)
        tr.n0 = s - lead
        ref = O.Chain(tr, O.FIRFilter(taps, True), O.Downsampler(Dn)).process(O.synth_white_iq(s - lead, cnt + lead, 1))[lead // Dn:]
        close(fetch(lib, y1, (s // Dn) * 8, len(ref), np.complex64), ref, "window @%d" % s)
    # linearity over the whole stream, evaluated on the device with torch (plumbing): f(x1 + a x2) = f(x1) + a f(x2)
    import torch

    def view(ptr, count):
        class _A:            # __cuda_array_interface__ view of a C-ABI device buffer
            __cuda_array_interface__ = {"shape": (count * 2,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_A(), device="cuda")

    a = 0.37
    run(g, x2, y2)
    vx1, vx2 = view(x1, n), view(x2, n)
    vx1.add_(vx2, alpha=a)                                      # x1 <- x1 + a x2
    torch.cuda.synchronize()
    run(g, x1, y3)
    vy1, vy2, vy3 = view(y1, n_out), view(y2, n_out), view(y3, n_out)
    err = float((vy3 - (vy1 + a * vy2)).abs().max())
    scale = max(1.0, float(vy3.abs().max()))
    assert err <= 1e-5 * scale, "linearity: %.3g" % err
    lib.lrb200_graph_destroy(g)
    for p in (x1, x2, y1, y2, y3):
        lib.lrb200_free(p)


@pytest.mark.parametrize("M", [16, 128, 2048])
def test_fir_1Gi_windows_and_dc_gain(M):
    """configs[3] (tap sweep points): LowpassFilterBlock(M) taps over 1 Gi complex samples."""
    lib = _lib.require_device()
    n = 1 << 30
    taps = O.f32_taps(O.firwin_lowpass(M, 0.2)).astype(np.float32)
    x, y = dev_alloc(lib, n * 8), dev_alloc(lib, n * 8)
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x), 0, n, 3), "synth")
    f = _lib.check_handle(lib.lrb200_fir_create_crcf(taps.ctypes.data, M, 1, D), "fir")
    no = ctypes.c_size_t(0)
    _lib.check(lib.lrb200_block_execute(f, ctypes.c_void_p(x), n, ctypes.c_void_p(y), ctypes.byref(no)), "execute")
    _lib.check(lib.lrb200_sync(), "sync")
    assert no.value == n
    for s, cnt in windows(n, 60000, 1):
        lead = HALO if s > 0 else 0
        ref = O.FIRFilter(taps, True).process(O.synth_white_iq(s - lead, cnt + lead, 3))[lead:]
        close(fetch(lib, y, s * 8, len(ref), np.complex64), ref, "M=%d window @%d" % (M, s))
    # DC gain: a constant stream comes out as const * sum(taps) once the history is full
    import torch

    class _A:
        __cuda_array_interface__ = {"shape": (n * 2,), "typestr": "<f4", "data": (x, False), "version": 2}
    vx = torch.as_tensor(_A(), device="cuda").view(-1, 2)
    vx[:, 0] = 0.75
    vx[:, 1] = -0.5
    torch.cuda.synchronize()
    _lib.check(lib.lrb200_block_reset(f), "reset")
    _lib.check(lib.lrb200_block_execute(f, ctypes.c_void_p(x), n, ctypes.c_void_p(y), ctypes.byref(no)), "execute")
    _lib.check(lib.lrb200_sync(), "sync")
    want = (0.75 - 0.5j) * float(np.sum(taps.astype(np.float64)))
    for s, cnt in windows(n, 60000, 1):
        got = fetch(lib, y, max(s, M) * 8, cnt - M, np.complex64)
        assert float(np.max(np.abs(got - want))) <= 1e-5, "DC gain M=%d @%d" % (M, s)
    lib.lrb200_block_destroy(f)
    lib.lrb200_free(x)
    lib.lrb200_free(y)
