"""Seeded random sweep over shapes the hand-picked cases do not enumerate: tap counts, decimations, type combinations,
algorithm selection, stream lengths and chunkings drawn at random (fixed seeds: reproducible), each compared with the
oracle at the north_star tolerance.  Covers the kernel-selection boundaries (register-tiled direct / catch-all direct /
overlap-save / partitioned overlap-save, interior / edge kernels, short-call fallbacks) from the outside."""
import ctypes

import numpy as np
import pytest

import luaradio_b200 as radio
from luaradio_b200 import _lib
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def close(got, ref, what, rel=1e-5):
    assert got.shape == ref.shape, "%s: length %s != %s" % (what, got.shape, ref.shape)
    if ref.size == 0:
        return
    scale = max(1.0, float(np.max(np.abs(ref))))
    err = float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128))))
    assert err <= rel * scale, "%s: max abs err %.3g > %.3g" % (what, err, rel * scale)


def chunks(rng, n):
    style = rng.integers(0, 3)
    hi = [5000, 70000, n + 1][style]
    cuts, i = [], 0
    while i < n:
        k = int(rng.integers(0, hi))
        cuts.append((i, min(n, i + k)))
        i += k
    return cuts or [(0, 0)]


def rnd(rng, n, cplx):
    if cplx:
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    return rng.uniform(-1, 1, n).astype(np.float32)


@pytest.mark.parametrize("seed", range(36))
def test_random_fir_shapes(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = ["crcf", "cccf", "rrrf"][seed % 3]
    M = int(rng.choice([int(rng.integers(1, 40)), int(rng.integers(40, 520)), int(rng.integers(520, 3000))], p=[0.35, 0.45, 0.2]))
    D = int(rng.choice([1, int(rng.integers(2, 13))]))
    algo = int(rng.choice([_lib.FIR_AUTO, _lib.FIR_DIRECT, _lib.FIR_FFT]))
    n = int(rng.choice([int(rng.integers(0, 3000)), int(rng.integers(3000, 400000))]))
    taps = rng.uniform(-1, 1, M)
    if kind == "cccf":
        taps = taps + 1j * rng.uniform(-1, 1, M)
    taps = (taps / max(1e-3, np.sum(np.abs(taps)))).astype(np.complex64 if kind == "cccf" else np.float32)
    x = rnd(rng, n, kind != "rrrf")
    lib = _lib.require_device()
    create = {"crcf": lib.lrb200_fir_create_crcf, "cccf": lib.lrb200_fir_create_cccf, "rrrf": lib.lrb200_fir_create_rrrf}[kind]
    h = _lib.check_handle(create(taps.ctypes.data, M, D, _lib.LRB200_HOST), "fir")
    _lib.check(lib.lrb200_fir_set_algorithm(h, algo))
    outs = []
    for a, b in chunks(rng, n):
        seg = np.ascontiguousarray(x[a:b])
        out = np.zeros(lib.lrb200_block_max_output(h, len(seg)) + 1, x.dtype)
        no = ctypes.c_size_t()
        _lib.check(lib.lrb200_fir_execute(h, seg.ctypes.data, len(seg), out.ctypes.data, ctypes.byref(no)))
        outs.append(out[:no.value])
    lib.lrb200_fir_destroy(h)
    ref = O.Chain(O.FIRFilter(taps, kind != "rrrf"), O.Downsampler(D)).process(x)
    close(np.concatenate(outs), ref, "kind=%s M=%d D=%d algo=%d n=%d" % (kind, M, D, algo, n))


@pytest.mark.parametrize("seed", range(12))
def test_random_tuner_and_decimator_graphs(seed):
    """TunerBlock / DecimatorBlock with random offsets, bandwidths, decimations and tap counts through the flow graph
    (fused and unfused): exercises every fusion rule's shape checks and fallbacks."""
    rng = np.random.default_rng(2000 + seed)
    rate = float(rng.choice([48000.0, 1102500.0, 2.4e6]))
    D = int(rng.integers(2, 11))
    M = int(rng.choice([128, int(rng.integers(8, 300))]))
    n = int(rng.integers(1000, 300000))
    x = rnd(rng, n, True)
    fuse = bool(seed % 2)
    chunk = int(rng.choice([1 << 22, int(rng.integers(100, 50000))]))
    if seed % 3:
        offset, bw = float(rng.uniform(-0.4, 0.4) * rate), float(rng.uniform(0.05, 0.4) * rate / D)
        blk = radio.TunerBlock(offset, bw, D, {"num_taps": M})
        ref = O.tuner(offset, bw, D, rate, M).process(x)
        what = "Tuner(%g, %g, %d, taps %d) rate %g" % (offset, bw, D, M, rate)
    else:
        blk = radio.DecimatorBlock(D, {"num_taps": M})
        ref = O.decimator(D, True, M).process(x)
        what = "Decimator(%d, taps %d)" % (D, M)
    src, snk = radio.ArraySource(x, rate, chunk), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, blk, snk)
    top.run(False, fuse=fuse)
    close(snk.result(), ref, "%s fuse=%s chunk=%d n=%d graph=%s" % (what, fuse, chunk, n, top.describe_gpu_graph()))


@pytest.mark.parametrize("seed", range(8))
def test_random_single_pole_iir(seed):
    rng = np.random.default_rng(3000 + seed)
    rate = float(rng.choice([44100.0, 220500.0, 2e6]))
    cplx = bool(seed % 2)
    n = int(rng.integers(1, 500000))
    x = rnd(rng, n, cplx)
    which = seed % 4
    if which == 0:
        tau = float(rng.uniform(10e-6, 200e-6))
        blk, (b, a) = radio.FMDeemphasisFilterBlock(tau), O.fm_deemphasis_taps(tau, rate)
    elif which in (1, 2):
        fc = float(rng.uniform(0.002, 0.3) * rate)
        blk, (b, a) = radio.SinglepoleLowpassFilterBlock(fc), O.singlepole_lowpass_taps(fc, rate)
    else:
        fc = float(rng.uniform(0.002, 0.3) * rate)
        blk, (b, a) = radio.SinglepoleHighpassFilterBlock(fc), O.singlepole_highpass_taps(fc, rate)
    src, snk = radio.ArraySource(x, rate, int(rng.integers(1000, 200000))), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, blk, snk)
    top.run(False)
    ref = O.IIRFilterFast(b, a, cplx).process(x)
    close(snk.result(), ref, "iir which=%d rate=%g n=%d" % (which, rate, n))
