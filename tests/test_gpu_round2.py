"""Round-2 GPU parity: the graph rewrites and kernels added after round 1.

  * audio tail  Lowpass(128) -> FMDeemphasis -> Downsampler(5)  rewritten by the noble identity into ONE 133-tap
    decimating real FIR + a pole c^5 at the output rate (graph.cu) -- against the oracle's block-by-block chain;
  * the fused Tuner -> FrequencyDiscriminator stage checked DIRECTLY at the reference's own 1e-6 absolute
    (tests/blocks/signal/frequencydiscriminator_spec.lua epsilon), not after the audio low-pass has averaged it;
  * the real-stream polyphase decimator through the C ABI with ragged calls.

Tolerance unless stated: |got - ref| <= 1e-5 * max(1, ||ref||_inf) (north_star)."""
import ctypes

import numpy as np
import pytest

import luaradio_b200 as radio
from luaradio_b200 import _lib
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def close(got, ref, rel=1e-5, absolute=None):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, "length %s != %s" % (got.shape, ref.shape)
    if ref.size == 0:
        return
    tol = absolute if absolute is not None else rel * max(1.0, float(np.max(np.abs(ref))))
    err = float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128))))
    assert err <= tol, "max abs err %.3g > %.3g" % (err, tol)


def run_graph(blocks, x, rate, chunk, fuse=True):
    src, snk = radio.ArraySource(x, rate, chunk), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, *blocks, snk)
    top.run(False, fuse=fuse)
    return snk.result(), top


def audio_tail_blocks():
    return [radio.LowpassFilterBlock(128, 15e3), radio.FMDeemphasisFilterBlock(75e-6), radio.DownsamplerBlock(5)]


def audio_tail_oracle(rate):
    b, a = O.fm_deemphasis_taps(75e-6, rate)
    return O.Chain(O.lowpass_filter(128, 15e3, rate, False), O.IIRFilterFast(b, a, False), O.Downsampler(5))


@pytest.mark.parametrize("chunk", [1 << 22, 100003, 4099, 7])
def test_audio_tail_noble_identity(chunk):
    """examples/rtlsdr_wbfm_mono.lua:15-18 on a real stream at 220.5 kHz: three blocks -> two stages, and the stream is
    the same whatever the vector size (the decimation phase, the 132-sample history and the pole state are carried)."""
    rate = 220500.0
    n = 400000 if chunk > 100 else 30000
    rng = np.random.default_rng(11)
    t = np.arange(n) / rate
    x = (0.4 * np.sin(2 * np.pi * 1000 * t) + 0.3 * np.sin(2 * np.pi * 9000 * t) + 0.2 * rng.uniform(-1, 1, n)).astype(np.float32)
    got, top = run_graph(audio_tail_blocks(), x, rate, chunk)
    desc = top.describe_gpu_graph()
    assert desc == "fir*iir1_rrrf(133,/5)+pole[fused x3]", desc
    ref = audio_tail_oracle(rate).process(x)
    close(got, ref)
    # chunked == whole to float32 rounding (the streaming state is exact, only summation grouping at tile edges differs)
    whole, _ = run_graph(audio_tail_blocks(), x, rate, 1 << 24)
    close(got, whole, absolute=2e-6)
    # and the unfused graph (three kernels, the reference's dataflow) agrees too
    unf, top2 = run_graph(audio_tail_blocks(), x, rate, chunk if chunk > 100 else 1 << 20, fuse=False)
    assert top2.describe_gpu_graph().count("|") == 2
    close(unf, ref)


def test_audio_tail_other_shapes_keep_the_old_fusion():
    """A shape the real polyphase kernel does not have (64 taps, /4) must still be correct: FIR | IIR+down."""
    rate = 48000.0
    rng = np.random.default_rng(12)
    x = rng.uniform(-1, 1, 200000).astype(np.float32)
    blocks = [radio.LowpassFilterBlock(64, 5e3), radio.SinglepoleLowpassFilterBlock(2e3), radio.DownsamplerBlock(4)]
    got, top = run_graph(blocks, x, rate, 50000)
    b, a = O.singlepole_lowpass_taps(2e3, rate)
    ref = O.Chain(O.lowpass_filter(64, 5e3, rate, False), O.IIRFilterFast(b, a, False), O.Downsampler(4)).process(x)
    close(got, ref)


@pytest.mark.parametrize("chunk", [1 << 22, 65536, 12345])
def test_tuner_discriminator_direct_1e6(chunk):
    """Tuner(-250e3, 200e3, 5) -> FrequencyDiscriminator(1.25) -> sink: the fused tuner+discriminator stage (packed
    atan2, tile phasor skipped) at the reference spec's 1e-6 absolute tolerance on a strong FM signal."""
    n = 500000
    x = O.synth_fm_iq(0, n)
    got, top = run_graph([radio.TunerBlock(-250e3, 200e3, 5), radio.FrequencyDiscriminatorBlock(1.25)], x, 1102500.0, chunk)
    assert top.describe_gpu_graph().startswith("tuner+discrim(128,/5)"), top.describe_gpu_graph()
    ref = O.Chain(O.tuner(-250e3, 200e3, 5, 1102500.0), O.FrequencyDiscriminator(1.25)).process(x)
    assert got.shape == ref.shape
    # the first 26 outputs come from the filter's start-up transient (|y| ~ 1e-4: the angle of a near-zero phasor is
    # ill-conditioned -- the reference's own vectors avoid it by construction); from there on: 1e-6 absolute
    err = np.abs(got[26:] - ref[26:])
    assert float(err.max()) <= 1e-6, "max abs err %.3g" % float(err.max())
    close(got[:26], ref[:26], absolute=2e-5)


# Added with the stream-start fix of the fused discriminator (tuner.cu, edge-tile kernel), after the round's GPU budget had
# reached zero: non-strict xfail until it has run on a GPU once.
@pytest.mark.xfail(strict=False, reason="stream-start signed-zero fix in the fused discriminator has not run on a GPU yet")
@pytest.mark.parametrize("quadrant", [(1, 1), (-1, 1), (1, -1), (-1, -1)])
@pytest.mark.parametrize("fuse", [True, False])
def test_discriminator_first_sample_quadrants(quadrant, fuse):
    """frequencydiscriminator.lua:33,72: the stream's first output is atan2f of y[0] * conj(0), a product of signed zeros --
    pi / gain when both parts of the tuner's first output are negative, 0 in the other three quadrants (pinned on the
    reference's own code: tests/test_lua_reference.py::test_discriminator_first_sample_signed_zero_artefact)."""
    rate, n = 1102500.0, 60000
    x = O.synth_fm_iq(0, n).copy()
    h0 = float(O.f32_taps(O.firwin_lowpass(128, 100e3 / (rate / 2)))[0])
    sr, si = quadrant
    x[0] = np.complex64(complex(sr * 0.3, si * 0.2) * (1.0 if h0 > 0 else -1.0))      # y[0] = h[0] x[0]: the translator's phase at n = 0 is 0
    got, top = run_graph([radio.TunerBlock(-250e3, 200e3, 5), radio.FrequencyDiscriminatorBlock(1.25)], x, rate, 1 << 22, fuse=fuse)
    ref = O.Chain(O.tuner(-250e3, 200e3, 5, rate), O.FrequencyDiscriminator(1.25)).process(x)
    first = np.pi / (2 * np.pi * 1.25) if quadrant == (-1, -1) else 0.0
    assert abs(float(ref[0]) - first) <= 1e-6
    assert abs(float(got[0]) - first) <= 1e-6, "first output %.6g, reference %.6g" % (float(got[0]), first)
    assert float(np.abs(got[26:] - ref[26:]).max()) <= 1e-6


@pytest.mark.parametrize("M,D", [(133, 5), (131, 5), (135, 5)])
def test_real_polyphase_decimator_c_abi(M, D):
    """lrb200_fir_create_rrrf(taps, M, decim=5): the real-stream polyphase kernel through HOST-mode calls of ragged
    length (one sample up to several tiles), against FIR -> Downsampler of the oracle."""
    rng = np.random.default_rng(M)
    n = 300000
    taps = O.f32_taps(O.firwin_lowpass(M, 1.0 / D))
    x = rng.uniform(-1, 1, n).astype(np.float32)
    lib = _lib.require_device()
    h = _lib.check_handle(lib.lrb200_fir_create_rrrf(taps.ctypes.data, M, D, _lib.LRB200_HOST), "fir")
    assert lib.lrb200_fir_get_algorithm(h) == _lib.FIR_DIRECT
    outs, i = [], 0
    sizes = [1, 2, 3, 4, 5, 17, 4096, 20481, 100000, 1, 7, 50000]
    k = 0
    while i < n:
        m = min(n - i, sizes[k % len(sizes)])
        k += 1
        seg = np.ascontiguousarray(x[i:i + m])
        out = np.zeros(lib.lrb200_block_max_output(h, m), np.float32)
        no = ctypes.c_size_t()
        _lib.check(lib.lrb200_fir_execute(h, seg.ctypes.data, m, out.ctypes.data, ctypes.byref(no)))
        outs.append(out[:no.value])
        i += m
    lib.lrb200_fir_destroy(h)
    ref = O.Chain(O.FIRFilter(taps, False), O.Downsampler(D)).process(x)
    close(np.concatenate(outs), ref)


def _chain_graph(lib):
    import bench
    return bench.build_chain_graph(lib, _lib)


def test_superchunk_mode_equals_synchronous_calls():
    """lrb200_graph_set_superchunk: the reference's 8192-sample vectors (zero.lua:30) packed into pinned super-chunks and
    processed asynchronously give the same stream as one synchronous call per vector, after lrb200_graph_flush."""
    lib = _lib.require_device()
    n, vec = 1000000, 8192
    x = O.synth_fm_iq(0, n)
    g = _chain_graph(lib)

    def run(sc, flush_every=None):
        _lib.check(lib.lrb200_graph_reset(g))
        _lib.check(lib.lrb200_graph_set_superchunk(g, sc))
        cap = lib.lrb200_graph_max_output(g, vec) + 16
        y = np.zeros(cap, np.float32)
        no = ctypes.c_size_t()
        outs, counts = [], []
        for o in range(0, n, vec):
            seg = np.ascontiguousarray(x[o:o + vec])
            _lib.check(lib.lrb200_graph_execute(g, seg.ctypes.data, len(seg), y.ctypes.data, ctypes.byref(no)))
            assert no.value <= cap
            counts.append(no.value)
            outs.append(y[:no.value].copy())
        _lib.check(lib.lrb200_graph_flush(g, y.ctypes.data, ctypes.byref(no)))
        outs.append(y[:no.value].copy())
        _lib.check(lib.lrb200_graph_set_superchunk(g, 0))
        return np.concatenate(outs), counts

    sync, _ = run(0)
    ref = O.wbfm_mono_chain().process(x)
    close(sync, ref)
    for sc in (1 << 16, 100000, 1 << 20):
        acc, counts = run(sc)
        assert acc.shape == sync.shape
        close(acc, sync, absolute=2e-6)
        assert counts.count(0) > len(counts) // 2         # most calls only append to the pinned slot
    lib.lrb200_graph_destroy(g)


def test_graph_halo_and_execute_shard_on_one_device():
    """lrb200_graph_halo / lrb200_graph_execute_shard: the chunks of a 4-way time-sharded stream, each run cold with only a
    head piece depending on the left neighbour's samples, concatenate to the single-stream result."""
    lib = _lib.require_device()
    total, world = 2000000, 4
    x = O.synth_fm_iq(0, total)
    g, gh = _chain_graph(lib), _chain_graph(lib)
    halo = lib.lrb200_graph_halo(g)
    assert halo % 25 == 0 and 2500 <= halo <= 4000, halo
    whole = np.zeros(total // 25 + 8, np.float32)
    no = ctypes.c_size_t()
    _lib.check(lib.lrb200_graph_execute(g, x.ctypes.data, total, whole.ctypes.data, ctypes.byref(no)))
    whole = whole[:no.value]
    per = total // world
    assert per % 25 == 0
    d_in = lib.lrb200_malloc((per + halo) * 8)
    d_out = lib.lrb200_malloc((per // 25 + 8) * 4)
    parts = []
    for r in range(world):
        start = r * per
        lead = halo if r > 0 else 0
        seg = np.ascontiguousarray(x[start - lead:start + per])
        _lib.check(lib.lrb200_memcpy_h2d(ctypes.c_void_p(d_in + (halo - lead) * 8), seg.ctypes.data, seg.nbytes))
        _lib.check(lib.lrb200_graph_execute_shard(g, gh, ctypes.c_void_p(d_in), halo, per, start, ctypes.c_void_p(d_out),
                                                  ctypes.byref(no), None))
        out = np.zeros(no.value, np.float32)
        _lib.check(lib.lrb200_memcpy_d2h(out.ctypes.data, ctypes.c_void_p(d_out), out.nbytes))
        _lib.check(lib.lrb200_sync())
        parts.append(out)
    got = np.concatenate(parts)
    assert got.shape == whole.shape
    close(got, whole, absolute=2e-6)
    lib.lrb200_free(d_in)
    lib.lrb200_free(d_out)
    lib.lrb200_graph_destroy(g)
    lib.lrb200_graph_destroy(gh)
