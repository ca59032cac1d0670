"""GPU parity for the blocks and the scheduler added around the hot path:

  * two-input blocks (Multiply, MultiplyConjugate, Add, Subtract) and DelayBlock against the reference's golden vectors
    (whole and split calls) and on long streams against the oracle;
  * spectrum_utils.PSD against the reference's committed PSD vectors at the reference's tolerances;
  * PLLBlock against the oracle (the reference has no pll_spec; the oracle is pinned on the reference's pll.lua executed in the test interpreter);
  * the DAG scheduler: the reference's tests/top_spec.lua:14-55 topology (two sources -> MultiplyConjugate -> GPU run ->
    sink), a graph with a host-side CPU block in the middle, fan-out, the WBFM-stereo and AM-synchronous composites, and
    start()/status()/wait()/stop().

Tolerance unless stated: |got - ref| <= 1e-5 * max(1, ||ref||_inf)."""
import time

import numpy as np
import pytest

import luaradio_b200 as radio
from luaradio_b200.block import Block, Input, Output
from luaradio_b200.types import ComplexFloat32, Float32, Vector
from oracle import lr_oracle as O
from tests.golden_util import GOLDEN_DIR, epsilon_ok, load_spec
from tests.oracle_blocks import BINARY_OPS

pytestmark = pytest.mark.gpu


def close(got, ref, rel=1e-5, absolute=None):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, "length %s != %s" % (got.shape, ref.shape)
    if ref.size == 0:
        return
    tol = absolute if absolute is not None else rel * max(1.0, float(np.max(np.abs(ref))))
    err = float(np.max(np.abs(got.astype(np.complex128) - ref.astype(np.complex128))))
    assert err <= tol, "max abs err %.3g > %.3g" % (err, tol)


def mk(cls, args, in_types, rate=2.0):
    b = cls(*args)
    b.get_rate = lambda: rate
    b.differentiate(in_types)
    b.initialize()
    return b


def rnd_c(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


@pytest.mark.parametrize("spec", ["multiply_spec", "multiplyconjugate_spec", "add_spec", "subtract_spec"])
def test_two_input_blocks_golden(spec):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        t = ComplexFloat32 if np.iscomplexobj(v["inputs"][0]) else Float32
        blk = mk(getattr(radio, block), [], [t, t])
        a, b = v["inputs"]
        got = np.array(blk.process(Vector.cast(a), Vector.cast(b)).data, copy=True)
        ok, msg = epsilon_ok(got, v["outputs"][0], eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)
        # split calls (the jig's sample-by-sample mode, in threes)
        outs = [np.array(blk.process(Vector.cast(a[i:i + 3]), Vector.cast(b[i:i + 3])).data, copy=True) for i in range(0, len(a), 3)]
        ok, msg = epsilon_ok(np.concatenate(outs), v["outputs"][0], eps)
        assert ok, msg
        blk.cleanup()


def test_two_input_blocks_long_streams():
    rng = np.random.default_rng(21)
    n = 1000003
    for cls, op in ((radio.MultiplyBlock, "multiply"), (radio.MultiplyConjugateBlock, "multiplyconjugate"),
                    (radio.AddBlock, "add"), (radio.SubtractBlock, "subtract")):
        a, b = rnd_c(rng, n), rnd_c(rng, n)
        blk = mk(cls, [], [ComplexFloat32, ComplexFloat32])
        close(np.array(blk.process(Vector.cast(a), Vector.cast(b)).data), O.binary_op(op, a, b))
        blk.cleanup()
        if op != "multiplyconjugate":
            ar, br = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
            blk = mk(cls, [], [Float32, Float32])
            close(np.array(blk.process(Vector.cast(ar), Vector.cast(br)).data), O.binary_op(op, ar, br))
            blk.cleanup()


def test_delay_golden_and_stream():
    block, vectors, eps = load_spec("delay_spec")
    done = 0
    for v in vectors:
        x = v["inputs"][0]
        if x.dtype not in (np.complex64, np.float32):
            continue                                   # Bit / Byte streams are not on this path
        t = ComplexFloat32 if np.iscomplexobj(x) else Float32
        blk = mk(radio.DelayBlock, v["args"], [t])
        ok, msg = epsilon_ok(np.array(blk.process(Vector.cast(x)).data, copy=True), v["outputs"][0], eps)
        assert ok, "%s: %s" % (v["desc"], msg)
        blk.cleanup()
        blk = mk(radio.DelayBlock, v["args"], [t])
        outs = [np.array(blk.process(Vector.cast(x[i:i + 1])).data, copy=True) for i in range(len(x))]
        assert np.array_equal(np.concatenate(outs), v["outputs"][0])
        blk.cleanup()
        done += 1
    assert done >= 4
    rng = np.random.default_rng(3)
    x = rnd_c(rng, 300000)
    for D in (1, 129, 5000):
        blk, ref = mk(radio.DelayBlock, [D], [ComplexFloat32]), O.Delay(D)
        outs, refs, i = [], [], 0
        for k in (7, 100, 4000, 200000, 1, 95892):
            outs.append(np.array(blk.process(Vector.cast(x[i:i + k])).data, copy=True))
            refs.append(ref.process(x[i:i + k]))
            i += k
        assert np.array_equal(np.concatenate(outs), np.concatenate(refs))       # pure data movement: bit exact
        blk.cleanup()


def test_psd_reference_vectors_and_frames():
    z = np.load(GOLDEN_DIR + "/spectrum_utils_vectors.npz")
    PSD = radio.spectrum_utils.PSD
    for kind in ("complex", "real"):
        x = z[kind + "_test_vector"]
        for win in ("rectangular", "hamming"):
            ok, msg = epsilon_ok(PSD(len(x), kind == "complex", win, 44100, False).compute(x), z["%s_test_vector_%s_psd" % (kind, win)], 1e-5)
            assert ok, (kind, win, msg)
            ok, msg = epsilon_ok(PSD(len(x), kind == "complex", win, 44100, True).compute(x), z["%s_test_vector_%s_psd_log" % (kind, win)], 3)
            assert ok, (kind, win, "log", msg)
    assert np.array_equal(radio.spectrum_utils.fftshift(z["complex_test_vector"]), z["complex_test_vector_fftshift"])
    # many frames at the spectrum sinks' sizes, against the oracle frame by frame
    rng = np.random.default_rng(8)
    for N in (64, 1024, 4096):
        x = (rnd_c(rng, N * 37) * 0.3 + np.exp(2j * np.pi * 0.123 * np.arange(N * 37))).astype(np.complex64)
        got = PSD(N, True, "hamming", 1e6, False).compute(x)
        ref = np.concatenate([O.psd(x[i:i + N], "hamming", 1e6, False) for i in range(0, len(x), N)])
        close(got, ref, rel=2e-5)
    with pytest.raises(radio._lib.LibraryError):
        PSD(1000, True)                                # frame lengths are powers of two on the GPU


def test_pll_matches_the_restatement():
    """PLLBlock(100, 19e3-50, 19e3+50, 2) on a noisy 19 kHz pilot at 220.5 kHz, in ragged calls.  UNPINNED against the
    reference (no pll_spec there): the oracle is the operation-by-operation restatement of pll.lua:140-170."""
    rate, n = 220500.0, 60000
    rng = np.random.default_rng(5)
    t = np.arange(n) / rate
    x = (0.8 * np.exp(2j * np.pi * 19000.3 * t + 0.4j) + 0.05 * rnd_c(rng, n)).astype(np.complex64)
    blk = mk(radio.PLLBlock, [100, 19e3 - 50, 19e3 + 50, 2], [ComplexFloat32], rate)
    ref = O.PLL(100, 19e3 - 50, 19e3 + 50, 2, rate)
    outs, errs, i = [], [], 0
    for k in (1, 999, 20000, 5, 38995):
        o, e = blk.process(Vector.cast(x[i:i + k]))
        outs.append(np.array(o.data, copy=True))
        errs.append(np.array(e.data, copy=True))
        i += k
    ro, re_ = ref.process(x)
    # float32 atan2f / double sincos of two libraries: 2e-6 after 60 000 steps of a contracting loop
    close(np.concatenate(errs), re_, absolute=2e-5)
    close(np.concatenate(outs), ro, absolute=2e-5)
    assert abs(np.mean(np.concatenate(errs)[-5000:])) < 0.05          # it locked
    blk.cleanup()


class HostGain(Block):
    """A CPU block in the middle of a GPU graph (numpy): the scheduler must leave it on the host and put a GPU run on
    either side."""
    name = "HostGain"

    def instantiate(self, g):
        self.g = g
        self.add_type_signature([Input("in", ComplexFloat32)], [Output("out", ComplexFloat32)])

    def process(self, x):
        return Vector.cast((np.asarray(x.data) * np.float32(self.g)).astype(np.complex64))


def test_top_spec_topology_two_sources_into_a_gpu_run():
    """tests/top_spec.lua:14-55: two sources -> MultiplyConjugate -> Lowpass(16, 100e3) -> FrequencyDiscriminator(5) ->
    Decimator(25) -> sink, on the reference's own source vectors and its committed sink vector."""
    z = np.load(GOLDEN_DIR + "/top_vectors.npz")
    src1, src2 = radio.ArraySource(z["SRC1_TEST_VECTOR"], 1e6, 100), radio.ArraySource(z["SRC2_TEST_VECTOR"], 1e6, 100)
    mixer, f1, fm, dec, snk = radio.MultiplyConjugateBlock(), radio.LowpassFilterBlock(16, 100e3), radio.FrequencyDiscriminatorBlock(5), \
        radio.DecimatorBlock(25, {"num_taps": 16}), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src1, "out", mixer, "in1")
    top.connect(src2, "out", mixer, "in2")
    top.connect(mixer, f1, fm, dec, snk)
    top.run()
    ok, msg = epsilon_ok(snk.result(), z["SNK_TEST_VECTOR"], 1e-6)
    assert ok, msg
    desc = top.describe_gpu_graph()
    assert desc.count(";") == 0 and desc.count("|") >= 1, desc        # ONE device flow graph behind the two-input block
    # and a second run() of the same top block works from clean ports
    for s in (src1, src2):
        s.pos = 0
    snk.chunks = []
    top.run()
    ok, msg = epsilon_ok(snk.result(), z["SNK_TEST_VECTOR"], 1e-6)
    assert ok, msg


def test_cpu_block_between_two_gpu_runs_and_fanout():
    rng = np.random.default_rng(4)
    n, rate = 300000, 1e6
    x = rnd_c(rng, n)
    src = radio.ArraySource(x, rate, 65536)
    a1, a2 = radio.FrequencyTranslatorBlock(1e5), radio.LowpassFilterBlock(64, 2e5)
    host = HostGain(0.5)
    b1, b2 = radio.DownsamplerBlock(2), radio.ComplexMagnitudeBlock()
    c1 = radio.ComplexToRealBlock()
    s1, s2 = radio.ArraySink(), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, a1, a2, host, b1)
    top.connect(b1, b2, s1)              # fan-out of b1: one branch magnitude ...
    top.connect(b1, c1, s2)              # ... one branch real part
    top.run()
    mid = (O.Chain(O.FrequencyTranslator(1e5, rate), O.lowpass_filter(64, 2e5, rate, True)).process(x) * np.float32(0.5)).astype(np.complex64)
    d = O.Downsampler(2).process(mid)
    close(s1.result(), O.complex_magnitude(d))
    close(s2.result(), O.complex_to_real(d))
    # one fused run before the host block, one device DAG (fan-out) behind it
    desc = top.describe_gpu_graph()
    assert "rot+fir" in desc and "dag{" in desc, desc


def stereo_mpx(n, rate, rng):
    """An FM-modulated stereo multiplex: L+R, 19 kHz pilot, (L-R) on 38 kHz DSB-SC."""
    t = np.arange(n) / rate
    left, right = 0.5 * np.sin(2 * np.pi * 700 * t), 0.4 * np.sin(2 * np.pi * 2300 * t)
    mpx = 0.45 * (left + right) + 0.1 * np.sin(2 * np.pi * 19e3 * t) + 0.45 * (left - right) * np.sin(2 * np.pi * 38e3 * t)
    phase = 2 * np.pi * 75e3 * np.cumsum(mpx) / rate
    return (np.exp(1j * phase) + 0.001 * rnd_c(rng, n)).astype(np.complex64), left, right


def test_wbfm_stereo_demodulator_dag():
    """composites/wbfmstereodemodulator.lua:22-64 at 220.5 kHz (the rate rtlsdr_wbfm_stereo.lua runs it at), against the
    oracle wired block by block the same way (that wiring is pinned on the CPU against the reference's own composite executed in
    the test interpreter: tests/test_oracle_golden.py::test_oracle_stereo_against_the_reference_executed_golden)."""
    rate, n = 220500.0, 150000
    rng = np.random.default_rng(9)
    x, left, right = stereo_mpx(n, rate, rng)
    src, sl, sr = radio.ArraySource(x, rate, 50000), radio.ArraySink(), radio.ArraySink()
    demod = radio.WBFMStereoDemodulator()
    top = radio.CompositeBlock()
    top.connect(src, demod)
    top.connect(demod, "left", sl, "in")
    top.connect(demod, "right", sr, "in")
    top.run()
    # oracle, block by block (wbfmstereodemodulator.lua:29-62)
    fm = O.FrequencyDiscriminator(1.25).process(x)
    hil = O.HilbertTransform(129).process(fm)
    pilot = O.complex_bandpass_filter(129, [18e3, 20e3], rate).process(hil)
    pll_out, _ = O.PLL(100, 19e3 - 50, 19e3 + 50, 2, rate).process(pilot)
    dly = O.Delay(129).process(hil)
    lpr = O.complex_to_real(O.lowpass_filter(128, 15e3, rate, True).process(dly))
    lmr = O.complex_to_real(O.lowpass_filter(128, 15e3, rate, True).process(O.binary_op("multiplyconjugate", dly, pll_out)))
    b, a = O.fm_deemphasis_taps(75e-6, rate)
    ref_l = O.IIRFilterFast(b, a, False).process(O.binary_op("add", lpr, lmr))
    ref_r = O.IIRFilterFast(b, a, False).process(O.binary_op("subtract", lpr, lmr))
    # While the PLL ACQUIRES (phase detector near +-pi: an unstable equilibrium) 1e-7 differences in its input are
    # amplified, so the comparison is loose there.  Once locked the loop contracts differences in phi_locked, but the
    # MULTIPLIED phase keeps them: phi_multiplied - 2 phi_locked = -alpha * sum of ALL past phase errors (pll.lua:155-157),
    # so what the acquisition amplified stays as a constant phase offset of the 38 kHz carrier (measured: 3e-4 rad ->
    # 7e-5 in the L-R path).  Hence 5e-4 when locked; the PLL alone is compared at 2e-5 on a clean pilot above.
    lock = 60000
    for got, ref in ((sl.result(), ref_l), (sr.result(), ref_r)):
        assert got.shape == ref.shape
        d = np.abs(got - ref)
        assert float(d[:lock].max()) <= 5e-3, "acquisition: max err %.3g at %d" % (float(d[:lock].max()), int(d[:lock].argmax()))
        assert float(d[lock:].max()) <= 5e-4, "locked: max err %.3g at %d" % (float(d[lock:].max()), lock + int(d[lock:].argmax()))
    # it separates the channels: after lock, left carries the 700 Hz tone and not the 2300 Hz one
    got_l = sl.result()[60000:]
    spec = np.abs(np.fft.rfft(got_l * np.hanning(len(got_l))))
    f = np.fft.rfftfreq(len(got_l), 1 / rate)
    p700, p2300 = spec[np.argmin(np.abs(f - 700))], spec[np.argmin(np.abs(f - 2300))]
    assert p700 > 10 * p2300, (p700, p2300)
    # the whole demodulator is ONE device DAG; its linear runs (discriminator -> hilbert, two lowpass -> c2r) are fused
    # flow graphs inside it
    desc = top.describe_gpu_graph()
    assert desc.startswith("dag{") and desc.count(";") >= 6, desc
    # and the host-level scheduling of the same graph (chains + single blocks, a host hop at every junction) agrees
    src2, sl2, sr2 = radio.ArraySource(x, rate, 50000), radio.ArraySink(), radio.ArraySink()
    demod2 = radio.WBFMStereoDemodulator()
    top2 = radio.CompositeBlock()
    top2.connect(src2, demod2)
    top2.connect(demod2, "left", sl2, "in")
    top2.connect(demod2, "right", sr2, "in")
    top2.run(device_dag=False)
    assert "dag{" not in top2.describe_gpu_graph()
    close(sl2.result(), sl.result(), absolute=2e-6)
    close(sr2.result(), sr.result(), absolute=2e-6)


def test_am_synchronous_demodulator_dag():
    rate, n, ifreq = 48000.0, 120000, 5000.0
    rng = np.random.default_rng(10)
    t = np.arange(n) / rate
    env = 0.5 * (1 + 0.5 * np.sin(2 * np.pi * 440 * t))
    x = (env * np.exp(2j * np.pi * (ifreq + 3.0) * t + 0.7j) + 0.002 * rnd_c(rng, n)).astype(np.complex64)
    src, snk = radio.ArraySource(x, rate, 30000), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, radio.AMSynchronousDemodulator(ifreq, 5e3), snk)
    top.run()
    rf = O.complex_bandpass_filter(129, [ifreq - 5e3, ifreq + 5e3], rate).process(x)
    pll_out, _ = O.PLL(1000, ifreq - 100, ifreq + 100, None, rate).process(rf)
    b, a = O.singlepole_highpass_taps(100, rate)
    ref = O.Chain(O.complex_to_real, O.IIRFilterFast(b, a, False), O.lowpass_filter(128, 5e3, rate, False)).process(
        O.binary_op("multiplyconjugate", rf, pll_out))
    close(snk.result(), ref, absolute=5e-5)


def test_start_status_wait_stop():
    """composite.lua:534-545 (start), :858-877 (status), :913-930 (wait), :886-906 (stop)."""
    rng = np.random.default_rng(2)
    x = rnd_c(rng, 400000)
    src, snk = radio.ArraySource(x, 1e6, 4096), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, radio.LowpassFilterBlock(64, 1e5), radio.DownsamplerBlock(4), snk)
    assert top.status() == {"running": False}
    top.start()
    assert top.status()["running"] in (True, False)
    with pytest.raises(RuntimeError):
        if top.status()["running"]:
            top.start()
        else:
            raise RuntimeError("already finished")
    top.wait()
    assert top.status() == {"running": False}
    close(snk.result(), O.Chain(O.lowpass_filter(64, 1e5, 1e6, True), O.Downsampler(4)).process(x))
    # stop() ends an endless source
    src2 = radio.ArraySource(x, 1e6, 4096)
    src2.process_orig = src2.process

    def forever():
        v = src2.process_orig()
        if v is None:
            src2.pos = 0
            v = src2.process_orig()
        return v
    src2.process = forever
    snk2 = radio.ArraySink()
    snk2.process = lambda v: None
    top2 = radio.CompositeBlock()
    top2.connect(src2, radio.LowpassFilterBlock(64, 1e5), snk2)
    top2.start()
    time.sleep(0.3)
    assert top2.status()["running"]
    top2.stop()
    assert top2.status() == {"running": False}


def test_superchunk_scheduler_with_small_vectors():
    """The reference's 8192-sample source vectors through the scheduler's super-chunk mode: same stream, flushed at EOF."""
    n = 700000
    x = O.synth_fm_iq(0, n)

    def run(superchunk):
        src, snk = radio.ArraySource(x, 1102500.0, 8192), radio.ArraySink()
        top = radio.CompositeBlock()
        top.connect(src, radio.TunerBlock(-250e3, 200e3, 5), radio.FrequencyDiscriminatorBlock(1.25),
                    radio.LowpassFilterBlock(128, 15e3), radio.FMDeemphasisFilterBlock(75e-6), radio.DownsamplerBlock(5), snk)
        top.run(superchunk=superchunk)
        return snk.result()

    ref = O.wbfm_mono_chain().process(x)
    close(run(0), ref)
    close(run(1 << 17), ref)


def test_pll_chunk_parallel_mode_when_locked():
    """lrb200_pll_set_mode(1): every chunk simulated by its own thread after a lead-in, the multiplied phase rebuilt from
    prefix sums.  Equal to the sequential recurrence while the loop is locked; the first chunk (carried state) is exact."""
    rate, n = 220500.0, 700000
    rng = np.random.default_rng(6)
    t = np.arange(n) / rate
    x = (0.8 * np.exp(2j * np.pi * 19000.3 * t + 0.4j) + 0.02 * rnd_c(rng, n)).astype(np.complex64)
    ref = O.PLL(100, 19e3 - 50, 19e3 + 50, 2, rate)
    ro, re_ = ref.process(x)
    for multiplier, ref_out in ((2, ro), (0.25, None)):
        blk = radio.PLLBlock(100, 19e3 - 50, 19e3 + 50, multiplier)
        blk.parallel = True
        blk.get_rate = lambda: rate
        blk.differentiate([ComplexFloat32])
        blk.initialize()
        # first call: acquisition inside chunk 0 of a short (sequential) call, then two long chunk-parallel calls
        cuts = [(0, 60000), (60000, 380000), (380000, n)]
        outs, errs = [], []
        for a, b in cuts:
            o, e = blk.process(Vector.cast(x[a:b]))
            outs.append(np.array(o.data, copy=True))
            errs.append(np.array(e.data, copy=True))
        got_o, got_e = np.concatenate(outs), np.concatenate(errs)
        if ref_out is None:
            ref_out, _ = O.PLL(100, 19e3 - 50, 19e3 + 50, multiplier, rate).process(x)
        close(got_e[:60000], re_[:60000], absolute=2e-5)            # sequential part: as test_pll_matches_the_restatement
        d_e = float(np.max(np.abs(got_e[60000:] - re_[60000:])))
        d_o = float(np.max(np.abs(got_o[60000:] - ref_out[60000:])))
        assert d_e <= 2e-4 and d_o <= 2e-4, (multiplier, d_e, d_o)
        blk.cleanup()
