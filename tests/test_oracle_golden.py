"""Pin the oracle: every committed golden vector of the reference's own specs for this path
(SURVEY.md section 8c), whole-vector AND sample-by-sample (tests/jigs.lua:191-250), at the
reference's epsilons."""
import numpy as np
import pytest

from oracle import lr_oracle as O
from tests.golden_util import all_block_specs, epsilon_ok, load_spec, GOLDEN_DIR
from tests.oracle_blocks import make_oracle

TWO_INPUT = ("multiplyconjugate_spec", "multiply_spec", "add_spec", "subtract_spec")
SPECS = [s for s in all_block_specs() if s not in TWO_INPUT]


@pytest.mark.parametrize("spec", SPECS)
def test_oracle_whole_vector(spec):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        o = make_oracle(block, v["args"], v["inputs"])
        got = o.process(v["inputs"][0])
        ok, msg = epsilon_ok(got, v["outputs"][0], eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)


@pytest.mark.parametrize("spec", [s for s in SPECS if s not in ("tuner_spec", "decimator_spec")])
def test_oracle_sample_by_sample(spec):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        o = make_oracle(block, v["args"], v["inputs"])
        x = v["inputs"][0]
        outs = [o.process(x[i:i + 1]) for i in range(len(x))]
        got = np.concatenate(outs) if outs else np.zeros(0, v["outputs"][0].dtype)
        ok, msg = epsilon_ok(got, v["outputs"][0], eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)


def test_oracle_multiply_conjugate():
    block, vectors, eps = load_spec("multiplyconjugate_spec")
    for v in vectors:
        ok, msg = epsilon_ok(O.multiply_conjugate(*v["inputs"]), v["outputs"][0], eps)
        assert ok, msg


@pytest.mark.parametrize("spec", TWO_INPUT)
def test_oracle_two_input_blocks(spec):
    from tests.oracle_blocks import BINARY_OPS
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        ok, msg = epsilon_ok(O.binary_op(BINARY_OPS[block], *v["inputs"]), v["outputs"][0], eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)


def test_oracle_psd_vectors():
    """tests/utilities/spectrum_utils_spec.lua:58-72: PSD of the committed test vectors, rectangular and hamming, linear at
    1e-5 and logarithmic within 3 (dB), exactly the reference's tolerances."""
    z = np.load(GOLDEN_DIR + "/spectrum_utils_vectors.npz")
    for kind in ("complex", "real"):
        x = z[kind + "_test_vector"]
        for win in ("rectangular", "hamming"):
            ok, msg = epsilon_ok(O.psd(x, win, 44100, False), z["%s_test_vector_%s_psd" % (kind, win)], 1e-5)
            assert ok, (kind, win, msg)
            ok, msg = epsilon_ok(O.psd(x, win, 44100, True), z["%s_test_vector_%s_psd_log" % (kind, win)], 3)
            assert ok, (kind, win, "log", msg)


def test_oracle_tap_design():
    """tests/utilities/filter_utils_spec.lua:8-58 and window_utils_spec.lua at 1e-6."""
    z = np.load(GOLDEN_DIR + "/filter_utils_vectors.npz")
    cases = {
        "firwin_lowpass": O.firwin_lowpass(128, 0.5),
        "firwin_highpass": O.firwin_highpass(129, 0.5),
        "firwin_bandpass": O.firwin_bandpass(129, [0.4, 0.6]),
        "firwin_bandstop": O.firwin_bandstop(129, [0.4, 0.6]),
        "firwin_complex_bandpass_positive": O.firwin_complex_bandpass(129, [0.1, 0.3]),
        "firwin_complex_bandpass_negative": O.firwin_complex_bandpass(129, [-0.1, -0.3]),
        "firwin_complex_bandpass_zero": O.firwin_complex_bandpass(129, [-0.2, 0.2]),
        "firwin_complex_bandstop_positive": O.firwin_complex_bandstop(129, [0.1, 0.3]),
        "firwin_complex_bandstop_negative": O.firwin_complex_bandstop(129, [-0.1, -0.3]),
        "firwin_complex_bandstop_zero": O.firwin_complex_bandstop(129, [-0.2, 0.2]),
        "fir_hilbert_transform": O.fir_hilbert_transform(129),
    }
    for k, h in cases.items():
        ok, msg = epsilon_ok(O.f32_taps(h), z[k], 1e-6)
        assert ok, "%s: %s" % (k, msg)
    w = np.load(GOLDEN_DIR + "/window_utils_vectors.npz")
    for name in ("rectangular", "hamming", "hanning", "bartlett", "blackman"):
        for per in (False, True):
            key = "window_" + name + ("_periodic" if per else "")
            ok, msg = epsilon_ok(O.window(len(w[key]), name, per).astype(np.float32), w[key], 1e-6)
            assert ok, "%s: %s" % (key, msg)


def test_oracle_dft():
    """tests/utilities/spectrum_utils_spec.lua:58-72 at 1e-5 (forward e^{-j}, inverse 1/N)."""
    z = np.load(GOLDEN_DIR + "/spectrum_utils_vectors.npz")
    ok, msg = epsilon_ok(O.dft(z["complex_test_vector"]), z["complex_test_vector_dft"], 1e-5)
    assert ok, msg
    ok, msg = epsilon_ok(O.dft(z["real_test_vector"]), z["real_test_vector_dft"], 1e-5)
    assert ok, msg
    ok, msg = epsilon_ok(O.idft(z["complex_test_vector_dft"]), z["complex_test_vector"], 1e-5)
    assert ok, msg


def test_oracle_top_chain():
    """tests/top_spec.lua:14-55 against tests/top_vectors.py:6-24: MultiplyConjugate -> Lowpass(16, 100e3)
    -> FrequencyDiscriminator(5) -> Decimator(25, num_taps=16) at rate 1e6, epsilon 1e-6."""
    z = np.load(GOLDEN_DIR + "/top_vectors.npz")
    x = O.multiply_conjugate(z["SRC1_TEST_VECTOR"], z["SRC2_TEST_VECTOR"])
    x = O.lowpass_filter(16, 100e3, 1e6, True).process(x)
    x = O.FrequencyDiscriminator(5.0).process(x)
    x = O.decimator(25, False, num_taps=16).process(x)
    ok, msg = epsilon_ok(x, z["SNK_TEST_VECTOR"], 1e-6)
    assert ok, msg


def test_oracle_chunking_invariance():
    """Streaming state: ragged chunking must not change the WBFM-mono chain output."""
    x = O.synth_fm_iq(0, 40000)
    whole = O.wbfm_mono_chain().process(x)
    c = O.wbfm_mono_chain()
    rng = np.random.default_rng(3)
    outs, i = [], 0
    while i < len(x):
        n = int(rng.integers(0, 3000))
        outs.append(c.process(x[i:i + n]))
        i += n
    got = np.concatenate(outs)
    assert got.shape == whole.shape
    assert np.max(np.abs(got - whole)) < 1e-6


def test_oracle_iq_file_formats():
    """tests/blocks/sources/iqfile_spec.py: all 14 sample formats of IQFileSource at 1e-6."""
    z = np.load(GOLDEN_DIR + "/iqfile_spec_raw.npz")
    for i, fmt in enumerate(z["formats"]):
        ok, msg = epsilon_ok(O.iq_file_convert(z["v%d_raw" % i], str(fmt)), z["v%d_out" % i], 1e-6)
        assert ok, "%s: %s" % (fmt, msg)


def test_oracle_real_file_formats():
    """tests/blocks/sources/realfile_spec.gen.lua: the 14 RealFileSource formats (raw bytes -> Float32), epsilon 1e-6."""
    z = np.load(GOLDEN_DIR + "/realfile_spec_raw.npz")
    assert len(z["formats"]) == 14
    for i, fmt in enumerate(z["formats"]):
        ok, msg = epsilon_ok(O.real_file_convert(z["v%d_raw" % i], str(fmt)), z["v%d_out" % i], 1e-6)
        assert ok, "%s: %s" % (fmt, msg)


# WAV headers of the reference's sink spec (tests/blocks/sinks/wavfile_spec.lua:22-35): 256 samples per channel at 44100 Hz
WAV_HEADERS = {
    (8, 1): "524946462401000057415645666d74201000000001000100 44ac0000 44ac0000 0100 0800 64617461 00010000",
    (8, 2): "524946462402000057415645666d74201000000001000200 44ac0000 88580100 0200 0800 64617461 00020000",
    (16, 1): "524946462402000057415645666d74201000000001000100 44ac0000 88580100 0200 1000 64617461 00020000",
    (16, 2): "524946462404000057415645666d74201000000001000200 44ac0000 10b10200 0400 1000 64617461 00040000",
    (32, 1): "524946462404000057415645666d74201000000001000100 44ac0000 10b10200 0400 2000 64617461 00040000",
    (32, 2): "524946462408000057415645666d74201000000001000200 44ac0000 20620500 0800 2000 64617461 00080000",
}


def test_oracle_file_sinks_round_trip_and_wav_header():
    """tests/blocks/sinks/{iqfile,realfile,wavfile}_spec.lua write random samples and read them back within a per-format
    epsilon (1e-2 for 8 bit, 1e-4 for 16 bit, 1e-6 otherwise); the WAV headers are the spec's golden bytes."""
    rng = np.random.default_rng(5)
    xr = rng.uniform(-1, 1, 1000).astype(np.float32)
    xc = (rng.uniform(-1, 1, 1000) + 1j * rng.uniform(-1, 1, 1000)).astype(np.complex64)
    for fmt in O.IQ_FORMATS:
        eps = 1e-2 if "8" in fmt and "f" not in fmt else (1e-4 if "16" in fmt else 1e-6)
        back = O.real_file_convert(O.file_sink_convert(xr, fmt), fmt)
        assert np.max(np.abs(back - xr)) <= eps, fmt
        back = O.iq_file_convert(O.file_sink_convert(xc, fmt), fmt)
        assert np.max(np.abs(back - xc)) <= eps * 1.5, fmt
    # truncation toward zero, end points
    assert list(O.file_sink_convert(np.array([-1.0, -0.004, 0.0, 0.004, 1.0], np.float32), "u8")) == [0, 126, 127, 128, 255]
    assert list(O.file_sink_convert(np.array([-1.0, 1.0], np.float32), "s16le").view("<i2")) == [-32767, 32767]
    for (bits, ch), hexs in WAV_HEADERS.items():
        assert O.wav_header(256, ch, bits, 44100) == bytes.fromhex(hexs.replace(" ", "")), (bits, ch)


def test_oracle_chain_against_the_reference_executed_golden():
    """tests/golden/wbfm_chain_ref_executed.npz: the WHOLE rtlsdr_wbfm_mono chain as the stock reference computed it (its
    pure-Lua process() branches and CompositeBlock run loop executed from the reference tree in the test interpreter,
    tests/golden/make_chain_golden.py).  The oracle's chain -- whole and in the same ragged chunks -- must reproduce it at
    the reference specs' own 1e-6."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "wbfm_chain_ref_executed.npz"))
    x, y, rate = g["x"], g["y"], float(g["rate"])
    assert len(y) == len(x) // 25 and np.max(np.abs(y)) > 1e-3

    def chain():
        return O.Chain(O.tuner(-250e3, 200e3, 5, rate), O.FrequencyDiscriminator(1.25), O.lowpass_filter(128, 15e3, rate / 5, False),
                       O.IIRFilter(*O.fm_deemphasis_taps(75e-6, rate / 5), False), O.Downsampler(5))
    whole = chain().process(x)
    assert np.max(np.abs(whole - y)) <= 1e-6
    c = chain()
    s = list(g["splits"])
    parts = np.concatenate([c.process(x[a:b]) for a, b in zip(s[:-1], s[1:])])
    assert np.max(np.abs(parts - y)) <= 1e-6


def test_oracle_stereo_against_the_reference_executed_golden():
    """tests/golden/wbfm_stereo_ref_executed.npz: WBFMStereoDemodulator as the stock reference computed it (its composite,
    the pure-Lua process() branches of its 14 blocks -- PLL included -- and its run loop executed from the reference tree in
    the test interpreter, tests/golden/make_stereo_golden.py).  The oracle, wired block by block the same way
    (wbfmstereodemodulator.lua:29-62), reproduces both channels through the PLL's acquisition and lock."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "wbfm_stereo_ref_executed.npz"))
    x, rate = g["x"], float(g["rate"])
    fm = O.FrequencyDiscriminator(1.25).process(x)
    hil = O.HilbertTransform(129).process(fm)
    pilot = O.complex_bandpass_filter(129, [18e3, 20e3], rate).process(hil)
    pll_out, _ = O.PLL(100, 19e3 - 50, 19e3 + 50, 2, rate).process(pilot)
    dly = O.Delay(129).process(hil)
    lpr = O.complex_to_real(O.lowpass_filter(128, 15e3, rate, True).process(dly))
    lmr = O.complex_to_real(O.lowpass_filter(128, 15e3, rate, True).process(O.binary_op("multiplyconjugate", dly, pll_out)))
    b, a = O.fm_deemphasis_taps(75e-6, rate)
    left = O.IIRFilter(b, a, False).process(O.binary_op("add", lpr, lmr))
    right = O.IIRFilter(b, a, False).process(O.binary_op("subtract", lpr, lmr))
    assert len(left) == len(g["left"]) == len(x) and np.max(np.abs(left)) > 0.05
    # measured 3e-8 (float32 accumulation in the reference's Lua loops against float64 here); the specs' own epsilon is 1e-6
    assert np.max(np.abs(left - g["left"])) <= 1e-6 and np.max(np.abs(right - g["right"])) <= 1e-6
    # the demodulator separated something: the two channels differ, as L (700 Hz) and R (2300 Hz) of the multiplex do
    assert np.max(np.abs(g["left"][3000:] - g["right"][3000:])) > 0.02


def _composite_oracles(rate):
    """The oracle wired the way the reference's composites wire their blocks (composites/*.lua), by fixture name."""
    def iir_hp(cutoff):
        b, a = O.singlepole_highpass_taps(cutoff, rate)
        return O.IIRFilter(b, a, False)

    def am_sync(x):
        f = O.complex_bandpass_filter(129, [10e3 - 5e3, 10e3 + 5e3], rate).process(x)         # amsynchronousdemodulator.lua:25-45
        pll, _ = O.PLL(1000, 10e3 - 100, 10e3 + 100, 1.0, rate).process(f)
        d = O.complex_to_real(O.binary_op("multiplyconjugate", f, pll))
        return O.Chain(iir_hp(100), O.lowpass_filter(128, 5e3, rate, False)).process(d)
    return {
        "nbfm": O.Chain(O.lowpass_filter(128, 5e3 + 4e3, rate, True), O.FrequencyDiscriminator(5e3 / 4e3), O.lowpass_filter(128, 4e3, rate, False)).process,
        "am_envelope": O.Chain(O.complex_magnitude, iir_hp(100), O.lowpass_filter(128, 5e3, rate, False)).process,
        "ssb_usb": O.Chain(O.complex_bandpass_filter(129, [0, 3e3], rate), O.complex_to_real, O.lowpass_filter(128, 3e3, rate, False)).process,
        "ssb_lsb": O.Chain(O.complex_bandpass_filter(129, [0, -3e3], rate), O.complex_to_real, O.lowpass_filter(128, 3e3, rate, False)).process,
        "am_synchronous": am_sync,
        "decimator_c": O.decimator(4, True, 64).process,
        "interpolator_c": O.interpolator(3, True, 48).process,
        "resampler_c": O.rational_resampler(3, 2, True, 48).process,
        "interpolator_r": O.interpolator(2, False, 32).process,
    }


@pytest.mark.parametrize("name", ["nbfm", "am_envelope", "ssb_usb", "ssb_lsb", "am_synchronous", "decimator_c", "interpolator_c", "resampler_c",
                                  "interpolator_r"])
def test_oracle_composites_against_the_reference_executed_goldens(name):
    """tests/golden/composites_ref_executed.npz: the reference's NBFM / AM-envelope / SSB / AM-synchronous demodulators and
    its Decimator / Interpolator / RationalResampler composites as the stock reference computed them (composite wiring, the
    pure-Lua process() branches and the run loop executed from the reference tree in the test interpreter,
    tests/golden/make_composite_goldens.py).  The oracle wired the same way must reproduce them at the reference specs' 1e-6
    (the resampler composites scale by L: 1e-6 relative to that)."""
    import os
    g = np.load(os.path.join(GOLDEN_DIR, "composites_ref_executed.npz"))
    x, y, rate = g[name + "_x"], g[name + "_y"], float(g[name + "_rate"])
    got = _composite_oracles(rate)[name](x)
    assert len(got) == len(y) and np.max(np.abs(y)) > 1e-3
    assert np.max(np.abs(got - y)) <= 1e-6 * max(1.0, float(np.max(np.abs(y)))), float(np.max(np.abs(got - y)))
