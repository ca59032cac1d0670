"""GPU parity against the reference's committed golden vectors, through the block API and the C ABI.

Same procedure as the reference's jig (tests/jigs.lua:191-250): every vector is run (a) whole and
(b) sample by sample (256 process() calls of length 1, which pins the carried streaming state), and
compared with the reference's absolute epsilon (1e-6; 1e-5 for translator/tuner)."""
import numpy as np
import pytest

from tests.blocks_util import create_block, run_composite, run_sample_by_sample, run_whole
from tests.golden_util import epsilon_ok, load_spec

pytestmark = pytest.mark.gpu

BLOCK_SPECS = [
    "firfilter_spec", "lowpassfilter_spec", "highpassfilter_spec", "bandpassfilter_spec", "bandstopfilter_spec",
    "complexbandpassfilter_spec", "complexbandstopfilter_spec", "hilberttransform_spec", "frequencytranslator_spec",
    "frequencydiscriminator_spec", "downsampler_spec", "fmdeemphasisfilter_spec", "singlepolelowpassfilter_spec",
    "singlepolehighpassfilter_spec", "iirfilter_spec", "complexmagnitude_spec", "complextoreal_spec",
    "multiplyconstant_spec", "upsampler_spec",
]


def _golden_out(block, v):
    """The reference's FFT-mode FIR vectors are truncated to whole overlap-save blocks
    (firfilter_spec.py:22-28); the GPU block is length-preserving, so compare that prefix."""
    return v["outputs"][0]


@pytest.mark.parametrize("spec", BLOCK_SPECS)
def test_golden_whole_vector(spec):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        blk = create_block(block, v["args"], v["inputs"])
        got = run_whole(blk, v["inputs"][0])
        want = _golden_out(block, v)
        if block == "FIRFilterBlock" and len(want) < len(got):
            got = got[:len(want)]
        ok, msg = epsilon_ok(got, want, eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)
        blk.cleanup()


@pytest.mark.parametrize("spec", BLOCK_SPECS)
def test_golden_sample_by_sample(spec):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        blk = create_block(block, v["args"], v["inputs"])
        want = _golden_out(block, v)
        got = run_sample_by_sample(blk, v["inputs"][0], want.dtype)
        if block == "FIRFilterBlock" and len(want) < len(got):
            got = got[:len(want)]
        ok, msg = epsilon_ok(got, want, eps)
        assert ok, "%s / %s: %s" % (block, v["desc"], msg)
        blk.cleanup()


@pytest.mark.parametrize("spec", ["tuner_spec", "decimator_spec", "interpolator_spec", "rationalresampler_spec"])
@pytest.mark.parametrize("fuse", [True, False])
def test_golden_composites(spec, fuse):
    block, vectors, eps = load_spec(spec)
    for v in vectors:
        got, top = run_composite(block, v["args"], v["inputs"][0], fuse=fuse)
        ok, msg = epsilon_ok(got, v["outputs"][0], eps)
        assert ok, "%s / %s (fuse=%s, graph=%s): %s" % (block, v["desc"], fuse, top.describe_gpu_graph(), msg)
        if fuse:
            assert "fused" in top.describe_gpu_graph() or "tuner" in top.describe_gpu_graph(), top.describe_gpu_graph()


@pytest.mark.parametrize("chunk", [1 << 22, 7, 1])
def test_golden_composites_ragged_source(chunk):
    """Same composites with the source delivering tiny vectors (state carried across graph executes)."""
    for spec in ("tuner_spec", "decimator_spec", "interpolator_spec", "rationalresampler_spec"):
        block, vectors, eps = load_spec(spec)
        for v in (vectors[:2] if spec in ("tuner_spec", "decimator_spec") else vectors):
            got, _ = run_composite(block, v["args"], v["inputs"][0], chunk=chunk)
            ok, msg = epsilon_ok(got, v["outputs"][0], eps)
            assert ok, "%s / %s chunk=%d: %s" % (block, v["desc"], chunk, msg)


def test_golden_top_chain():
    """tests/top_spec.lua:14-55: Lowpass(16,100e3) -> FrequencyDiscriminator(5) -> Decimator(25,{num_taps=16})
    at 1e6 S/s, on the reference's own source vectors (the MultiplyConjugate of the two sources is
    formed on the host: that block is outside the hot path)."""
    import luaradio_b200 as radio
    z = np.load(__import__("tests.golden_util", fromlist=["GOLDEN_DIR"]).GOLDEN_DIR + "/top_vectors.npz")
    x = (z["SRC1_TEST_VECTOR"].astype(np.complex128) * np.conj(z["SRC2_TEST_VECTOR"].astype(np.complex128))).astype(np.complex64)
    src, snk = radio.ArraySource(x, 1e6), radio.ArraySink()
    top = radio.CompositeBlock()
    top.connect(src, radio.LowpassFilterBlock(16, 100e3), radio.FrequencyDiscriminatorBlock(5.0),
                radio.DecimatorBlock(25, {"num_taps": 16}), snk)
    top.run(False)
    ok, msg = epsilon_ok(snk.result(), z["SNK_TEST_VECTOR"], 1e-6)
    assert ok, msg


def test_golden_iq_file_source_formats():
    """tests/blocks/sources/iqfile_spec.py: the 14 IQFileSource sample formats, converted on the GPU:
    (a) source.process() chunk by chunk through the C ABI, (b) as the first stage of a GPU flow graph fed raw bytes."""
    import luaradio_b200 as radio
    from tests.golden_util import GOLDEN_DIR
    z = np.load(GOLDEN_DIR + "/iqfile_spec_raw.npz")
    for i, fmt in enumerate(z["formats"]):
        raw, want = z["v%d_raw" % i].tobytes(), z["v%d_out" % i]
        src = radio.IQFileSource(raw, str(fmt), 1, chunk=100)
        src.differentiate([])
        src.initialize()
        outs = []
        while True:
            v = src.process()
            if v is None:
                break
            outs.append(np.array(v.data, copy=True))
        ok, msg = epsilon_ok(np.concatenate(outs), want, 1e-6)
        assert ok, "%s process(): %s" % (fmt, msg)
        src.cleanup()
        # in a flow graph: IQFileSource -> ComplexToReal -> sink  (raw bytes in, converter fused as stage 0)
        src, snk = radio.IQFileSource(raw, str(fmt), 1, chunk=77), radio.ArraySink()
        top = radio.CompositeBlock()
        top.connect(src, radio.ComplexToRealBlock(), snk)
        top.run(False)
        ok, msg = epsilon_ok(snk.result(), want.real.astype(np.float32), 1e-6)
        assert ok, "%s graph: %s" % (fmt, msg)
        assert top.describe_gpu_graph().startswith("iqconv(%s)" % fmt), top.describe_gpu_graph()
    with pytest.raises(AssertionError):
        radio.IQFileSource(b"", "u7", 1)


def test_golden_real_file_source_formats():
    """tests/blocks/sources/realfile_spec.gen.lua: the 14 RealFileSource formats, converted on the GPU."""
    import luaradio_b200 as radio
    from tests.golden_util import GOLDEN_DIR
    z = np.load(GOLDEN_DIR + "/realfile_spec_raw.npz")
    for i, fmt in enumerate(z["formats"]):
        src = radio.RealFileSource(z["v%d_raw" % i].tobytes(), str(fmt), 1, chunk=100)
        src.differentiate([])
        src.initialize()
        outs = []
        while True:
            v = src.process()
            if v is None:
                break
            outs.append(np.array(v.data, copy=True))
        ok, msg = epsilon_ok(np.concatenate(outs), z["v%d_out" % i], 1e-6)
        assert ok, "%s: %s" % (fmt, msg)
        src.cleanup()


def test_file_sinks_bit_exact_and_wav_header():
    """IQFileSink / RealFileSink / WAVFileSink (tests/blocks/sinks/*_spec.lua): the device conversion equals the oracle
    byte for byte for all 14 formats (odd lengths, unaligned tails), and the WAV file equals header + data."""
    import io
    import luaradio_b200 as radio
    from oracle import lr_oracle as O
    from tests.test_oracle_golden import WAV_HEADERS
    rng = np.random.default_rng(11)
    for n in (1, 7, 256, 100003):
        xr = rng.uniform(-1, 1, n).astype(np.float32)
        xc = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
        xr[:1], xc[:1] = 1.0, -1.0 + 1.0j
        for fmt in O.IQ_FORMATS:
            for cls, x, t in ((radio.RealFileSink, xr, radio.types.Float32), (radio.IQFileSink, xc, radio.types.ComplexFloat32)):
                f = io.BytesIO()
                snk = cls(f, fmt)
                snk.differentiate([t])
                snk.initialize()
                snk.process(radio.types.Vector.cast(x))
                snk.cleanup()
                assert f.getvalue() == O.file_sink_convert(x, fmt).tobytes(), "%s %s n=%d" % (cls.name, fmt, n)
    x1, x2 = rng.uniform(-1, 1, 256).astype(np.float32), rng.uniform(-1, 1, 256).astype(np.float32)
    for (bits, ch), hexs in WAV_HEADERS.items():
        f = io.BytesIO()
        snk = radio.WAVFileSink(f, ch, bits)
        snk.get_rate = lambda: 44100
        snk.differentiate([radio.types.Float32] * ch)
        snk.initialize()
        snk.process(*[radio.types.Vector.cast(c) for c in (x1, x2)[:ch]])
        snk.cleanup()
        data = x1 if ch == 1 else np.stack([x1, x2], 1).reshape(-1)
        want = bytes.fromhex(hexs.replace(" ", "")) + O.file_sink_convert(data, radio.WAVFileSink.WAVE_FORMATS[bits]).tobytes()
        assert f.getvalue() == want, (bits, ch)
    with pytest.raises(AssertionError):
        radio.WAVFileSink(io.BytesIO(), 1, 24)


def test_iqconv_8bit_all_values_bit_exact():
    """Every possible u8 / s8 byte through the device converters equals the oracle's double -> float32 result bit for
    bit (the device uses a 3-instruction division), on the vectorised and the scalar (unaligned tail) paths."""
    import luaradio_b200 as radio
    from oracle import lr_oracle as O
    raw = np.tile(np.arange(256, dtype=np.uint8), 9)[:2 * 1027]          # odd sample count: vector body + scalar tail
    for fmt in ("u8", "s8"):
        for cls, conv in ((radio.IQFileSource, O.iq_file_convert), (radio.RealFileSource, O.real_file_convert)):
            src = cls(raw.tobytes(), fmt, 1, chunk=4096)
            src.differentiate([])
            src.initialize()
            got = np.array(src.process().data, copy=True)
            src.cleanup()
            assert np.array_equal(got.view(np.uint32), conv(raw, fmt).view(np.uint32)), (fmt, cls.name)


def test_iqconv_16bit_all_values_bit_exact():
    """Every 16-bit sample value, both signednesses and byte orders, through the device converters (vector body and
    scalar tail) == the oracle's double -> float32 result bit for bit."""
    import luaradio_b200 as radio
    from oracle import lr_oracle as O
    vals = np.arange(65536, dtype=np.uint16)
    vals = np.concatenate([vals, vals[:3]])                       # 65539 components -> odd sample count for the real path
    for fmt in ("u16le", "u16be", "s16le", "s16be"):
        raw = vals.astype(">u2" if fmt.endswith("be") else "<u2").tobytes()
        src = radio.RealFileSource(raw, fmt, 1, chunk=1 << 20)
        src.differentiate([])
        src.initialize()
        got = np.array(src.process().data, copy=True)
        src.cleanup()
        assert np.array_equal(got.view(np.uint32), O.real_file_convert(np.frombuffer(raw, np.uint8), fmt).view(np.uint32)), fmt
        raw_iq = raw[:4 * 32767]                                  # an odd number of I/Q samples: vector body + scalar tail
        src = radio.IQFileSource(raw_iq, fmt, 1, chunk=1 << 20)
        src.differentiate([])
        src.initialize()
        got = np.array(src.process().data, copy=True)
        src.cleanup()
        assert np.array_equal(got.view(np.uint32), O.iq_file_convert(np.frombuffer(raw_iq, np.uint8), fmt).view(np.uint32)), fmt
