"""Build ORACLE blocks (oracle/lr_oracle.py) by reference block name, for the parity tests."""
import numpy as np

from oracle import lr_oracle as O


def _is_c(x):
    return np.iscomplexobj(x)


def make_oracle(block, args, inputs, rate=2.0):
    cin = _is_c(inputs[0]) if inputs else False
    a = list(args)
    opt = lambda i, d=None: a[i] if len(a) > i and a[i] is not None else d
    if block == "FIRFilterBlock":
        taps = np.asarray(a[0])
        return (O.FIRFilterFFT if opt(1, False) else O.FIRFilter)(taps, cin)
    if block == "LowpassFilterBlock":
        return O.lowpass_filter(a[0], a[1], rate, cin, opt(2), opt(3, "hamming"))
    if block == "HighpassFilterBlock":
        return O.highpass_filter(a[0], a[1], rate, cin, opt(2), opt(3, "hamming"))
    if block == "BandpassFilterBlock":
        return O.bandpass_filter(a[0], a[1], rate, cin, opt(2), opt(3, "hamming"))
    if block == "BandstopFilterBlock":
        return O.bandstop_filter(a[0], a[1], rate, cin, opt(2), opt(3, "hamming"))
    if block == "ComplexBandpassFilterBlock":
        return O.complex_bandpass_filter(a[0], a[1], rate, opt(2), opt(3, "hamming"))
    if block == "ComplexBandstopFilterBlock":
        return O.complex_bandstop_filter(a[0], a[1], rate, opt(2), opt(3, "hamming"))
    if block == "HilbertTransformBlock":
        return O.HilbertTransform(a[0], opt(1, "hamming"))
    if block == "FrequencyTranslatorBlock":
        return O.FrequencyTranslator(a[0], rate)
    if block == "FrequencyDiscriminatorBlock":
        return O.FrequencyDiscriminator(a[0])
    if block == "DownsamplerBlock":
        return O.Downsampler(a[0])
    if block == "FMDeemphasisFilterBlock":
        b, aa = O.fm_deemphasis_taps(a[0], rate)
        return O.IIRFilter(b, aa, cin)
    if block == "SinglepoleLowpassFilterBlock":
        b, aa = O.singlepole_lowpass_taps(a[0], rate)
        return O.IIRFilter(b, aa, cin)
    if block == "SinglepoleHighpassFilterBlock":
        b, aa = O.singlepole_highpass_taps(a[0], rate)
        return O.IIRFilter(b, aa, cin)
    if block == "IIRFilterBlock":
        return O.IIRFilter(np.asarray(a[0]), np.asarray(a[1]), cin)
    if block == "ComplexMagnitudeBlock":
        return O.Chain(O.complex_magnitude)
    if block == "ComplexToRealBlock":
        return O.Chain(O.complex_to_real)
    if block == "TunerBlock":
        o = opt(3, {}) or {}
        return O.tuner(a[0], a[1], a[2], rate, o.get("num_taps", 128), o.get("window", "hamming"))
    if block == "DecimatorBlock":
        o = opt(1, {}) or {}
        return O.decimator(a[0], cin, o.get("num_taps", 128), o.get("window", "hamming"))
    if block == "MultiplyConstantBlock":
        c = a[0]
        if isinstance(c, dict) and "complex" in c:
            c = complex(c["complex"][0], c["complex"][1])
        return O.MultiplyConstant(c)
    if block == "UpsamplerBlock":
        return O.Upsampler(a[0])
    if block == "InterpolatorBlock":
        o = opt(1, {}) or {}
        return O.interpolator(a[0], cin, o.get("num_taps", 128), o.get("window", "hamming"))
    if block == "RationalResamplerBlock":
        o = opt(2, {}) or {}
        return O.rational_resampler(a[0], a[1], cin, o.get("num_taps", 128), o.get("window", "hamming"))
    if block == "DelayBlock":
        return O.Delay(a[0])
    raise KeyError(block)


BINARY_OPS = {"MultiplyBlock": "multiply", "MultiplyConjugateBlock": "multiplyconjugate", "AddBlock": "add", "SubtractBlock": "subtract"}
