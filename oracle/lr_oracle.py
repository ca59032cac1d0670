"""CPU oracle for the LuaRadio sample-stream DSP hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's `process()` bodies (vsergeev/luaradio
v0.11.0 @ 7346868).  It exists to CHECK the CUDA path; nothing in `luaradio_b200/` imports it,
and only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may.  Every function cites the reference file:line it follows (paths relative to the
reference root).

Parity pinning: the reference stack itself (LuaJIT + VOLK v2.1.0 + liquid-dsp v1.3.2 + FFTW3f)
cannot run in the build container (none of them is installed, no network).  The oracle is
instead pinned against all 131 committed golden vectors of the reference's own busted specs for
this path (`tests/golden/*.npz`, extracted verbatim by `tests/golden/make_golden.py`), by
`tests/test_oracle_golden.py`, at the reference's own epsilons (1e-6 abs; 1e-5 for
translator/tuner).

Arithmetic convention: accumulate in float64, round once to float32 on output.  That is the
"infinitely precise then rounded" answer the reference's golden generators use
(`scipy.signal.lfilter` in float64 then `.astype(float32)`, e.g. tests/blocks/signal/
firfilter_spec.py:7-9) -- every backend of the reference (VOLK / liquid / pure Lua) is held to
it at 1e-6 absolute.
"""
import math

import numpy as np

C64 = np.complex64
F32 = np.float32

# ----------------------------------------------------------------------------------------------
# Tap design: radio/utilities/window_utils.lua, radio/utilities/filter_utils.lua (float64, host)
# ----------------------------------------------------------------------------------------------


def window(M, window_type="hamming", periodic=False):
    """radio/utilities/window_utils.lua:11-50 -- w[n], n=0..M-1, with M+1 in the formula if periodic."""
    Mf = (M + 1) if periodic else M
    n = np.arange(M, dtype=np.float64)
    if window_type == "rectangular":
        return np.ones(M)
    if window_type == "hamming":
        return 0.54 - 0.46 * np.cos((2 * np.pi * n) / (Mf - 1))
    if window_type == "hanning":
        return 0.5 - 0.5 * np.cos((2 * np.pi * n) / (Mf - 1))
    if window_type == "bartlett":
        return (2.0 / (Mf - 1)) * ((Mf - 1) / 2.0 - np.abs(n - (Mf - 1) / 2.0))
    if window_type == "blackman":
        return 0.42 - 0.5 * np.cos((2 * np.pi * n) / (Mf - 1)) + 0.08 * np.cos((4 * np.pi * n) / (Mf - 1))
    raise ValueError('Unsupported window "%s".' % window_type)


def _sinc_term(num_taps, cutoff):
    """sin(pi*fc*(n-c))/(pi*(n-c)) with the centre tap left to the caller (filter_utils.lua:21-33)."""
    c = (num_taps - 1) / 2.0
    n = np.arange(num_taps, dtype=np.float64) - c
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.sin(np.pi * cutoff * n) / (np.pi * n)
    return h, n


def fir_lowpass(num_taps, cutoff):
    """filter_utils.lua:21-33."""
    h, n = _sinc_term(num_taps, cutoff)
    h[n == 0] = cutoff
    return h


def fir_highpass(num_taps, cutoff):
    """filter_utils.lua:45-59 (odd taps only)."""
    assert num_taps % 2 == 1, "Number of taps must be odd."
    h, n = _sinc_term(num_taps, cutoff)
    h = -h
    h[n == 0] = 1 - cutoff
    return h


def fir_bandpass(num_taps, cutoffs):
    """filter_utils.lua:71-86."""
    assert num_taps % 2 == 1, "Number of taps must be odd."
    assert len(cutoffs) == 2
    h2, n = _sinc_term(num_taps, cutoffs[1])
    h1, _ = _sinc_term(num_taps, cutoffs[0])
    h = h2 - h1
    h[n == 0] = cutoffs[1] - cutoffs[0]
    return h


def fir_bandstop(num_taps, cutoffs):
    """filter_utils.lua:92-107."""
    assert num_taps % 2 == 1, "Number of taps must be odd."
    assert len(cutoffs) == 2
    h2, n = _sinc_term(num_taps, cutoffs[1])
    h1, _ = _sinc_term(num_taps, cutoffs[0])
    h = h1 - h2
    h[n == 0] = 1 - (cutoffs[1] - cutoffs[0])
    return h


def firwin(h, window_type, scale_freq):
    """filter_utils.lua:121-141 -- window, then scale to unit magnitude response at scale_freq."""
    h = h * window(len(h), window_type or "hamming")
    n = np.arange(len(h), dtype=np.float64)
    scale = np.sum(h * np.cos(np.pi * (n - (len(h) - 1) / 2.0) * scale_freq))
    return h / scale


def firwin_lowpass(num_taps, cutoff, window_type="hamming"):
    """filter_utils.lua:152-157."""
    return firwin(fir_lowpass(num_taps, cutoff), window_type, 0.0)


def firwin_highpass(num_taps, cutoff, window_type="hamming"):
    """filter_utils.lua:168-173."""
    return firwin(fir_highpass(num_taps, cutoff), window_type, 1.0)


def firwin_bandpass(num_taps, cutoffs, window_type="hamming"):
    """filter_utils.lua:184-189."""
    return firwin(fir_bandpass(num_taps, cutoffs), window_type, (cutoffs[0] + cutoffs[1]) / 2.0)


def firwin_bandstop(num_taps, cutoffs, window_type="hamming"):
    """filter_utils.lua:200-205."""
    return firwin(fir_bandstop(num_taps, cutoffs), window_type, 0.0)


def complex_firwin(h, center_freq, window_type, scale_freq):
    """filter_utils.lua:220-249 -- modulate to center_freq, window, scale at scale_freq (complex gain)."""
    n = np.arange(len(h), dtype=np.float64)
    hc = h * np.exp(1j * np.pi * center_freq * n)
    hc = hc * window(len(h), window_type or "hamming")
    scale = np.sum(hc * np.exp(-1j * np.pi * (n - (len(h) - 1) / 2.0) * scale_freq))
    return hc / scale


def firwin_complex_bandpass(num_taps, cutoffs, window_type="hamming"):
    """filter_utils.lua:261-266."""
    h = fir_lowpass(num_taps, (max(cutoffs) - min(cutoffs)) / 2.0)
    c = (cutoffs[0] + cutoffs[1]) / 2.0
    return complex_firwin(h, c, window_type, c)


def firwin_complex_bandstop(num_taps, cutoffs, window_type="hamming"):
    """filter_utils.lua:278-285."""
    h = fir_highpass(num_taps, (max(cutoffs) - min(cutoffs)) / 2.0)
    scale_freq = 1.0 if (cutoffs[0] < 0.0 and 0.0 < cutoffs[1]) else 0.0
    return complex_firwin(h, (cutoffs[0] + cutoffs[1]) / 2.0, window_type, scale_freq)


def fir_hilbert_transform(num_taps, window_type="hamming"):
    """filter_utils.lua:350-376 -- 2/(pi*n) at odd n about the centre, windowed, unscaled."""
    assert num_taps % 2 == 1, "Number of taps must be odd."
    ns = np.arange(num_taps) - (num_taps - 1) // 2
    h = np.zeros(num_taps)
    odd = (ns % 2) != 0
    h[odd] = 2.0 / (ns[odd] * np.pi)
    return h * window(num_taps, window_type or "hamming")


def f32_taps(h):
    """types/cstruct.lua:53-63 vector_from_array: float64 design -> float32 (or complex64) taps."""
    h = np.asarray(h)
    return h.astype(C64) if np.iscomplexobj(h) else h.astype(F32)


def singlepole_lowpass_taps(cutoff, rate):
    """singlepolelowpassfilter.lua:55-64 -- bilinear transform with pre-warp. Returns (b[2], a[2]) float32."""
    tau = 1.0 / (2 * math.pi * cutoff)
    tau = 1.0 / (2 * rate * math.tan(1.0 / (2 * rate * tau)))
    b = [1 / (1 + 2 * tau * rate), 1 / (1 + 2 * tau * rate)]
    a = [1.0, (1 - 2 * tau * rate) / (1 + 2 * tau * rate)]
    return np.array(b, F32), np.array(a, F32)


def singlepole_highpass_taps(cutoff, rate):
    """singlepolehighpassfilter.lua (same structure as lowpass, :initialize)."""
    tau = 1.0 / (2 * math.pi * cutoff)
    tau = 1.0 / (2 * rate * math.tan(1.0 / (2 * rate * tau)))
    b = [(2 * tau * rate) / (1 + 2 * tau * rate), -(2 * tau * rate) / (1 + 2 * tau * rate)]
    a = [1.0, (1 - 2 * tau * rate) / (1 + 2 * tau * rate)]
    return np.array(b, F32), np.array(a, F32)


def fm_deemphasis_taps(tau, rate):
    """fmdeemphasisfilter.lua:24-27 -- SinglepoleLowpass with cutoff 1/(2*pi*tau)."""
    return singlepole_lowpass_taps(1.0 / (2 * math.pi * tau), rate)


# ----------------------------------------------------------------------------------------------
# DFT conventions: radio/utilities/spectrum_utils.lua:206-244 (forward e^{-j..}), :456-504 (inverse, 1/N)
# ----------------------------------------------------------------------------------------------


def dft(x):
    return np.fft.fft(np.asarray(x, dtype=np.complex128)).astype(C64)


def idft(X):
    return np.fft.ifft(np.asarray(X, dtype=np.complex128)).astype(C64)


# ----------------------------------------------------------------------------------------------
# Streaming blocks.  Each mirrors initialize()/process() of the reference block, with the carried
# state the reference carries.  process() accepts any chunk length >= 0.
# ----------------------------------------------------------------------------------------------


class FIRFilter:
    """firfilter.lua:90-163 (dot-product form): y[n] = sum_k h[k] x[n-k], zero initial history
    (:100, Vector.new zero-fills), history of M-1 samples carried (:115-119), len(out)==len(in)."""

    def __init__(self, taps, complex_input):
        self.taps = np.asarray(taps)
        self.ctaps = np.iscomplexobj(self.taps)
        self.cin = bool(complex_input) or self.ctaps
        acc = np.complex128 if self.cin else np.float64
        self.h = self.taps.astype(np.complex128 if self.ctaps else np.float64)
        self.hist = np.zeros(len(self.taps) - 1, dtype=acc)
        self.out_dtype = C64 if self.cin else F32

    def process(self, x):
        x = np.asarray(x)
        M = len(self.h)
        ext = np.concatenate([self.hist, x.astype(self.hist.dtype)])
        if len(x) == 0:
            return np.zeros(0, self.out_dtype)
        # full convolution restricted to the n outputs aligned with the new samples
        y = np.convolve(ext, self.h)[M - 1:M - 1 + len(x)]
        if M > 1:
            self.hist = ext[len(ext) - (M - 1):]
        return y.astype(self.out_dtype)


class FIRFilterFFT:
    """firfilter.lua:320-398 overlap-save: N = 2^floor(log2(8M)), L = N-M+1; emits
    floor((fill+n)/L)*L samples per call (:362); residual input is held in input_block."""

    def __init__(self, taps, complex_input):
        self.taps = np.asarray(taps)
        self.cin = bool(complex_input) or np.iscomplexobj(self.taps)
        M = len(self.taps)
        self.N = 2 ** int(math.floor(math.log(8 * M) / math.log(2)))
        self.L = self.N - M + 1
        self.M = M
        ext = np.zeros(self.N, dtype=np.complex128)
        ext[:M] = self.taps
        self.taps_dft = np.fft.fft(ext)
        self.block = np.zeros(self.N, dtype=np.complex128)
        self.fill = 0
        self.out_dtype = C64 if self.cin else F32

    def process(self, x):
        x = np.asarray(x)
        outs = []
        i = 0
        while i < len(x):
            ln = min(len(x) - i, self.L - self.fill)
            self.block[self.M - 1 + self.fill:self.M - 1 + self.fill + ln] = x[i:i + ln]
            self.fill += ln
            i += ln
            if self.fill < self.L:
                break
            y = np.fft.ifft(np.fft.fft(self.block) * self.taps_dft)
            outs.append(y[self.M - 1:])
            self.block[:self.M - 1] = self.block[self.N - (self.M - 1):].copy()
            self.fill = 0
        y = np.concatenate(outs) if outs else np.zeros(0, np.complex128)
        return y.astype(C64) if self.cin else y.real.astype(F32)


class FrequencyTranslator:
    """frequencytranslator.lua:93-110 (pure-Lua path; VOLK rotator :26-53 is the same map up to
    drift): y[n] = x[n] * exp(j*omega*n), omega = 2*pi*offset/rate, phase carried across calls.
    The oracle evaluates the phase of the GLOBAL sample index exactly (Python big-int turns), i.e.
    the drift-free limit of the reference's recurrences and what its golden generator computes
    (tests/blocks/signal/frequencytranslator_spec.py:7)."""

    def __init__(self, offset, rate):
        self.turns = float(offset) / float(rate)   # cycles per sample
        self.n0 = 0

    def phasor(self, n0, n):
        # exact-ish: split turns*idx into integer and fractional parts in extended precision
        idx = np.arange(n0, n0 + n, dtype=np.float64)
        t = np.longdouble(self.turns) * idx.astype(np.longdouble)
        t = (t - np.floor(t)).astype(np.float64)
        return np.exp(2j * np.pi * t)

    def process(self, x):
        x = np.asarray(x)
        p = self.phasor(self.n0, len(x))
        self.n0 += len(x)
        return (x.astype(np.complex128) * p.astype(C64).astype(np.complex128)).astype(C64)


class FrequencyDiscriminator:
    """frequencydiscriminator.lua:25-38,68-88: y[n] = arg(x[n]*conj(x[n-1])) / (2*pi*k);
    prev_sample carried, initially 0 (:34)."""

    def __init__(self, modulation_index):
        self.gain = 2 * math.pi * modulation_index
        self.prev = np.complex128(0)

    def process(self, x):
        x = np.asarray(x).astype(np.complex128)
        if len(x) == 0:
            return np.zeros(0, F32)
        sh = np.concatenate([[self.prev], x[:-1]])
        t = x * np.conj(sh)
        self.prev = x[-1]
        return (np.arctan2(t.imag, t.real) / self.gain).astype(F32)


class Downsampler:
    """downsampler.lua:40-56: y[m] = x[index + m*D]; out length ceil((n-index)/D); index carried."""

    def __init__(self, factor):
        self.factor = int(factor)
        self.index = 0

    def process(self, x):
        x = np.asarray(x)
        n_out = max(0, -(-(len(x) - self.index) // self.factor))
        y = x[self.index:self.index + n_out * self.factor:self.factor].copy()
        assert len(y) == n_out
        self.index = self.index + n_out * self.factor - len(x)
        return y


class IIRFilter:
    """iirfilter.lua:113-179 (pure-Lua path; liquid :79-109 is the same recurrence):
    y[n] = (sum_j b[j] x[n-j] - sum_{j>=1} a[j] y[n-j]) / a[0], zero initial state."""

    def __init__(self, b_taps, a_taps, complex_input):
        self.b = np.asarray(b_taps, dtype=np.float64)
        self.a = np.asarray(a_taps, dtype=np.float64)
        assert len(self.a) >= 1
        self.cin = bool(complex_input)
        acc = np.complex128 if self.cin else np.float64
        self.xs = np.zeros(len(self.b), dtype=acc)
        self.ys = np.zeros(len(self.a) - 1, dtype=acc)

    def process(self, x):
        x = np.asarray(x)
        out = np.zeros(len(x), dtype=self.xs.dtype)
        b, a, xs, ys = self.b, self.a, self.xs, self.ys
        for i in range(len(x)):
            xs[1:] = xs[:-1].copy()
            xs[0] = x[i]
            acc = np.dot(xs, b) - (np.dot(ys, a[1:]) if len(ys) else 0.0)
            acc = acc / a[0]
            if len(ys):
                ys[1:] = ys[:-1].copy()
                ys[0] = acc
            out[i] = acc
        return out.astype(C64 if self.cin else F32)


class HilbertTransform:
    """hilberttransform.lua:27-37,132-167: out.real = x delayed by (M-1)/2, out.imag = (x * h_hilb)."""

    def __init__(self, num_taps, window_type="hamming"):
        assert num_taps % 2 == 1, "Number of taps must be odd"
        self.h = f32_taps(fir_hilbert_transform(num_taps, window_type)).astype(np.float64)
        self.hist = np.zeros(num_taps - 1)
        self.M = num_taps

    def process(self, x):
        x = np.asarray(x, dtype=np.float64)
        if len(x) == 0:
            return np.zeros(0, C64)
        M = self.M
        ext = np.concatenate([self.hist, x])
        im = np.convolve(ext, self.h)[M - 1:M - 1 + len(x)]
        re = ext[(M - 1) // 2:(M - 1) // 2 + len(x)]
        if M > 1:
            self.hist = ext[len(ext) - (M - 1):]
        return (re + 1j * im).astype(C64)


def complex_magnitude(x):
    """complexmagnitude.lua:28-36 / types/complexfloat32.lua:163-165: sqrt(re^2+im^2)."""
    x = np.asarray(x).astype(np.complex128)
    return np.sqrt(x.real * x.real + x.imag * x.imag).astype(F32)


def complex_to_real(x):
    """complextoreal.lua:27-35."""
    return np.asarray(x).real.astype(F32)


def multiply_conjugate(x, y):
    """multiplyconjugate.lua process: x * conj(y)."""
    return (np.asarray(x).astype(np.complex128) * np.conj(np.asarray(y).astype(np.complex128))).astype(C64)


def binary_op(op, x, y):
    """multiply.lua / add.lua / subtract.lua process bodies: element-wise on two equal-length vectors."""
    x, y = np.asarray(x), np.asarray(y)
    acc = np.complex128 if np.iscomplexobj(x) else np.float64
    a, b = x.astype(acc), y.astype(acc)
    r = {"multiply": a * b, "add": a + b, "subtract": a - b, "multiplyconjugate": a * np.conj(b)}[op]
    return r.astype(C64 if np.iscomplexobj(x) else F32)


class Delay:
    """delay.lua:26-60: out[n] = in[n - num_samples], zeros first; the last num_samples inputs are carried."""

    def __init__(self, num_samples):
        self.D = int(num_samples)
        self.state = None

    def process(self, x):
        x = np.asarray(x)
        if self.state is None:
            self.state = np.zeros(self.D, x.dtype)
        cat = np.concatenate([self.state, x])
        self.state = cat[len(cat) - self.D:]
        return cat[:len(x)].astype(x.dtype)


class PLL:
    """pll.lua:113-170 restated operation by operation: Lua numbers are float64; the VCO output, the phase-detector
    product (complexfloat32.lua:79-81) and the error (atan2f, :150-152) are float32 cells.  The reference has no pll_spec;
    this restatement is pinned instead against the reference's own pll.lua EXECUTED statement by statement in the test
    interpreter on float32-faithful sample cells (tests/test_lua_reference.py::
    test_pll_oracle_pinned_against_the_reference_pll_lua_executed: bit-identical over acquisition and lock, state carried
    across calls)."""

    def __init__(self, loop_bandwidth, frequency_min, frequency_max, multiplier, rate):
        bw = 2 * math.pi * (loop_bandwidth / rate)
        self.fmin = 2 * math.pi * (frequency_min / rate)
        self.fmax = 2 * math.pi * (frequency_max / rate)
        damping = math.sqrt(2) / 2
        bw = bw / (damping + 1 / (4 * damping))
        denom = 1 + 2 * damping * bw + bw * bw
        self.alpha = (4 * damping * bw) / denom
        self.beta = (4 * bw * bw) / denom
        self.mult = 1.0 if multiplier is None else multiplier
        self.phi = 0.0
        self.phim = 0.0
        self.freq = (self.fmin + self.fmax) / 2.0

    def process(self, x):
        x = np.asarray(x, C64)
        out = np.zeros(len(x), C64)
        err = np.zeros(len(x), F32)
        two_pi = 2 * math.pi
        phi, phim, freq = self.phi, self.phim, self.freq
        for i in range(len(x)):
            vr, vi = F32(math.cos(phi)), F32(math.sin(phi))
            out[i] = complex(F32(math.cos(phim)), F32(math.sin(phim)))
            xr, xi = float(x[i].real), float(x[i].imag)
            pr = F32(xr * float(vr) - xi * float(-vi))
            pi = F32(xr * float(-vi) + xi * float(vr))
            e = float(np.arctan2(pi, pr, dtype=F32))
            err[i] = e
            freq = freq + self.beta * e
            phi = phi + freq + self.alpha * e
            phim = phim + freq * self.mult + self.alpha * e
            freq = min(max(freq, self.fmin), self.fmax)
            if phi > two_pi:
                phi -= two_pi
            if phi < -two_pi:
                phi += two_pi
            if phim > two_pi:
                phim -= two_pi
            if phim < -two_pi:
                phim += two_pi
        self.phi, self.phim, self.freq = phi, phim, freq
        return out, err


def psd(samples, window_type="hamming", sample_rate=2, logarithmic=True):
    """spectrum_utils.lua:524-642 (PSD:compute, pure-Lua branch): periodic window as float32, DFT, |X_k|^2 / (rate *
    window energy), optionally 10*log10."""
    x = np.asarray(samples)
    n = len(x)
    w = np.array(window(n, window_type, True), F32)
    energy = float(np.sum(w.astype(np.float64) ** 2))
    if np.iscomplexobj(x):
        xw = (x.astype(np.complex128) * w.astype(np.float64)).astype(C64)
    else:
        xw = (x.astype(np.float64) * w.astype(np.float64)).astype(F32)
    X = np.fft.fft(xw.astype(np.complex128)).astype(C64)
    p = (X.real.astype(np.float64) ** 2 + X.imag.astype(np.float64) ** 2) / (sample_rate * energy)
    if logarithmic:
        with np.errstate(divide="ignore"):
            p = 10 * np.log10(p)
    return p.astype(F32)


# ----------------------------------------------------------------------------------------------
# Filter wrappers (taps designed from get_rate() in initialize()) and composites
# ----------------------------------------------------------------------------------------------


def lowpass_filter(num_taps, cutoff, rate, complex_input, nyquist=None, window_type="hamming"):
    """lowpassfilter.lua:32-50."""
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_lowpass(num_taps, cutoff / nyq, window_type)), complex_input)


def highpass_filter(num_taps, cutoff, rate, complex_input, nyquist=None, window_type="hamming"):
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_highpass(num_taps, cutoff / nyq, window_type)), complex_input)


def bandpass_filter(num_taps, cutoffs, rate, complex_input, nyquist=None, window_type="hamming"):
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_bandpass(num_taps, [c / nyq for c in cutoffs], window_type)), complex_input)


def bandstop_filter(num_taps, cutoffs, rate, complex_input, nyquist=None, window_type="hamming"):
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_bandstop(num_taps, [c / nyq for c in cutoffs], window_type)), complex_input)


def complex_bandpass_filter(num_taps, cutoffs, rate, nyquist=None, window_type="hamming"):
    """complexbandpassfilter.lua:35-54."""
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_complex_bandpass(num_taps, [c / nyq for c in cutoffs], window_type)), True)


def complex_bandstop_filter(num_taps, cutoffs, rate, nyquist=None, window_type="hamming"):
    nyq = nyquist or rate / 2.0
    return FIRFilter(f32_taps(firwin_complex_bandstop(num_taps, [c / nyq for c in cutoffs], window_type)), True)


class Chain:
    """Blocks connected output->input; rates follow get_rate() (downsampler.lua:36-38)."""

    def __init__(self, *blocks):
        self.blocks = blocks

    def process(self, x):
        for b in self.blocks:
            x = b.process(x) if hasattr(b, "process") else b(x)
        return x


def tuner(offset, bandwidth, decimation, rate, num_taps=128, window_type="hamming"):
    """composites/tuner.lua:32-48: Translator(offset) -> Lowpass(num_taps, bandwidth/2) -> Downsampler(D)."""
    return Chain(FrequencyTranslator(offset, rate),
                 lowpass_filter(num_taps, bandwidth / 2.0, rate, True, None, window_type),
                 Downsampler(decimation))


def decimator(decimation, complex_input, num_taps=128, window_type="hamming"):
    """composites/decimator.lua:28-42: Lowpass(num_taps, 1/D, nyquist=1.0) -> Downsampler(D)."""
    return Chain(lowpass_filter(num_taps, 1.0 / decimation, 2.0, complex_input, 1.0, window_type),
                 Downsampler(decimation))


def wbfm_mono_chain(rate=1102500.0, tune_offset=-250e3, bandwidth=200e3, tuner_decim=5,
                    deviation_index=1.25, af_taps=128, af_cutoff=15e3, tau=75e-6, af_decim=5):
    """examples/rtlsdr_wbfm_mono.lua:12-28 (= composites/wbfmmonodemodulator.lua:22-35 + tuner +
    downsampler): Tuner -> FrequencyDiscriminator(1.25) -> Lowpass(128, 15e3) -> FMDeemphasis(75e-6)
    -> Downsampler(5).  complex64 @ rate  ->  float32 @ rate/25."""
    r1 = rate / tuner_decim
    b, a = fm_deemphasis_taps(tau, r1)
    return Chain(tuner(tune_offset, bandwidth, tuner_decim, rate),
                 FrequencyDiscriminator(deviation_index),
                 lowpass_filter(af_taps, af_cutoff, r1, False),
                 IIRFilterFast(b, a, False),
                 Downsampler(af_decim))


class IIRFilterFast(IIRFilter):
    """Same recurrence as IIRFilter, evaluated by scipy.signal.lfilter (float64) with carried state,
    for long streams (the per-sample Python loop above is for golden-vector sizes)."""

    def __init__(self, b_taps, a_taps, complex_input):
        super().__init__(b_taps, a_taps, complex_input)
        import scipy.signal
        self._lfilter = scipy.signal.lfilter
        n = max(len(self.a), len(self.b)) - 1
        self.zi = np.zeros(n, dtype=self.xs.dtype)

    def process(self, x):
        x = np.asarray(x).astype(self.xs.dtype)
        if len(x) == 0:
            return np.zeros(0, C64 if self.cin else F32)
        if len(self.zi) == 0:
            y = self._lfilter(self.b, self.a, x)
        else:
            y, self.zi = self._lfilter(self.b, self.a, x, zi=self.zi)
        return y.astype(C64 if self.cin else F32)


# ----------------------------------------------------------------------------------------------
# Synthetic inputs: counter-based, so any window can be regenerated anywhere.
# The CUDA generator in luaradio_b200/csrc/synth.cu implements the same integer hash.
# ----------------------------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def splitmix64(z):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _u01_pair(idx, seed):
    with np.errstate(over="ignore"):
        h = splitmix64(idx.astype(np.uint64) * np.uint64(2) + (np.uint64(seed) << np.uint64(40)))
    lo = (h & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (h >> np.uint64(32)).astype(np.uint32)
    # 24-bit mantissa uniform in [-1, 1)
    a = ((lo >> np.uint32(8)).astype(np.float32) * F32(2.0 ** -23) - F32(1.0)).astype(F32)
    b = ((hi >> np.uint32(8)).astype(np.float32) * F32(2.0 ** -23) - F32(1.0)).astype(F32)
    return a, b


def synth_white_iq(n0, n, seed=1):
    """White IQ: I,Q ~ U[-1,1) float32 from a hash of the global sample index."""
    idx = np.arange(n0, n0 + n, dtype=np.uint64)
    a, b = _u01_pair(idx, seed)
    return (a.astype(np.float32) + 1j * b.astype(np.float32)).astype(C64)


def synth_fm_iq(n0, n, seed=1, rate=1102500.0, carrier=250e3, deviation=75e3, amp=0.5, noise=0.01):
    """Synthetic FM: x[n] = A*exp(j(2*pi*fc/fs*n + phi[n])) + w[n], phi from two audio tones
    (closed-form integral so any window is reproducible), w white U*noise.  float64 -> complex64."""
    idx = np.arange(n0, n0 + n, dtype=np.float64)
    t = idx / rate
    # integral of m(t) = 0.5 sin(2 pi f1 t) + 0.5 sin(2 pi f2 t)
    f1, f2 = 1e3, 7e3
    integ = 0.5 * (1 - np.cos(2 * np.pi * f1 * t)) / (2 * np.pi * f1) + 0.5 * (1 - np.cos(2 * np.pi * f2 * t)) / (2 * np.pi * f2)
    carrier_turns = np.longdouble(carrier / rate) * np.arange(n0, n0 + n, dtype=np.uint64).astype(np.longdouble)
    carrier_turns = (carrier_turns - np.floor(carrier_turns)).astype(np.float64)
    ph = 2 * np.pi * carrier_turns + 2 * np.pi * deviation * integ
    a, b = _u01_pair(np.arange(n0, n0 + n, dtype=np.uint64), seed)
    x = amp * np.exp(1j * ph) + noise * (a.astype(np.float64) + 1j * b.astype(np.float64))
    return x.astype(C64)


# ----------------------------------------------------------------------------------------------
# the source boundary.  IQFileSource sample formats -> ComplexFloat32
# ----------------------------------------------------------------------------------------------

IQ_FORMATS = {
    # radio/utilities/format_utils.lua:82-97: numpy dtype (file byte order), offset, scale
    "u8": ("u1", 127.5, 127.5), "s8": ("i1", 0.0, 127.5),
    "u16le": ("<u2", 32767.5, 32767.5), "u16be": (">u2", 32767.5, 32767.5),
    "s16le": ("<i2", 0.0, 32767.5), "s16be": (">i2", 0.0, 32767.5),
    "u32le": ("<u4", 2147483647.5, 2147483647.5), "u32be": (">u4", 2147483647.5, 2147483647.5),
    "s32le": ("<i4", 0.0, 2147483647.5), "s32be": (">i4", 0.0, 2147483647.5),
    "f32le": ("<f4", 0.0, 1.0), "f32be": (">f4", 0.0, 1.0),
    "f64le": ("<f8", 0.0, 1.0), "f64be": (">f8", 0.0, 1.0),
}


def iq_file_convert(raw, fmt):
    """radio/blocks/sources/iqfile.lua:82-116: interleaved I/Q of `fmt` -> complex64:
    byte swap if needed (:96-101), then (value - offset) / scale in double, stored as float32 (:105-108)."""
    dt, offset, scale = IQ_FORMATS[fmt]
    v = np.frombuffer(np.asarray(raw, dtype=np.uint8).tobytes(), dtype=np.dtype(dt)).astype(np.float64)
    y = ((v - offset) / scale).astype(F32)
    return (y[0::2] + 1j * y[1::2]).astype(C64)


def real_file_convert(raw, fmt):
    """radio/blocks/sources/realfile.lua:86-104: real samples of `fmt` -> float32 (same map as iq_file_convert)."""
    dt, offset, scale = IQ_FORMATS[fmt]
    v = np.frombuffer(np.asarray(raw, dtype=np.uint8).tobytes(), dtype=np.dtype(dt)).astype(np.float64)
    return ((v - offset) / scale).astype(F32)


def file_sink_convert(x, fmt):
    """IQFileSink:process (radio/blocks/sinks/iqfile.lua:66-80), RealFileSink / WAVFileSink:process
    (sinks/wavfile.lua:170-186): raw = x*scale + offset in double, stored into the C integer type (truncation toward
    zero; float formats: rounding to float32 / exact float64), in the file's byte order.  Complex input is written as
    interleaved I/Q.  Returns the file bytes (uint8).  Values outside [-1, 1] saturate (undefined in the reference)."""
    dt, offset, scale = IQ_FORMATS[fmt]
    x = np.asarray(x)
    if np.iscomplexobj(x):
        x = np.stack([x.real, x.imag], axis=1).reshape(-1)
    v = x.astype(F32).astype(np.float64) * scale + offset
    dt = np.dtype(dt)
    if dt.kind == "f":
        out = v.astype(dt)
    else:
        info = np.iinfo(dt)
        out = np.clip(np.trunc(v), info.min, info.max).astype(np.int64).astype(dt)
    return np.frombuffer(out.tobytes(), dtype=np.uint8).copy()


def wav_header(num_samples, num_channels, bits_per_sample, rate):
    """WAVFileSink headers (radio/blocks/sinks/wavfile.lua:135-160, sizes filled in by cleanup()): RIFF/WAVE, PCM
    'fmt ' chunk of 16 bytes, 'data' chunk; all little endian."""
    import struct
    data = num_samples * num_channels * (bits_per_sample // 8)
    return (b"RIFF" + struct.pack("<I", 36 + data) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, 1, num_channels, int(rate), int(rate) * num_channels * (bits_per_sample // 8),
                        num_channels * (bits_per_sample // 8), bits_per_sample) + b"data" + struct.pack("<I", data))


# ----------------------------------------------------------------------------------------------
# resampling family
# ----------------------------------------------------------------------------------------------
class MultiplyConstant:
    """multiplyconstant.lua:50-70: y = x * c (float32 arithmetic: complex x complex, complex x real, real x real)."""

    def __init__(self, constant):
        self.c = constant

    def process(self, x):
        x = np.asarray(x)
        if np.iscomplexobj(self.c) or isinstance(self.c, complex):
            return (x.astype(C64) * C64(self.c)).astype(C64)
        return (x * F32(self.c)).astype(x.dtype)


class Upsampler:
    """upsampler.lua:44-52: y[i*L] = x[i], zeros in between; len(out) = L*len(in); stateless."""

    def __init__(self, factor):
        self.factor = int(factor)

    def process(self, x):
        x = np.asarray(x)
        y = np.zeros(len(x) * self.factor, x.dtype)
        y[::self.factor] = x
        return y


def interpolator(interpolation, complex_input, num_taps=128, window_type="hamming"):
    """composites/interpolator.lua:31-41."""
    return Chain(MultiplyConstant(float(interpolation)), Upsampler(interpolation),
                 lowpass_filter(num_taps, 1.0 / interpolation, 2.0, complex_input, 1.0, window_type))


def rational_resampler(interpolation, decimation, complex_input, num_taps=128, window_type="hamming"):
    """composites/rationalresampler.lua:33-46."""
    cutoff = min(1.0 / interpolation, 1.0 / decimation)
    return Chain(MultiplyConstant(float(interpolation)), Upsampler(interpolation),
                 lowpass_filter(num_taps, cutoff, 2.0, complex_input, 1.0, window_type), Downsampler(decimation))
