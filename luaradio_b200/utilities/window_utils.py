"""Window functions (host, float64).  Mirrors radio/utilities/window_utils.lua:11-50."""
import math

_WINDOWS = {
    "rectangular": lambda n, M: 1.0,
    "hamming": lambda n, M: 0.54 - 0.46 * math.cos((2 * math.pi * n) / (M - 1)),
    "hanning": lambda n, M: 0.5 - 0.5 * math.cos((2 * math.pi * n) / (M - 1)),
    "bartlett": lambda n, M: (2.0 / (M - 1)) * ((M - 1) / 2.0 - abs(n - (M - 1) / 2.0)),
    "blackman": lambda n, M: 0.42 - 0.5 * math.cos((2 * math.pi * n) / (M - 1)) + 0.08 * math.cos((4 * math.pi * n) / (M - 1)),
}


def window(M, window_type, periodic=False):
    if window_type not in _WINDOWS:
        raise ValueError('Unsupported window "%s".' % (window_type,))
    f = _WINDOWS[window_type]
    Mf = (M + 1) if periodic else M
    return [f(n, Mf) for n in range(M)]
