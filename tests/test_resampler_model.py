"""The index arithmetic of the register-tiled polyphase resampler kernel (resample.cu: rs_poly_kernel), checked on the CPU
through its numpy model (tests/rs_kernel_model.py) against the oracle, for every (L, D) pair the library instantiates."""
import os
import re

import numpy as np
import pytest

from oracle import lr_oracle as O
from tests import rs_kernel_model as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "luaradio_b200", "csrc", "resample.cu")).read()
PAIRS = sorted({(int(a), int(b)) for a, b in re.findall(r"LRB_RS\((\d+), (\d+)\)", SRC)})


def test_model_and_kernel_agree_on_the_constants():
    assert int(re.search(r"constexpr int RS_C = (\d+);", SRC).group(1)) == K.C
    assert len(PAIRS) == 18 and all(K.rs_rb(L, D) > 0 for L, D in PAIRS)
    # periods per thread: the C++ rs_rb() restated in the model; spot-check against the source text
    assert "return L == 2 ? 8 : (L <= 4 ? 4 : (L == 5 ? 3 : (L <= 8 ? 2 : 0)));" in SRC
    assert [K.rs_rb(L, 1) for L in range(2, 9)] == [8, 4, 4, 3, 2, 2, 2]
    assert (K.rs_rb(2, 3), K.rs_rb(2, 5), K.rs_rb(3, 2), K.rs_rb(5, 4), K.rs_rb(7, 5)) == (4, 4, 4, 3, 2)
    # the circular window must hold the live window plus one refill
    for L, D in PAIRS:
        RI = K.rs_rb(L, D) * D
        WN = RI + K.C - 1
        WP = (WN + 2 * K.C - 1) // K.C * K.C
        assert WP >= WN + K.C


@pytest.mark.parametrize("L,D", PAIRS)
@pytest.mark.parametrize("cplx", [True, False])
def test_kernel_model_matches_the_oracle(L, D, cplx):
    rng = np.random.default_rng(L * 16 + D)
    for M in (128, 33, 7):
        n = 1500
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64) if cplx else rng.uniform(-1, 1, n).astype(np.float32)
        ref = (O.rational_resampler(L, D, cplx, M) if D > 1 else O.interpolator(L, cplx, M)).process(x)
        taps = O.f32_taps(O.firwin_lowpass(M, 1.0 / max(L, D), "hamming"))
        got = K.stream(x, taps, L, D, K.rs_rb(L, D), float(L), [700, 1, 0, 299, 500])
        assert len(got) == len(ref)
        assert np.max(np.abs(got - ref)) < 2e-6 * max(1.0, float(np.max(np.abs(ref))))
