import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import luaradio_b200 as radio
from luaradio_b200.types import Float32, ComplexFloat32, Vector
from oracle import lr_oracle as O

rng = np.random.default_rng(8)
n = 400000
x = rng.uniform(-1, 1, n).astype(np.float32)
for name, cls, args, rate, taps in (
    ("deemph", radio.FMDeemphasisFilterBlock, [75e-6], 220500.0, O.fm_deemphasis_taps(75e-6, 220500.0)),
    ("lp10", radio.SinglepoleLowpassFilterBlock, [10.0], 1e6, O.singlepole_lowpass_taps(10.0, 1e6)),
    ("hp1k", radio.SinglepoleHighpassFilterBlock, [1e3], 48e3, O.singlepole_highpass_taps(1e3, 48e3)),
):
    ref = O.IIRFilterFast(taps[0], taps[1], False).process(x)
    for mode in ("whole", "chunks"):
        b = cls(*args); b.get_rate = lambda: rate; b.differentiate([Float32]); b.initialize()
        if mode == "whole":
            got = np.array(b.process(Vector.cast(x)).data, copy=True)
        else:
            outs = []
            for a in range(0, n, 50000):
                outs.append(np.array(b.process(Vector.cast(x[a:a+50000])).data, copy=True))
            got = np.concatenate(outs)
        err = np.abs(got - ref)
        bad = np.nonzero(err > 1e-5 * max(1, np.max(np.abs(ref))))[0]
        print(name, mode, "taps", taps, "max err %.3g" % err.max(), "first bad", bad[:5], "count", len(bad), "ref max %.3g" % np.max(np.abs(ref)))
