import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import luaradio_b200 as radio
from luaradio_b200.types import Float32, Vector
from oracle import lr_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
rng = np.random.default_rng(8)
x = rng.uniform(-1, 1, n).astype(np.float32)
taps = O.fm_deemphasis_taps(75e-6, 220500.0)
ref = O.IIRFilterFast(taps[0], taps[1], False).process(x)
for rep in range(3):
    b = radio.FMDeemphasisFilterBlock(75e-6); b.get_rate = lambda: 220500.0; b.differentiate([Float32]); b.initialize()
    got = np.array(b.process(Vector.cast(x)).data, copy=True)
    err = got - ref
    bad = np.nonzero(np.abs(err) > 1e-5)[0]
    print("rep", rep, "n", n, "max err %.3g" % np.abs(err).max(), "first bad", bad[:3], "count", len(bad))
    if len(bad):
        i0 = bad[0]
        print(" err[%d..]:" % (i0 - 10), np.array2string(err[i0 - 10:i0 + 12], precision=4))
        t = (i0 // 2048) * 2048
        print(" err at tile start %d:" % t, np.array2string(err[t:t + 20], precision=4))
        # ratio structure: does err decay by c per sample within a thread?
        print(" err/err_prev:", np.array2string(err[i0 + 1:i0 + 9] / err[i0:i0 + 8], precision=4))
