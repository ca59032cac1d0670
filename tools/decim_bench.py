#!/usr/bin/env python3
"""Decimating FIR shapes (DecimatorBlock / FIRFilterBlock + Downsampler) on device-resident white IQ: time per call and
achieved GB/s (8 + 8/D B per complex input sample, 4 + 4/D real) for the automatic, forced-direct and forced-FFT
algorithms.  python tools/decim_bench.py [log2_samples]  -> one JSON line (profiles/r02_decim_shapes.json)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from luaradio_b200 import _lib
    import luaradio_b200 as radio
    lib = _lib.require_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.lrb200_set_stream(ctypes.c_void_p(stream.cuda_stream)))
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x.data_ptr()), 0, n, 1))
    peak = 6486.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"samples": n, "peak_gbs": peak, "shapes": []}
    for kind in ("crcf", "rrrf", "cccf"):
        for D, M in ((2, 16), (2, 128), (3, 33), (4, 128), (5, 64), (7, 200), (8, 128), (10, 128), (16, 256), (25, 513)):
            taps = np.array(radio.filter_utils.firwin_lowpass(M, 1.0 / D), np.float32)
            if kind == "cccf":
                taps = (taps * np.exp(1j * 0.3 * np.arange(M))).astype(np.complex64)
            create = getattr(lib, "lrb200_fir_create_" + kind)
            nn = n * 2 if kind == "rrrf" else n           # the complex buffer holds 2n floats
            row = {"kind": kind, "D": D, "M": M}
            for algo, name in ((_lib.FIR_AUTO, "auto"), (_lib.FIR_DIRECT, "direct"), (_lib.FIR_FFT, "fft")):
                h = _lib.check_handle(create(taps.ctypes.data, M, D, _lib.LRB200_DEVICE), "fir")
                _lib.check(lib.lrb200_fir_set_algorithm(h, algo))
                no = ctypes.c_size_t(0)

                def run():
                    _lib.check(lib.lrb200_fir_execute(h, ctypes.c_void_p(x.data_ptr()), nn, ctypes.c_void_p(y.data_ptr()), ctypes.byref(no)))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(5):
                    run()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                esz = 4 if kind == "rrrf" else 8
                gbs = (esz + esz / D) * nn / (ms * 1e-3) / 1e9
                row[name] = {"ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3),
                             "algo": "fft" if lib.lrb200_fir_get_algorithm(h) == _lib.FIR_FFT else "direct"}
                lib.lrb200_fir_destroy(h)
            out["shapes"].append(row)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
