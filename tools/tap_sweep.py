#!/usr/bin/env python3
"""BASELINE config 3: FIR tap sweep 16/64/128/512/2048 taps x N complex samples on one GPU, HBM GB/s vs roofline.

    python tools/tap_sweep.py [--samples 1073741824] > gpurun_out/tap_sweep.json

Per point: complex in x real taps (crcf) and x complex taps (cccf); algorithmic bytes 16 B/sample; also the FP32
rate the direct form would need (4M / 8M flop per sample) so the bound in force is visible."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 30)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--taps", default="16,64,128,512,2048", help="comma-separated tap counts")
    args = ap.parse_args()
    import torch
    import luaradio_b200 as radio
    from luaradio_b200 import _lib
    import bench
    lib = _lib.require_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.lrb200_set_stream(ctypes.c_void_p(stream.cuda_stream)))
    n = args.samples
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x.data_ptr()), 0, n, 1))
    peak, src = bench.peaks()
    rows = []
    for M in [int(t) for t in args.taps.split(",")]:
        taps = np.array(radio.filter_utils.firwin_lowpass(M, 0.25), np.float32)
        for kind in ("crcf", "cccf"):
            if kind == "crcf":
                h = lib.lrb200_fir_create_crcf(taps.ctypes.data, M, 1, _lib.LRB200_DEVICE)
            else:
                ct = (taps * np.exp(1j * 0.5 * np.arange(M))).astype(np.complex64)
                h = lib.lrb200_fir_create_cccf(ct.ctypes.data, M, 1, _lib.LRB200_DEVICE)
            _lib.check_handle(h, "fir")
            no = ctypes.c_size_t()
            run = lambda: _lib.check(lib.lrb200_fir_execute(h, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(y.data_ptr()), ctypes.byref(no)))
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            torch.cuda.synchronize()
            probe = e0.elapsed_time(e1)
            steps = max(1, min(args.steps, int(2000.0 / max(probe, 1e-3))))
            e0.record(stream)
            for _ in range(steps):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            gbs = 16.0 * n / (ms * 1e-3) / 1e9
            direct_flops = (4 if kind == "crcf" else 8) * M
            rows.append({"taps": M, "kind": kind, "samples": n, "steps": steps, "ms": round(ms, 3),
                         "msamples_per_s": round(n / ms / 1e3, 1), "hbm_GBs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4),
                         "algorithm": "fft-overlap-save" if lib.lrb200_fir_get_algorithm(h) == _lib.FIR_FFT else "direct",
                         "direct_form_equiv_TFLOPs": round(direct_flops * n / (ms * 1e-3) / 1e12, 2)})
            lib.lrb200_fir_destroy(h)
            print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps({"peak_GBs": peak, "peak_source": src, "rows": rows}))


if __name__ == "__main__":
    main()
