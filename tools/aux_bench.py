#!/usr/bin/env python3
"""Device-side throughput of the kernels on the rows next to the hot path: file sample-format converters
at the graph boundaries and the resampling family.  One JSON object; achieved GB/s counts ALGORITHMIC bytes (input +
output of the block), against the measured HBM peak (MEASURED_PEAKS.json, see bench.peaks()).

    python tools/aux_bench.py [--samples N] [--steps K] > profiles/rNN_aux_bench.json
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 28)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only-resample", action="store_true", help="skip the file-format rows")
    args = ap.parse_args()
    import torch
    import luaradio_b200 as radio
    from luaradio_b200 import _lib
    import bench
    lib = _lib.require_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.lrb200_set_stream(ctypes.c_void_p(stream.cuda_stream)))
    D = _lib.LRB200_DEVICE
    n = args.samples
    peak, src = bench.peaks()
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x.data_ptr()), 0, n, 1))
    xs = x * 0.7                                   # inside [-1, 1] for the sink converters
    raw = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    y = torch.empty(n * 2 + 64, dtype=torch.complex64, device="cuda")
    rows = []

    def timed(name, make, in_ptr, n_in, out_ptr, bytes_per_in, note=""):
        g = _lib.check_handle(lib.lrb200_graph_create(), "graph")
        for h in make():
            _lib.check(lib.lrb200_graph_append(g, _lib.check_handle(h, name)), "append")
        _lib.check(lib.lrb200_graph_commit(g, 1), "commit")
        no = ctypes.c_size_t(0)

        def step():
            _lib.check(lib.lrb200_graph_reset(g), "reset")
            _lib.check(lib.lrb200_graph_execute_device(g, ctypes.c_void_p(in_ptr), n_in, ctypes.c_void_p(out_ptr), ctypes.byref(no)), name)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        gbs = bytes_per_in * n_in / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "graph": lib.lrb200_graph_describe(g).decode(), "input_samples": n_in, "output_samples": no.value,
                     "ms": round(ms, 4), "msamples_per_s": round(n_in / ms / 1e3, 1), "algorithmic_bytes_per_input_sample": bytes_per_in,
                     "hbm_GBs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4), "note": note})
        lib.lrb200_graph_destroy(g)

    # ---- file formats: fill `raw` with the sink's own output so the source converters read realistic bytes
    for fmt, b in (() if args.only_resample else (("u8", 1), ("s16le", 2), ("f32be", 4))):
        timed("iqsink(%s)" % fmt, lambda: [lib.lrb200_iqsink_create(fmt.encode(), D)], xs.data_ptr(), n, raw.data_ptr(), 8 + 2 * b)
        timed("iqconv(%s)" % fmt, lambda: [lib.lrb200_iqconv_create(fmt.encode(), D)], raw.data_ptr(), n, y.data_ptr(), 2 * b + 8)
    if not args.only_resample:
        timed("realsink(s16le)", lambda: [lib.lrb200_realsink_create(b"s16le", D)], xs.data_ptr(), 2 * n, raw.data_ptr(), 4 + 2,
              "WAVFileSink, 16 bits per sample")
    # ---- resampling family (LowpassFilterBlock(128, 1/L or min(1/L, 1/D), nyquist 1.0) taps)
    for L, Dn in ((2, 1), (3, 1), (4, 1), (8, 1), (2, 3), (3, 2), (5, 4), (2, 5), (160, 147)):
        m = n // (2 * L) if Dn == 1 else n // 2
        taps = np.array(radio.filter_utils.firwin_lowpass(128, min(1.0 / L, 1.0 / Dn)), np.float32)

        def make():
            hs = [lib.lrb200_mulconst_create(float(L), 0.0, 1, 0, D), lib.lrb200_upsample_create(L, 8, D),
                  lib.lrb200_fir_create_crcf(taps.ctypes.data, 128, 1, D)]
            if Dn > 1:
                hs.append(lib.lrb200_downsample_create(Dn, 8, D))
            return hs
        timed("interpolator x%d" % L if Dn == 1 else "rational resampler %d/%d" % (L, Dn), make, x.data_ptr(), m, y.data_ptr(),
              8 + 8.0 * L / Dn, "128 taps, complex")
    print(json.dumps({"peak_GBs": peak, "peak_source": src, "samples": n, "steps": args.steps, "rows": rows}))


if __name__ == "__main__":
    main()
