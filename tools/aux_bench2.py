#!/usr/bin/env python3
"""Device-side throughput of the round-2 kernels next to the hot path: two-input element-wise
blocks, DelayBlock, PSD, the PLL (sequential and chunk-parallel), and the WBFM-stereo demodulator end to end through the
scheduler (one device DAG vs host-level scheduling).  Algorithmic bytes = inputs + outputs of the block.

    python tools/aux_bench2.py [log2_samples] > profiles/r02_aux_bench.json
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import luaradio_b200 as radio
    from luaradio_b200 import _lib
    import bench
    lib = _lib.require_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.lrb200_set_stream(ctypes.c_void_p(stream.cuda_stream)))
    D = _lib.LRB200_DEVICE
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
    peak, src = bench.peaks()
    a = torch.empty(n, dtype=torch.complex64, device="cuda")
    b = torch.empty(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    e = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(a.data_ptr()), 0, n, 1))
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(b.data_ptr()), 1 << 40, n, 2))
    rows = []

    def timed(name, h, ins, outs, count, bytes_per_sample, steps=10, note=""):
        _lib.check_handle(h, name)
        xs = (ctypes.c_void_p * len(ins))(*ins)
        ys = (ctypes.c_void_p * len(outs))(*outs)
        no = ctypes.c_size_t(0)

        def step():
            _lib.check(lib.lrb200_block_execute_multi(h, xs, len(ins), count, ys, len(outs), ctypes.byref(no)), name)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = bytes_per_sample * count / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "samples": count, "ms": round(ms, 4), "msamples_per_s": round(count / ms / 1e3, 1),
                     "algorithmic_bytes_per_sample": bytes_per_sample, "hbm_GBs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4), "note": note})
        lib.lrb200_block_destroy(h)

    for op in ("multiply", "multiplyconjugate", "add", "subtract"):
        timed(op + " (complex)", lib.lrb200_binary_create(op.encode(), 1, D), [a.data_ptr(), b.data_ptr()], [y.data_ptr()], n, 24)
    timed("add (real)", lib.lrb200_binary_create(b"add", 0, D), [a.data_ptr(), b.data_ptr()], [y.data_ptr()], 2 * n, 12)
    timed("delay(129) complex", lib.lrb200_delay_create(129, 8, D), [a.data_ptr()], [y.data_ptr()], n, 16)
    win = np.array(radio.window_utils.window(1024, "hamming", True), np.float32)
    timed("psd(1024, hamming, log) complex", lib.lrb200_psd_create(1024, win.ctypes.data, 2.0 * float(np.sum(win.astype(np.float64) ** 2)), 1, 1, D),
          [a.data_ptr()], [e.data_ptr()], n, 12, note="one CTA per 1024-sample frame, radix-2 in shared memory")
    # PLL on a clean 19 kHz pilot at 220.5 kHz
    m = 1 << 22
    t = torch.arange(m, device="cuda", dtype=torch.float64) / 220500.0
    pilot = torch.polar(torch.full((m,), 0.8, device="cuda", dtype=torch.float64), 2 * np.pi * 19000.3 * t).to(torch.complex64)
    for mode, name in ((0, "pll sequential (exact)"), (1, "pll chunk-parallel (locked loop)")):
        h = lib.lrb200_pll_create(100.0, 19e3 - 50, 19e3 + 50, 2.0, 220500.0, D)
        _lib.check(lib.lrb200_pll_set_mode(h, mode))
        timed(name, h, [pilot.data_ptr()], [y.data_ptr(), e.data_ptr()], m if mode else 1 << 18, 20, steps=3,
              note="PLLBlock(100, 19 kHz +- 50, x2) at 220.5 kHz; the reference's Lua loop: 5 MS/s on an i5")
    # WBFM stereo demodulator through the scheduler, host vectors of 1 Mi samples
    from oracle import lr_oracle as O
    ns = 1 << 24
    x = O.synth_fm_iq(0, ns, 1, 220500.0, 0.0, 75e3, 0.9, 0.001)
    for dag in (True, False):
        src_b, sl, sr = radio.ArraySource(x, 220500.0, 1 << 20), radio.ArraySink(), radio.ArraySink()
        sl.process = lambda v: None
        sr.process = lambda v: None
        demod = radio.WBFMStereoDemodulator()
        for blk in demod._blocks:
            if isinstance(blk, radio.PLLBlock):
                blk.parallel = True
        top = radio.CompositeBlock()
        top.connect(src_b, demod)
        top.connect(demod, "left", sl, "in")
        top.connect(demod, "right", sr, "in")
        t0 = time.perf_counter()
        top.run(device_dag=dag)
        el = time.perf_counter() - t0
        rows.append({"kernel": "WBFMStereoDemodulator via CompositeBlock.run(device_dag=%s)" % dag, "samples": ns, "ms": round(el * 1e3, 1),
                     "msamples_per_s": round(ns / el / 1e6, 2), "graph": top.describe_gpu_graph()[:200],
                     "note": "host vectors of 1 Mi samples, includes graph construction; PLL chunk-parallel"})
    print(json.dumps({"peak_GBs": peak, "peak_source": src, "rows": rows}))


if __name__ == "__main__":
    main()
