#!/usr/bin/env python3
"""bench.py -- WBFM-mono chain throughput on B200 (BASELINE.json metric), roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload = "wbfm_mono_256Mi", BASELINE.json configs[2]): the full
examples/rtlsdr_wbfm_mono.lua flow graph  Tuner(-250e3, 200e3, 5) -> FrequencyDiscriminator(1.25) ->
Lowpass(128, 15e3) -> FMDeemphasis(75e-6) -> Downsampler(5)  on 256 Mi complex-float32 samples of synthetic
FM IQ at 1.1025 MS/s per GPU.  A "step" is one pass of the chain over one rank's 256 Mi-sample chunk.

  value  : whole-job M input-samples/s, inputs resident in HBM (device-generated synthetic IQ), CUDA events,
           max over ranks.
  e2e    : same metric through the C ABI with HOST buffers (lrb200_graph_execute: pinned host in, host out,
           H2D + kernels + D2H pipelined inside the timed call).
  N > 1  : the stream is sharded by time chunk, one 256 Mi-sample chunk per rank (weak scaling).  Each step
           rank r sends the last HALO input samples of its chunk to rank r+1 (NCCL P2P over NVLink) and every
           rank r > 0 runs from a cold state HALO samples early, discarding the corresponding outputs.
  --impl reference : the C restatement of the reference's VOLK code path (oracle/lr_oracle.c) on the host
           cores -- the reference stack (LuaJIT/VOLK/liquid/FFTW) is not installable here.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 1102500.0
TUNE_OFFSET = -250e3
BANDWIDTH = 200e3
HALO = 4000            # input samples, multiple of 25: 127 (tuner FIR) + 5*(1 + 127 + 640 IIR warm-up) = 3967
ALG_BYTES_CHAIN = 8.0 + 4.0 / 25.0          # fused minimum per input sample
ALG_BYTES_TUNER = 8.0 + 8.0 / 5.0           # tuner kernel: read 8 B, write 8/5 B per input sample
ALG_BYTES_TUNER_DISC = 8.0 + 4.0 / 5.0      # tuner with the discriminator fused: the 1/5-rate output is float32
ALG_BYTES_FIR = 16.0                        # plain complex FIR: 8 in + 8 out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def chain_taps():
    """Tap design exactly as the blocks' initialize() does it (product-side filter_utils)."""
    import luaradio_b200 as radio
    fu = radio.filter_utils
    t1 = np.array(fu.firwin_lowpass(128, (BANDWIDTH / 2) / (RATE / 2)), np.float32)
    r1 = RATE / 5
    t2 = np.array(fu.firwin_lowpass(128, 15e3 / (r1 / 2)), np.float32)
    tau = 75e-6
    cutoff = 1 / (2 * math.pi * tau)
    tt = 1 / (2 * math.pi * cutoff)
    tt = 1 / (2 * r1 * math.tan(1 / (2 * r1 * tt)))
    b = np.array([1 / (1 + 2 * tt * r1)] * 2, np.float32)
    a = np.array([1.0, (1 - 2 * tt * r1) / (1 + 2 * tt * r1)], np.float32)
    return t1, t2, b, a


def build_chain_graph(lib, _lib, iq_format=None):
    """The rtlsdr_wbfm_mono.lua chain as C-ABI blocks appended to a GPU flow graph (what CompositeBlock.run builds).
    `iq_format`: prepend the source-boundary sample-format converter (raw file/dongle bytes in)."""
    t1, t2, b, a = chain_taps()
    D = _lib.LRB200_DEVICE
    g = _lib.check_handle(lib.lrb200_graph_create(), "graph")
    blocks = [lib.lrb200_iqconv_create(iq_format.encode(), D)] if iq_format else []
    blocks += [
        lib.lrb200_rotator_create(TUNE_OFFSET / RATE, D),
        lib.lrb200_fir_create_crcf(t1.ctypes.data, 128, 1, D),
        lib.lrb200_downsample_create(5, 8, D),
        lib.lrb200_discrim_create(2 * math.pi * 1.25, D),
        lib.lrb200_fir_create_rrrf(t2.ctypes.data, 128, 1, D),
        lib.lrb200_iir_create_rrrf(b.ctypes.data, 2, a.ctypes.data, 2, D),
        lib.lrb200_downsample_create(5, 4, D),
    ]
    for h in blocks:
        _lib.check_handle(h, "block")
        _lib.check(lib.lrb200_graph_append(g, h), "graph_append")
    _lib.check(lib.lrb200_graph_commit(g, 1), "graph_commit")
    return g


def run_b200(args):
    import torch
    import torch.distributed as dist
    from luaradio_b200 import _lib, sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and rank == 0:
        sys.stderr.write("note: WORLD_SIZE=%d, --gpus=%d; using WORLD_SIZE\n" % (world, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.require_device(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.check(lib.lrb200_set_stream(ctypes.c_void_p(stream.cuda_stream)), "set_stream")

    n = (args.samples // 25) * 25            # chunk boundaries on multiples of 25 keep both decimators aligned
    start = rank * n                          # global index of this rank's first sample
    lead = HALO if rank > 0 else 0
    x = torch.empty(n + HALO, dtype=torch.complex64, device="cuda")
    xp = x.data_ptr()
    # this rank's chunk lives at x[HALO:], the halo from the left neighbour is received into x[:HALO]
    _lib.check(lib.lrb200_synth_fm_iq(ctypes.c_void_p(xp + HALO * 8), start, n, 1, RATE, 250e3, 75e3, 0.5, 0.01), "synth")
    g = build_chain_graph(lib, _lib)
    n_out_max = lib.lrb200_graph_max_output(g, n + HALO)
    y = torch.empty(n_out_max + 16, dtype=torch.float32, device="cuda")
    desc = lib.lrb200_graph_describe(g).decode()
    nstages = lib.lrb200_graph_num_stages(g)
    n_out = ctypes.c_size_t(0)

    def step():
        if world > 1:
            sharding.exchange_halo(dist, x[HALO:HALO + n], x[0:HALO], rank, world, HALO)
        _lib.check(lib.lrb200_graph_reset(g), "reset")
        _lib.check(lib.lrb200_graph_seek(g, start - lead), "seek")
        _lib.check(lib.lrb200_graph_execute_device(g, ctypes.c_void_p(xp + (HALO - lead) * 8), n + lead,
                                                   ctypes.c_void_p(y.data_ptr()), ctypes.byref(n_out)), "execute")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _lib.check(lib.lrb200_graph_set_timing(g, 1), "timing")
    launches0 = lib.lrb200_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = lib.lrb200_launch_count() - launches0
    stage_ms = []
    for k in range(nstages):
        cnt = ctypes.c_int(0)
        t = lib.lrb200_graph_stage_time_ms(g, k, ctypes.byref(cnt))
        stage_ms.append((lib.lrb200_graph_stage_name(g, k).decode(), t / max(cnt.value, 1)))
    _lib.check(lib.lrb200_graph_set_timing(g, 0), "timing")
    tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6          # M input-samples/s, whole job

    if args.profile:
        if rank == 0:
            emit(json.dumps({"profile_run": True, "ms_per_step": ms_step, "stages_ms": dict(stage_ms)}))
            bench_fir128(lib, _lib, torch, stream, args)
        return
    # ---- e2e: HOST buffers through the C ABI (pinned in, host out), H2D/D2H inside the timed call
    hin = lib.lrb200_host_alloc((n + lead) * 8)
    hout = lib.lrb200_host_alloc((n_out_max + 16) * 4)
    assert hin and hout, _lib.last_error()
    _lib.check(lib.lrb200_memcpy_d2h(hin, ctypes.c_void_p(xp + (HALO - lead) * 8), (n + lead) * 8), "d2h")
    _lib.check(lib.lrb200_sync(), "sync")
    e2e_steps = max(2, min(args.steps, 8))

    def e2e_step():
        _lib.check(lib.lrb200_graph_reset(g), "reset")
        _lib.check(lib.lrb200_graph_seek(g, start - lead), "seek")
        _lib.check(lib.lrb200_graph_execute(g, hin, n + lead, hout, ctypes.byref(n_out)), "graph_execute(host)")

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    te = torch.tensor([(t1 - t0) / e2e_steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n / float(te.item()) / 1e6
    n_out_step = int(n_out.value)

    # ---- the same end-to-end call fed the RTL-SDR's native u8 I/Q bytes (2 B/sample over PCIe, converted on the
    # device by the graph's first stage) -- extra information, N == 1 only; the headline e2e stays the f32 boundary
    e2e_u8 = None
    if world == 1:
        g8 = build_chain_graph(lib, _lib, "u8")
        h8 = lib.lrb200_host_alloc(n * 2)
        assert h8, _lib.last_error()
        xr = torch.view_as_real(x[HALO:HALO + n])
        CH = 1 << 24
        for o in range(0, n, CH):      # quantise the synthetic samples to u8 on the device, park them in pinned host memory
            q = (xr[o:o + CH] * 127.5 + 127.5).round_().clamp_(0, 255).to(torch.uint8).contiguous()
            _lib.check(lib.lrb200_memcpy_d2h(ctypes.c_void_p(h8 + o * 2), ctypes.c_void_p(q.data_ptr()), q.numel()), "d2h")
            _lib.check(lib.lrb200_sync(), "sync")
            torch.cuda.synchronize()
        del q, xr

        def e2e8_step():
            _lib.check(lib.lrb200_graph_reset(g8), "reset")
            _lib.check(lib.lrb200_graph_execute(g8, h8, n, hout, ctypes.byref(n_out)), "graph_execute(host,u8)")

        e2e8_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e8_step()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e2e_u8 = {"value": round(n / ((t1 - t0) / e2e_steps) / 1e6, 1), "unit": "Msamples/s", "h2d_bytes_per_step": n * 2,
                  "d2h_bytes_per_step": int(n_out.value) * 4, "graph": lib.lrb200_graph_describe(g8).decode(),
                  "note": "input = raw u8 I/Q bytes (IQFileSource/RtlSdrSource format), converted by the graph's first stage"}
        lib.lrb200_host_free(h8)
        lib.lrb200_graph_destroy(g8)

    result = None
    if rank == 0:
        peak, peak_src = peaks()
        dom = max(stage_ms, key=lambda kv: kv[1])
        dom_bytes = (ALG_BYTES_TUNER_DISC if dom[0].startswith("tuner+discrim") else
                     ALG_BYTES_TUNER if dom[0].startswith("tuner") else ALG_BYTES_FIR)
        achieved = dom_bytes * (n + lead) / (dom[1] * 1e-3) / 1e9 if dom[1] > 0 else 0.0
        traffic = None
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this workload size,
            # from the committed `ncu --set full` capture (profiles/); GB
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom[0])
            if traffic is not None and n != 268435450:
                traffic = None
        except Exception:
            pass
        result = {
            "metric": "WBFM-mono chain Msamples/s (input samples) @N B200; FIR-128 HBM GB/s vs roofline",
            "value": round(value, 1), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "wbfm_mono_256Mi" if n == 268435450 else "wbfm_mono_%d" % n,
                       "samples_per_gpu": n, "sample_rate_hz": RATE, "chain": "Tuner(-250e3,200e3,5)>FreqDiscriminator(1.25)>Lowpass(128,15e3)>FMDeemphasis(75e-6)>Downsampler(5)",
                       "graph": desc, "halo_samples": HALO if world > 1 else 0, "sharding": "time-chunk, NCCL P2P halo" if world > 1 else "single GPU",
                       "l2": "inputs (2 GiB/step) larger than L2, no flush"},
            "gpu_launches": int(launches),
            "stages_ms": {k: round(v, 4) for k, v in stage_ms},
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "peak_source": peak_src, "traffic": traffic,
                         "algorithmic_bytes_per_input_sample": dom_bytes,
                         "chain_fused_min_GBs": round(ALG_BYTES_CHAIN * world * n / (ms_step * 1e-3) / 1e9, 1)},
            "e2e": {"value": round(e2e_value, 1), "unit": "Msamples/s", "h2d_bytes_per_step": (n + lead) * 8,
                    "d2h_bytes_per_step": n_out_step * 4, "api": "lrb200_graph_execute (host pointers)", "steps": e2e_steps},
            "clocks": clocks,
        }
        if e2e_u8:
            result["e2e_u8"] = e2e_u8
    lib.lrb200_host_free(hin)
    lib.lrb200_host_free(hout)
    lib.lrb200_graph_destroy(g)
    del x, y

    # ---- FIR-128 alone (second half of the BASELINE metric) and the CPU baseline: rank 0, N == 1 only
    if rank == 0 and world == 1:
        result["fir128"] = bench_fir128(lib, _lib, torch, stream, args)
        result["cpu_baseline"] = cpu_baseline(lib, _lib, args)
    if rank == 0:
        emit(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_fir128(lib, _lib, torch, stream, args):
    """LowpassFilterBlock(128, 15e3) @ 220.5 kHz taps on white IQ resident in HBM: GB/s = 16 B * n / t."""
    import luaradio_b200 as radio
    n = args.fir_samples
    taps = np.array(radio.filter_utils.firwin_lowpass(128, 15e3 / 110250.0), np.float32)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    y = torch.empty(n, dtype=torch.complex64, device="cuda")
    _lib.check(lib.lrb200_synth_white_iq(ctypes.c_void_p(x.data_ptr()), 0, n, 1), "synth")
    out = {}
    peak, _ = peaks()
    for algo, name in ((_lib.FIR_AUTO, "auto"),):
        h = _lib.check_handle(lib.lrb200_fir_create_crcf(taps.ctypes.data, 128, 1, _lib.LRB200_DEVICE), "fir")
        _lib.check(lib.lrb200_fir_set_algorithm(h, algo), "algo")
        no = ctypes.c_size_t(0)
        steps = max(3, min(args.steps, 20))
        for _ in range(3):
            _lib.check(lib.lrb200_fir_execute(h, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(y.data_ptr()), ctypes.byref(no)), "fir")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            _lib.check(lib.lrb200_fir_execute(h, ctypes.c_void_p(x.data_ptr()), n, ctypes.c_void_p(y.data_ptr()), ctypes.byref(no)), "fir")
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = ALG_BYTES_FIR * n / (ms * 1e-3) / 1e9
        out = {"workload": "LowpassFilterBlock(128,15e3) crcf, %d complex samples" % n, "ms": round(ms, 4),
               "msamples_per_s": round(n / (ms * 1e-3) / 1e6, 1), "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
               "frac": round(gbs / peak, 4), "algorithm": "fft-overlap-save" if lib.lrb200_fir_get_algorithm(h) == _lib.FIR_FFT else "direct"}
        lib.lrb200_fir_destroy(h)
    return out


def host_fm_samples(lib, _lib, n):
    d = lib.lrb200_malloc(n * 8)
    host = np.empty(n, np.complex64)
    _lib.check(lib.lrb200_synth_fm_iq(d, 0, n, 1, RATE, 250e3, 75e3, 0.5, 0.01), "synth")
    _lib.check(lib.lrb200_memcpy_d2h(host.ctypes.data, d, n * 8), "d2h")
    _lib.check(lib.lrb200_sync(), "sync")
    lib.lrb200_free(d)
    return host


def time_cpu_chain(olib, x, threads, min_seconds=10.0, max_reps=50):
    t1, t2, b, a = chain_taps()
    out = np.zeros(len(x) // 25 + 8, np.float32)
    omega = 2 * math.pi * (TUNE_OFFSET / RATE)

    def once():
        return olib.lro_wbfm_chain(x.ctypes.data, len(x), 0, t1.ctypes.data, 128, 5, omega, 2 * math.pi * 1.25,
                                   t2.ctypes.data, 128, b.ctypes.data, a.ctypes.data, 5, out.ctypes.data, threads, 25 * 1024)
    once()
    reps, t0 = 0, time.perf_counter()
    while True:
        once()
        reps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or reps >= max_reps:
            break
    return len(x) * reps / el / 1e6, el, reps


def cpu_baseline(lib, _lib, args):
    """The C restatement of the reference's VOLK path (oracle/lr_oracle.c) on this box's host cores."""
    from oracle import cbuild
    olib = cbuild.load(native=True)
    cores = olib.lro_max_threads()
    n = args.cpu_samples
    x = host_fm_samples(lib, _lib, n)
    v, el, reps = time_cpu_chain(olib, x, cores, min_seconds=10.0)
    return {"value": round(v, 1), "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": "%d x %d-sample synthetic FM IQ (same generator), whole chain, %d threads time-chunked, %.1f s" % (reps, n, cores, el),
            "note": "C port of the reference's VOLK dot-product path (firfilter.lua:129-145 etc.); LuaJIT/VOLK/liquid/FFTW are not installable here"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from luaradio_b200 import _lib
    from oracle import cbuild
    olib = cbuild.load(native=True)
    cores = olib.lro_max_threads()
    n = args.cpu_samples
    try:
        lib = _lib.require_device(0)
        x = host_fm_samples(lib, _lib, n)
    except Exception:
        from oracle import lr_oracle as O
        x = np.concatenate([O.synth_fm_iq(i, min(1 << 22, n - i)) for i in range(0, n, 1 << 22)])
    t1, t2, b, a = chain_taps()
    out = np.zeros(n // 25 + 8, np.float32)
    omega = 2 * math.pi * (TUNE_OFFSET / RATE)

    def once():
        olib.lro_wbfm_chain(x.ctypes.data, n, 0, t1.ctypes.data, 128, 5, omega, 2 * math.pi * 1.25, t2.ctypes.data, 128,
                            b.ctypes.data, a.ctypes.data, 5, out.ctypes.data, cores, 25 * 1024)
    steps = max(1, min(args.steps, 20))
    for _ in range(max(1, min(args.warmup, 3))):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    el = time.perf_counter() - t0
    v = n * steps / el / 1e6
    emit(json.dumps({
        "impl": "reference", "metric": "WBFM-mono chain Msamples/s (input samples) @N B200; FIR-128 HBM GB/s vs roofline",
        "value": round(v, 1), "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 3)),
        "ms_per_step": round(el / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wbfm_mono_256Mi", "bounded_sample": n, "sample_rate_hz": RATE},
        "cpu_baseline": {"value": round(v, 1), "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": "%d-sample synthetic FM IQ per step, whole chain, %d threads time-chunked" % (n, cores)},
        "e2e": {"value": round(v, 1), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


_RESULT_OUT = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; see main()."""
    out = _RESULT_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    # Libraries write to fd 1 behind Python's back (NCCL prints "NCCL version ..." there at init): keep a private copy of
    # the real stdout for the result line and point fd 1 at stderr for everything else.
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=268435450, help="input samples per GPU per step (multiple of 25; 256 Mi)")
    ap.add_argument("--fir-samples", type=int, default=1 << 28)
    ap.add_argument("--cpu-samples", type=int, default=1 << 26)
    ap.add_argument("--profile", action="store_true", help="profiling run (ncu): skip the e2e, FIR-128 sweep and CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
