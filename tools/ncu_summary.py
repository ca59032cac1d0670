#!/usr/bin/env python3
"""Summarise an .ncu-rep (read HERE, no GPU needed) into the text committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_tuner.ncu-rep > profiles/r01_tuner.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print("  %-68s %s %s" % (k, d[k], u.get(k, "")))
        stalls = []
        for h, v in d.items():
            if "issue_stalled" in h and h.endswith(".pct"):
                try:
                    stalls.append((float(v), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            print("  stall %-62s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))
    src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv"]).splitlines()))
    if len(src) > 2:
        h = src[1]
        try:
            ia, isamp = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
        except ValueError:
            return
        data = []
        for r in src[2:]:
            try:
                data.append((int(r[isamp]), r[ia].strip()))
            except (ValueError, IndexError):
                pass
        tot = sum(s for s, _ in data) or 1
        # regions delimited by CTA barriers: cumulative share of stall samples and of executed instructions
        try:
            iex = h.index("Instructions Executed")
            ex = []
            for r in src[2:]:
                try:
                    ex.append(int(r[iex]))
                except (ValueError, IndexError):
                    ex.append(0)
            ex = ex[:len(data)]
            bars = [i for i, (_, ins) in enumerate(data) if ins.startswith("BAR") or " BAR." in ins]
            edges = [0] + bars + [len(data)]
            etot = sum(ex) or 1
            print("  regions between barriers (instruction index range: %% of stall samples, %% of executed warp instructions):")
            for a, b in zip(edges[:-1], edges[1:]):
                if b > a:
                    print("    [%5d, %5d)  samples %5.1f%%  executed %5.1f%%" % (a, b, 100.0 * sum(x for x, _ in data[a:b]) / tot, 100.0 * sum(ex[a:b]) / etot))
        except ValueError:
            pass
        print("  top stall-sample instructions (of %d samples):" % tot)
        for s, ins in sorted(data, reverse=True)[:10]:
            print("    %5.1f%%  %s" % (100.0 * s / tot, ins[:100]))


if __name__ == "__main__":
    main()
