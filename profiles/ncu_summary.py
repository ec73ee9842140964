#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): key metrics, SASS opcode mix, hot source lines.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 25]
"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-70s %-10s %s" % (w, units[i], [r[i] for r in data]))
    print("\nwarp stall reasons (% of warp-active cycles per issue):")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            st.append((float(data[0][i] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, n in sorted(st, reverse=True)[:10]:
        print("   %-28s %.2f" % (n, v))
    rows = page(rep, "source")
    d = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= 6 and r[0].startswith("0x"):
            d.append(r)
    tot_e = sum(int(r[5]) for r in d) or 1
    tot_s = sum(int(r[2]) for r in d) or 1
    h, hs = collections.Counter(), collections.Counter()
    for r in d:
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        h[op] += int(r[5]); hs[op] += int(r[2])
    print("\nSASS opcode mix (executed warp-instructions / stall samples): total %d / %d" % (tot_e, tot_s))
    for op, c in h.most_common(18):
        print("   %-10s %5.1f%%   samples %5.1f%%" % (op, 100.0 * c / tot_e, 100.0 * hs[op] / tot_s))
    print("\nhottest SASS instructions by stall samples:")
    for r in sorted(d, key=lambda r: -int(r[2]))[:nlines]:
        print("   %6s %5.1f%%  exec %8s  %s" % (r[2], 100.0 * int(r[2]) / tot_s, r[5], r[1].strip()[:70]))


if __name__ == "__main__":
    main()
