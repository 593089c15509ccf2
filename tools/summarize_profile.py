"""Turns an ncu report (gpurun_out/<tag>/*.ncu-rep) into the small tracked summaries under profiles/:
   <name>_metrics.csv  one row per kernel: duration, DRAM bytes, instructions, issue %, pipe %, stall breakdown
   <name>_hot_sass.txt the hottest SASS lines of each kernel (needs --import-source / -lineinfo)
usage: python tools/summarize_profile.py gpurun_out/q06/prof512.ncu-rep profiles/r01_entropy_512frames
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    stalls = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h]
    with open(dst + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")
            for k in KEYS:
                if k in hdr and r[hdr.index(k)] != "":
                    w.writerow([name, k, r[hdr.index(k)], units[hdr.index(k)]])
            for i, h in enumerate(hdr):  # every execution pipe: share of its peak issue rate while the SM was active
                if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active") and r[i] not in ("", "0"):
                    w.writerow([name, h, r[i], units[i]])
            top = sorted(((float(r[i]) if r[i] else 0.0, h) for i, h in stalls), reverse=True)[:8]
            for v, h in top:
                w.writerow([name, "stall:" + h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), "%.3f" % v, "warps per issue"])
    src = ncu_csv(rep, "source")
    with open(dst + "_hot_sass.txt", "w") as f:
        kernel = None
        body = []

        def flush():
            if not body:
                return
            h = body[0]
            ia, isrc, iex, ith, ismp = (h.index(c) for c in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
            lines = body[1:]
            tot = sum(int(x[iex]) for x in lines) or 1
            smp = sum(int(x[ismp]) for x in lines) or 1
            f.write("== %s: %d warp instructions executed, %d stall samples\n" % (kernel, tot, smp))
            f.write("   offset   executed  thr/inst  samples%%  SASS\n")
            base = int(lines[0][ia], 16)
            hot = sorted(lines, key=lambda x: -int(x[ismp]))[:40]
            for x in sorted(hot, key=lambda x: int(x[ia], 16)):
                ex = int(x[iex])
                f.write("   %05x %10d  %6.1f  %7.2f  %s\n" % (int(x[ia], 16) - base, ex, int(x[ith]) / max(ex, 1), 100.0 * int(x[ismp]) / smp, x[isrc].strip()[:100]))
            f.write("\n")

        for r in src:
            if r and r[0] == "Kernel Name":
                flush()
                kernel = r[1].split("(")[0]
                body = []
            elif r:
                body.append(r)
        flush()
    print("wrote", dst + "_metrics.csv", dst + "_hot_sass.txt")


if __name__ == "__main__":
    main()
