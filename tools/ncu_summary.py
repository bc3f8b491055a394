"""Summarise ncu outputs: launch list CSV -> per-kernel totals; .ncu-rep raw page -> key metrics per launch."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        n = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("pbd::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        tot[n] += v
        cnt[n] += 1
    s = sum(tot.values())
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:28s} launches={cnt[k]:4d} total={v / 1e6:9.3f} ms  avg={v / cnt[k] / 1e3:9.1f} us  share={100 * v / s:5.1f}%")
    print(f"{'TOTAL':28s} {'':13s} total={s / 1e6:9.3f} ms")


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def report(path, stalls=True):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:70])
        for k in KEYS:
            if k in hdr:
                print(f"    {k:70s} {r[hdr.index(k)]}")
        if stalls:
            items = [(hdr[i], float(r[i] or 0)) for i in range(len(hdr)) if "issue_stalled" in hdr[i] and "per_warp_active" in hdr[i]]
            for k, v in sorted(items, key=lambda x: -x[1])[:7]:
                print(f"    stall {k.split('issue_stalled_')[1]:55s} {v:.2f}")


if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a.endswith(".csv"):
            launches(a)
        else:
            report(a)
