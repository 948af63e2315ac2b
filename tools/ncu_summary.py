"""Summarise an .ncu-rep: speed-of-light numbers, stall reasons, per-role stall samples."""
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, row = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tc", "sm__pipe_tc_cycles_active",
            "sm__pipe_tensor_cycles_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
            "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
            "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__inst_executed.sum"]
    for h, u, v in zip(hdr, units, row):
        if any(h == w or h.startswith(w) for w in want):
            print("%-75s %-10s %s" % (h, u, v))
    det = run([rep, "--page", "details"])
    for line in det.splitlines():
        if any(k in line for k in ("highest-utilized", "Executed Ipc", "Issue Slots", "Eligible",
                                   "One or More", "Tensor", "Duration", "Elapsed Cycles",
                                   "L2 Hit", "Mem Busy", "Max Bandwidth")):
            print(line.strip()[:150])
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hi = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
    h2 = src[hi]
    idx = {h: i for i, h in enumerate(h2)}
    si = idx["# Samples"]
    data = [r for r in src[hi + 1:] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data)
    print("total stall samples", tot, "SASS instructions", len(data))
    reasons = [h for h in h2 if h.startswith("stall_") and "(" not in h]
    agg = {h: sum(int(r[idx[h]]) for r in data if r[idx[h]].isdigit()) for h in reasons}
    print("stall reasons:", ", ".join("%s=%.0f%%" % (k[6:], 100 * v / max(tot, 1))
                                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print("top instructions by samples:")
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][si]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        why = sorted(((int(r[idx[h]]), h[6:]) for h in reasons if r[idx[h]].isdigit() and int(r[idx[h]]) > 0),
                     reverse=True)[:2]
        print("  #%-5d %5s (%4.1f%%) exec=%-8s %-64s %s" % (k, r[si], 100 * int(r[si]) / tot,
              r[idx["Instructions Executed"]], r[idx["Source"]][:64], why))


if __name__ == "__main__":
    main()
