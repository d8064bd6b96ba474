"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (run here, no GPU)."""
import collections, csv, re, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        t = float(row["Metric Value"].replace(",", ""))
        t = t / 1e3 if row["Metric Unit"] == "ns" else (t * 1e3 if row["Metric Unit"] == "ms" else t)
        m = re.search(r"k_gemm<(?:\(int\))?(\d), (?:\(int\))?(\d)(?:, (?:\(int\))?(\d))?>", row["Kernel Name"])
        key = f"k_gemm<INIT={m.group(1)},EPI={m.group(2)},MT={m.group(3) or 8}>" if m else row["Kernel Name"].split("(")[0]
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    out = [f"{'kernel':34s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:34s} {v[0]:8d} {v[1]/1e3:10.3f} {v[1]/tot*100:6.1f}% {v[1]/v[0]:10.1f}")
    out.append(f"{'all':34s} {sum(v[0] for v in agg.values()):8d} {tot/1e3:10.3f}")
    return "\n".join(out)

def raw(rep, wanted):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units, vals = r[0], r[1], r[2]
    out = []
    for h, u, v in zip(hdr, units, vals):
        if h in wanted or h in ("Kernel Name", "Grid Size", "Block Size"):
            out.append(f"{h:80s} {v:>22s} {u}")
    return "\n".join(out)

def dram(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rd = wr = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"].lower()
        v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        if row["Metric Name"].startswith("dram__bytes_read"):
            rd += v; n += 1
        else:
            wr += v
    return f"launches {n}\ndram__bytes_read.sum  total {rd/1e9:.3f} GB\ndram__bytes_write.sum total {wr/1e9:.3f} GB\nsum {(rd+wr)/1e9:.3f} GB"


WANTED = {"gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
          "sm__ops_path_tensor_src_fp64.sum.per_second", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"}

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    elif sys.argv[1] == "dram":
        print(dram(sys.argv[2]))
    else:
        print(raw(sys.argv[2], WANTED))
