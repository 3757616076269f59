"""Turn ncu output brought back from the GPU box into small, committable summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_tron.md
  python tools/ncu_summary.py full gpurun_out/prof_tron.ncu-rep profiles/r01_tron_step_full.md [traffic_json]
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg, order = {}, []
    for r in rows[1:]:
        v = num(r[vi])
        if v is None or r[mi] != "gpu__time_duration.sum":
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        k = r[ki].split("(")[0]
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(a[1] for a in agg.values())
    mine = sum(a[1] for k, a in agg.items() if not k.startswith("void at::") and "elementwise" not in k)
    out = ["# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
           "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k[:90], n, t, t / n, 100 * t / tot))
    out += ["", "Total %.1f us over %d launches; library (non-torch) kernels: %.1f %%." % (tot, sum(a[0] for a in agg.values()), 100 * mine / tot)]
    open(dst, "w").write("\n".join(out) + "\n")


def full(src, dst, traffic=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full --clock-control none: `%s`" % src.split("/")[-1], ""]
    rec = []
    for row in rows[2:]:
        d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
        rec.append(d)
        out.append("## %s" % d["Kernel Name"][0].split("(")[0])
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        out.append("")
    # stall samples from the source page (first kernel)
    srcpage = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(srcpage.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if hi:
        h = rows[hi[0]]
        col = {x: i for i, x in enumerate(h)}
        end = hi[1] - 1 if len(hi) > 1 else len(rows)
        data = [r for r in rows[hi[0] + 1:end] if len(r) == len(h)]
        stall = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot = sum(num(r[col["# Samples"]]) or 0 for r in data) or 1
        agg = {x: sum(num(r[col[x]]) or 0 for r in data) for x in stall}
        out += ["## warp-state samples (first captured launch)", "", "| stall reason | samples | share |", "|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
            out.append("| %s | %d | %.1f %% |" % (k, v, 100 * v / tot))
        out += ["", "Top instructions by samples:", "", "| samples | SASS |", "|---|---|"]
        for r in sorted(data, key=lambda r: -(num(r[col["# Samples"]]) or 0))[:10]:
            out.append("| %s | `%s` |" % (r[col["# Samples"]], r[col["Source"]].strip()[:90]))
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    if traffic and rec:
        def b(d, k):
            v, u = d[k]
            return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per = [b(d, "dram__bytes_read.sum") + b(d, "dram__bytes_write.sum") for d in rec]
        json.dump({"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per),
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full; ncu runs each "
                           "kernel in isolation, so write-backs that leave L2 after the kernel ends are not counted",
                   "source": dst}, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
