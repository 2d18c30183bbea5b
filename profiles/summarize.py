"""Turn the raw ncu captures in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/r1_launches_X.md "title"
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep   profiles/r1_full_X.md     "title"
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst, title):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, x in enumerate(rows) if "init_ctl" in x["Kernel Name"]]
    sel = rows[idx[-1]:] if idx else rows
    agg = collections.OrderedDict()
    tot = 0.0
    for x in sel:
        name = re.sub(r"\(.*", "", x["Kernel Name"])[:70]
        t = float(x["Metric Value"])
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] = max(a[2], t)
        tot += t
    with open(dst, "w") as o:
        o.write("# %s\n\n" % title)
        o.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                "compare SHARES, not absolutes).  One solve = %d launches, %.3f ms of kernel time.\n\n" % (len(sel), tot / 1e6))
        o.write("| kernel | launches | total us | avg us | max us | share |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("| `%s` | %d | %.1f | %.1f | %.1f | %.1f%% |\n" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], v[2] / 1e3, 100 * v[1] / tot))
        o.write("\nPer-launch durations (us) in launch order:\n\n")
        for key in ("mv_tma", "rr_kernel", "subproj", "ritz", "orth_finish"):
            vals = [round(float(x["Metric Value"]) / 1e3, 1) for x in sel if key in x["Kernel Name"]]
            if vals:
                o.write("* `%s`: %s\n" % (key, vals))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]


def full(src, dst, title):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as o:
        o.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (`%s`).\n\n" % (title, src))
        for d in data:
            o.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % d[hdr.index("Kernel Name")][:90])
            for w in hdr:
                stall = w.startswith("smsp__average_warps_issue_stalled") and w.endswith("per_issue_active.ratio")
                if w in WANT or (stall and float(d[hdr.index(w)] or 0) > 0.05):
                    o.write("| %s | %s | %s |\n" % (w, d[hdr.index(w)], units[hdr.index(w)]))
            try:
                rd = float(d[hdr.index("dram__bytes_read.sum")])
                wr = float(d[hdr.index("dram__bytes_write.sum")])
                ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                o.write("\ntraffic = dram read + write = %.4f GB per launch\n\n" % ((rd * mul[ur] + wr * mul[uw]) / 1e9))
            except Exception:
                pass


if __name__ == "__main__":
    kind, src, dst, title = sys.argv[1:5]
    (launches if kind == "launches" else full)(src, dst, title)
