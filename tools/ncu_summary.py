#!/usr/bin/env python3
"""Key figures of an .ncu-rep per kernel launch (raw page) and, optionally, per source line sorted by stall samples.
usage: python tools/ncu_summary.py rep.ncu-rep [kernel-name-substring] [--lines N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_op_read_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if sub and sub not in name:
        continue
    print("==", name[:110])
    for w in want:
        if w in hdr:
            print("   %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = sorted(((float(r[hdr.index(s)] or 0), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)
    print("   stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
if nlines:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    sections, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "File Path":
            cur = {"file": r[1], "func": None, "hdr": None, "rows": []}; sections.append(cur)
        elif r and r[0] == "Function Name" and cur is not None: cur["func"] = r[1]
        elif r and r[0] == "Line No" and cur is not None: cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None: cur["rows"].append(r)
    agg, ti_, ts_ = collections.OrderedDict(), 0, 0
    for s in sections:
        if sub and sub not in (s["func"] or ""):
            continue
        h = s["hdr"]; ii, ti, si = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
        line = src_ = None
        num = lambda x: int(float(x)) if x not in ("-", "") else 0
        for r in s["rows"]:
            if r[0] != "":
                line, src_ = r[0], r[1]; continue
            a = agg.setdefault((s["file"].split("/")[-1], int(line), src_.strip()[:105]), [0, 0, 0])
            a[0] += num(r[ii]); a[1] += num(r[ti]); a[2] += num(r[si]); ti_ += num(r[ii]); ts_ += num(r[si])
    print("per line, by stall samples (of %d) / instructions (of %d):" % (ts_, ti_))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:nlines]:
        print("%5.1f%% stall %5.1f%% inst thr %4.1f %s:%d %s" % (100 * a[2] / max(ts_, 1), 100 * a[0] / max(ti_, 1), a[1] / max(a[0], 1), k[0], k[1], k[2]))
