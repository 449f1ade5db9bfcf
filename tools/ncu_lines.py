#!/usr/bin/env python3
"""Aggregate an ncu `--page source --csv --print-source sass,cuda` dump per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda | python tools/ncu_lines.py [kernel_index] [top]"""
import csv, sys, collections
want = int(sys.argv[1]) if len(sys.argv) > 1 else -1
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(sys.stdin))
# split into per-kernel, per-file sections: a section starts with "File Path"
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "func": None, "hdr": None, "rows": []}
        sections.append(cur)
    elif r and r[0] == "Function Name" and cur is not None:
        cur["func"] = r[1]
    elif r and r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
# group sections by kernel occurrence: new kernel when file list restarts is unknown; use order of (func) changes
funcs = []
for s in sections:
    if not funcs or funcs[-1][0] != s["func"] or any(x["file"] == s["file"] for x in funcs[-1][1]):
        funcs.append((s["func"], []))
    funcs[-1][1].append(s)
print("kernels in report:", [(i, f[0][:60]) for i, f in enumerate(funcs)])
f = funcs[want]
agg = collections.OrderedDict()
tot_inst = tot_samp = 0
for s in f[1]:
    h = s["hdr"]
    ii, ti, si = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    line, src = None, None
    for r in s["rows"]:
        if r[0] != "":
            line, src = r[0], r[1]
            continue
        key = (s["file"].split("/")[-1], int(line), src.strip()[:110])
        a = agg.setdefault(key, [0, 0, 0])
        num = lambda x: int(float(x)) if x not in ('-', '') else 0
        a[0] += num(r[ii]); a[1] += num(r[ti]); a[2] += num(r[si])
        tot_inst += num(r[ii]); tot_samp += num(r[si])
print("kernel:", f[0][:100], "warp-inst", tot_inst, "samples", tot_samp)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  thr/inst %4.1f  %s:%d  %s" % (100.0 * a[0] / max(tot_inst, 1), 100.0 * a[2] / max(tot_samp, 1), a[1] / max(a[0], 1), k[0], k[1], k[2]))
