#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list (gpurun_out/*.csv) per kernel: launches, total time, share, tensor-pipe activity.
usage: summarize_ncu_launches.py <csv> [skip_fraction]   (skip_fraction: ignore the first part of the launches, e.g. 0.5 = second half only)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {n: i for i, n in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    key = (int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Grid Size"]])
    try:
        per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        pass
keys = list(per)[int(len(per) * skip):]
agg = collections.OrderedDict()
for k in keys:
    m = per[k]
    name = k[1].split("(")[0].replace("void ", "").replace("wb::", "")[:44] + " grid " + k[2]
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    t = m.get("gpu__time_duration.sum", 0.0) / 1e3
    a[0] += 1; a[1] += t
    a[2] += m.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    a[3] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("| kernel (grid) | launches | total µs | share | tensor pipe active (time-weighted) | DRAM bytes |")
print("|---|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% | %.1f %% | %.3g |" % (n, a[0], a[1], 100 * a[1] / tot, a[2] / a[1] if a[1] else 0.0, a[3]))
print("\ntotal %.1f µs over %d launches (under ncu: serialised, cold caches; shares are what to read, not absolutes)" % (tot, len(keys)))
