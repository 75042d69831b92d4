"""Aggregate an ncu --csv launch list (gpu__time_duration + dram bytes) per kernel name.

    python tools/launch_summary.py gpurun_out/cfg5_launches.csv
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ik, im, iv, iid, iu = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
d = collections.OrderedDict()
for r in rows:
    d.setdefault((r[iid], r[ik]), {})[r[im]] = (float(r[iv].replace(",", "")), r[iu])
agg = collections.OrderedDict()
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for (_, k), m in d.items():
    a = agg.setdefault(k.split("(")[0], [0, 0.0, 0.0, 0.0])
    t, u = m["gpu__time_duration.sum"]
    a[0] += 1
    a[1] += t / {"ns": 1e3, "us": 1.0, "ms": 1e-3}.get(u, 1e3)
    for j, n in ((2, "dram__bytes_read.sum"), (3, "dram__bytes_write.sum")):
        if n in m:
            a[j] += m[n][0] * scale[m[n][1]]
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print("%-44s n=%3d total %9.1f us (%4.1f%%)  avg %8.1f us  rd %8.1f MB wr %8.1f MB  %5.0f GB/s" % (
        k[:44], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / a[0] / 1e6, a[3] / a[0] / 1e6, (a[2] + a[3]) / a[1] / 1e3))
