"""Summarise an .ncu-rep (run where ncu is installed): key raw metrics per launch + hot SASS blocks."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print("   %-70s %s %s" % (h, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# several kernels are concatenated; split on the "Kernel Name" marker rows
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
stalls = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_math", "stall_lg",
          "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst", "stall_membar"]
seen = set()
for kq in kern:
    if not kq["rows"] or "Address" not in kq["rows"][0]:
        continue  # the source page repeats each kernel's marker row: keep the block that carries the table
    h = kq["rows"][0]
    ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = []
    for r in kq["rows"][1:]:
        try:
            data.append((r[isrc], int(r[iex]), int(r[ismp]), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[1] for d in data)
    ts = sum(d[2] for d in data) or 1
    print("== SASS", kq["name"][:70], "instr", tot, "samples", ts)
    agg = {s: sum(int(d[3][h.index(s)] or 0) for d in data) for s in stalls if s in h}
    print("   stalls %:", {k.replace("stall_", ""): round(100 * v / ts, 1) for k, v in agg.items() if v})
    ops = collections.Counter()
    for s, ex, smp, _ in data:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        ops[op.split(".")[0]] += ex
    print("   opcodes %:", {k: round(100 * v / tot, 1) for k, v in ops.most_common(14)})
    blocks, cb = [], None
    for i, (s, ex, smp, _) in enumerate(data):
        if cb and cb["ex"] == ex:
            cb["n"] += 1
            cb["smp"] += smp
            cb["end"] = i
        else:
            cb = {"ex": ex, "n": 1, "smp": smp, "start": i, "end": i}
            blocks.append(cb)
    blocks.sort(key=lambda b: -b["smp"])
    for b in blocks[:top]:
        print("   block idx %d-%d: %d instrs x %d exec = %.1f%% instr, %.1f%% samples" % (
            b["start"], b["end"], b["n"], b["ex"], 100.0 * b["ex"] * b["n"] / tot, 100.0 * b["smp"] / ts))
    if len(sys.argv) > 3:
        lo, hi = (int(v) for v in sys.argv[3].split("-"))
        for i in range(lo, min(hi, len(data))):
            print("     %5d %9d %5d  %s" % (i, data[i][1], data[i][2], data[i][0][:100]))
