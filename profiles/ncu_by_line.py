"""Stall samples and executed instructions of ONE kernel of an .ncu-rep, attributed to source lines.

    python profiles/ncu_by_line.py <report.ncu-rep> <object or .so with -lineinfo> <kernel name substring> [top N]

ncu's CLI exports per-SASS-instruction metrics (--page source --print-source sass) but not per-source-line ones; nvdisasm
--print-line-info gives the source line of every SASS instruction of the same binary.  The two listings are aligned by
instruction order inside the kernel."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, binary, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
col = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait",
                               "stall_lg", "stall_membar", "stall_math", "stall_branch_resolving", "stall_no_inst", "stall_not_selected", "stall_mio")}
inst = []
for r in rows[hdr + 1:]:
    if len(r) <= col["stall_mio"]:
        continue
    try:
        inst.append({k: (r[v] if k == "Source" else int(r[v])) for k, v in col.items()})
    except ValueError:
        pass
# nvdisasm: find the function, collect (line info, instruction) in order
cubins = subprocess.run(["cuobjdump", "-lelf", binary], capture_output=True, text=True).stdout.split()
cubins = [c for c in cubins if c.endswith(".cubin")]
import os, tempfile
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
lines_of = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, fn_lines, fname, in_fn = None, [], None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if in_fn and fn_lines:
                break
            in_fn = kname in m.group(1)
            continue
        if not in_fn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            fn_lines.append((cur, m.group(1).strip()))
    if fn_lines and len(fn_lines) == len(inst):
        lines_of = fn_lines
        break
    if fn_lines and lines_of is None:
        lines_of = fn_lines  # keep the first candidate; report the mismatch below
if lines_of is None:
    sys.exit("kernel not found in " + binary)
if len(lines_of) != len(inst):
    print("WARNING: %d SASS instructions in the report, %d in the binary -- not the same build?" % (len(inst), len(lines_of)))
agg = defaultdict(lambda: defaultdict(int))
n = min(len(inst), len(lines_of))
tot_s = sum(i["# Samples"] for i in inst) or 1
tot_i = sum(i["Instructions Executed"] for i in inst) or 1
for k in range(n):
    key = lines_of[k][0] or ("?", 0)
    for m in col:
        if m != "Source":
            agg[key][m] += inst[k][m]
print("kernel %s: %d SASS instructions, %d samples, %d warp instructions executed" % (kname, len(inst), tot_s, tot_i))
print("%-28s %7s %7s  %s" % ("file:line", "samp%", "inst%", "dominant stalls (samples)"))
src_cache = {}
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = sorted(((m[6:], c) for m, c in v.items() if m.startswith("stall_") and c), key=lambda x: -x[1])[:3]
    print("%-28s %6.1f%% %6.1f%%  %s" % ("%s:%d" % key, 100.0 * v["# Samples"] / tot_s, 100.0 * v["Instructions Executed"] / tot_i,
                                       ", ".join("%s %d" % s for s in st)))
