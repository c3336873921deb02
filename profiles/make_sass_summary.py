"""Per-kernel SASS evidence for the hot kernels of libwholegraph_b200.so (cuobjdump -sass): which memory / atomic / copy-engine /
tensor instructions each one contains.  Run here (no GPU needed):  python profiles/make_sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cugraph-gnn_b200", "lib", "libwholegraph_b200.so")
PAT = re.compile(r"\b(LDG\.E(?:\.[A-Z0-9_]+)*|STG\.E(?:\.[A-Z0-9_]+)*|ATOMG(?:\.[A-Z0-9_]+)*|ATOMS(?:\.[A-Z0-9_]+)*|REDG(?:\.[A-Z0-9_]+)*|RED\.E(?:\.[A-Z0-9_]+)*|"
                 r"UBLKCP(?:\.[A-Z0-9_]+)*|UTMA[A-Z]*(?:\.[A-Z0-9_]+)*|UTC[A-Z]*MMA(?:\.[A-Z0-9_]+)*|UTCBAR(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Z0-9_]+)*|STTM(?:\.[A-Z0-9_]+)*|SYNCS(?:\.[A-Z0-9_]+)*|"
                 r"MATCH(?:\.[A-Z0-9_]+)*|UCGABAR[A-Z_]*|HMMA(?:\.[A-Z0-9_]+)*|LDGSTS(?:\.[A-Z0-9_]+)*)")
WANT = ("fz_label_kernel", "fz_emit", "rows_bulk_gather_kernel", "rows_copy_kernel", "uniform_small_kernel", "mh_insert_kernel", "mh_compact_kernel",
        "csr_aggregate", "sage_", "weighted_kernel")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, stats, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        continue
    sizes[cur] = sizes.get(cur, 0) + 1
    for tok in PAT.findall(line):
        stats.setdefault(cur, collections.Counter())[tok] += 1
names = subprocess.run(["c++filt"] + list(sizes), capture_output=True, text=True).stdout.splitlines()
demangled = dict(zip(sizes, names))
print("# SASS evidence, sm_100a, %s" % os.path.relpath(LIB, ROOT))
print("# mnemonics: LDG.E.128/.256 vector loads, ATOMG.* global atomics (CAS.64 / CAS.128 / MIN), UBLKCP = cp.async.bulk (copy engine),")
print("# SYNCS = mbarrier ops, UCGABAR = cluster barrier, UTC*MMA / LDTM = tcgen05 (tensor cores / tensor memory)")
for fn in sizes:
    d = demangled.get(fn, fn)
    if not any(w in d for w in WANT):
        continue
    c = stats.get(fn, {})
    print("\n%s\n  %d SASS instructions" % (d[:230], sizes[fn]))
    for tok, n in sorted(c.items()):
        print("    %-34s %d" % (tok, n))
