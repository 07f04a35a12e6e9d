#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel, registers / spills / shared
memory from the cubin's resource usage and counts of the SASS mnemonics the design relies on —
128-bit global accesses (LDG/STG.E.128), cp.async (LDGSTS), programmatic dependent launch
(ACQBULK = griddepcontrol.wait, PREEXIT = launch_dependents), local-memory traffic (LDL/STL)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "herald_b200", "lib", "libherald_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function ([^:]+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        usage[cur] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    for key, pat in (("LDG.128", r"\bLDG\.E[\w.]*\.128"), ("STG.128", r"\bSTG\.E[\w.]*\.128"),
                     ("LDGSTS", r"\bLDGSTS"), ("ACQBULK", r"\bACQBULK"), ("PREEXIT", r"\bPREEXIT"),
                     ("LDL", r"\bLDL\b"), ("STL", r"\bSTL\b"), ("ATOMG/RED", r"\b(ATOMG|RED)\."),
                     ("VOTE", r"\bVOTE\."), ("BAR", r"\bBAR\."), ("SYNCS", r"\bSYNCS\."),
                     ("UBLKCP", r"\bUBLKCP"), ("ARRIVES", r"\bARRIVES\.")):
        if re.search(pat, line):
            counts[cur][key] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n
cols = ["LDG.128", "STG.128", "LDGSTS", "ACQBULK", "PREEXIT", "LDL", "STL", "ATOMG/RED", "VOTE", "BAR", "SYNCS",
        "UBLKCP", "ARRIVES"]
print("%-78s %4s %5s %6s %5s | %s" % ("kernel", "regs", "stack", "smem", "local", " ".join("%9s" % c for c in cols)))
for fn in sorted(counts, key=demangle):
    name = re.sub(r"\(anonymous namespace\)::|hb::", "", demangle(fn))
    name = re.sub(r"\(.*", "", name)[:78]
    u = usage.get(fn, (0, 0, 0, 0))
    print("%-78s %4d %5d %6d %5d | %s" % (name, u[0], u[1], u[2], u[3], " ".join("%9d" % counts[fn][c] for c in cols)))
missing = [demangle(f) for f in counts if counts[f]["ACQBULK"] == 0]
print("\nkernels: %d; without griddepcontrol.wait: %d %s" % (len(counts), len(missing), missing[:3]))
