#!/usr/bin/env python3
"""Attribute the executed warp instructions of k_tick to source lines.
usage: tools/ncu_lines.py <report.ncu-rep> [top_n] [kernel] [profiled .so] [mangled-name substring]
Joins `ncu --page source --print-source sass --csv` (per-SASS-instruction counters) with the line
table of the cubin inside libwap_b200.so (nvdisasm -g).  The .so must be the one that was profiled."""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel = sys.argv[3] if len(sys.argv) > 3 else "k_echo"
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "webrtc-audio-processing_b200", "libwap_b200.so")
# substring of the MANGLED name selecting the function in the cubin (template instances: k_echoILi1 = k_echo<1>)
mangled = sys.argv[5] if len(sys.argv) > 5 else kernel
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)],
               cwd=tmp, capture_output=True)
# the kernels live in several translation units: take the cubin whose text sections name the kernel
dis = ""
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    if any(l.startswith(".text.") and mangled in l for l in d.splitlines()):
        dis = d
        break
line_of = {}
cur = None
inside = False
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = mangled in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--launch-count", "1",
                      "-k", "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = None
per_line = collections.Counter()
per_file = collections.Counter()
samp_line = collections.Counter()
ops = collections.Counter()
total = 0
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    n = int(r[ie] or 0)
    s = int(r[isamp] or 0)
    loc, text = line_of.get(a - base, (None, ""))
    per_line[loc] += n
    samp_line[loc] += s
    per_file[loc[0] if loc else None] += n
    tk = text.split()
    ops[(tk[1] if tk and tk[0].startswith("@") and len(tk) > 1 else tk[0]) if tk else "?"] += n
    total += n
print("total warp instructions executed: %d" % total)
print("\nby file:")
for f, n in per_file.most_common():
    print("  %-28s %6.2f%%" % (f, 100.0 * n / total))
print("\ntop source lines (share of executed warp instructions | share of stall samples):")
tot_s = sum(samp_line.values()) or 1
for loc, n in per_line.most_common(top):
    src = ""
    if loc:
        try:
            src = open(os.path.join(ROOT, "webrtc-audio-processing_b200", "csrc", loc[0])).read().splitlines()[loc[1] - 1].strip()
        except Exception:
            pass
    print("  %5.2f%% %5.2f%%  %s:%s  %s" % (100.0 * n / total, 100.0 * samp_line[loc] / tot_s, loc[0] if loc else "?", loc[1] if loc else "?", src[:90]))
print("\ntop opcodes:")
for o, n in ops.most_common(16):
    print("  %-14s %5.2f%%" % (o, 100.0 * n / total))
