#!/usr/bin/env python3
"""Static SASS instruction count per source file / line for one kernel of libwap_b200.so:
tools/static_lines.py [kernel] [top] [.so]   (code size is instruction-cache pressure)"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
kernel = sys.argv[1] if len(sys.argv) > 1 else "k_echo"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "webrtc-audio-processing_b200", "libwap_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("wap_engine") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
per_line, per_file = collections.Counter(), collections.Counter()
cur, inside, total = None, False, 0
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kernel in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l):
        per_line[cur] += 1
        per_file[cur[0] if cur else None] += 1
        total += 1
print("%s: %d SASS instructions = %.0f KB" % (kernel, total, total * 16 / 1024))
for f, n in per_file.most_common():
    print("  %-28s %7d  %5.1f%%" % (f, n, 100.0 * n / total))
print("top lines:")
for loc, n in per_line.most_common(top):
    print("  %6d  %s" % (n, loc))
