#!/usr/bin/env python3
"""Per-source-line stall attribution of one kernel from an `ncu --set full --import-source on` report:
joins the SASS page (per-instruction sample counters by stall reason) with the line table of the kernel's
cubin (the object file of the translation unit, e.g. webrtc-audio-processing_b200/_obj/wap_k_echo_1.o).
usage: tools/ncu_stall_lines.py <report.ncu-rep> <object or .so> <mangled-name substring> <kernel regex> [top_n]
The object must be the build that was profiled ("opcode mismatches" says how many SASS lines disagree)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep=sys.argv[1]; so=sys.argv[2]; mangled=sys.argv[3]; kernel=sys.argv[4]
tmp=tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",os.path.abspath(so)],cwd=tmp,capture_output=True)
dis=""
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    d=subprocess.run(["nvdisasm","-g","-c",os.path.join(tmp,cubin)],capture_output=True,text=True).stdout
    if any(l.startswith(".text.") and mangled in l for l in d.splitlines()):
        dis=d;break
line_of={};cur=None;inside=False;stack=None
for l in dis.splitlines():
    if l.startswith(".text."):
        inside=mangled in l;continue
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(os.path.basename(m.group(1)),int(m.group(2)));continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);",l)
    if m: line_of[int(m.group(1),16)]=(cur,m.group(2).strip())
raw=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","sass","--csv","--launch-count","1","-k","regex:"+kernel],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
h=rows[hi]
ia=h.index("Address");isrc=h.index("Source")
cols={n:h.index(n) for n in ["# Samples","Instructions Executed","stall_long_sb","stall_short_sb","stall_wait","stall_no_inst","stall_mio","stall_selected","stall_branch_resolving","stall_not_selected"]}
base=None
agg=collections.defaultdict(lambda: collections.Counter())
mism=0;tot=0
for r in rows[hi+1:]:
    if len(r)<=ia or not r[ia].startswith("0x"): continue
    a=int(r[ia],16)
    if base is None: base=a
    loc,text=line_of.get(a-base,(None,""))
    tot+=1
    if text.split()[:1]!=r[isrc].split()[:1] and not r[isrc].strip().startswith("@"): mism+=1
    for n,c in cols.items(): agg[loc][n]+=int(r[c] or 0)
print("sass rows",tot,"opcode mismatches",mism)
T=collections.Counter()
for loc,c in agg.items(): T.update(c)
print("totals",dict(T))
byfile=collections.defaultdict(collections.Counter)
for loc,c in agg.items(): byfile[loc[0] if loc else None].update(c)
print("\nby file: inst% | samples% | long_sb% of all samples | short | wait | noinst")
for f,c in sorted(byfile.items(),key=lambda x:-x[1]["# Samples"]):
    print("  %-28s %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f"%(f,100*c["Instructions Executed"]/T["Instructions Executed"],100*c["# Samples"]/T["# Samples"],100*c["stall_long_sb"]/T["# Samples"],100*c["stall_short_sb"]/T["# Samples"],100*c["stall_wait"]/T["# Samples"],100*c["stall_no_inst"]/T["# Samples"]))
print("\ntop lines by long_sb")
for loc,c in sorted(agg.items(),key=lambda x:-x[1]["stall_long_sb"])[:int(sys.argv[5]) if len(sys.argv)>5 else 60]:
    print("  %-32s long %5.2f%% samples %5.2f%% inst %5.2f%%"%(loc,100*c["stall_long_sb"]/T["# Samples"],100*c["# Samples"]/T["# Samples"],100*c["Instructions Executed"]/T["Instructions Executed"]))
print("\ntop lines by samples: samples% inst% | noinst% wait% short% sel%")
for loc,c in sorted(agg.items(),key=lambda x:-x[1]["# Samples"])[:40]:
    print("  %-32s %5.2f %5.2f | %5.2f %5.2f %5.2f %5.2f"%(loc,100*c["# Samples"]/T["# Samples"],100*c["Instructions Executed"]/T["Instructions Executed"],100*c["stall_no_inst"]/T["# Samples"],100*c["stall_wait"]/T["# Samples"],100*c["stall_short_sb"]/T["# Samples"],100*c["stall_selected"]/T["# Samples"]))
