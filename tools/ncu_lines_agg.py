"""Instruction and stall-sample shares per source line of one kernel of an .ncu-rep (captured with --import-source on, built with
-lineinfo), summed over all captured launches and template instances.  usage: python tools/ncu_lines_agg.py report.ncu-rep kernel_regex top"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg=collections.defaultdict(lambda:[0,0,""])
for r in rows:
    if len(r)>=2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if len(r)>8 and r[0]=="Line No": hdr=r; ii,si=hdr.index("Instructions Executed"),hdr.index("# Samples"); continue
    if hdr is None or len(r)<=ii or not r[0]: continue
    try: n,s=int(r[ii]),int(r[si])
    except ValueError: continue
    a=agg[(cur,r[0])]; a[0]+=n; a[1]+=s; a[2]=r[1].strip()[:110]
tot=sum(a[0] for a in agg.values()); tots=sum(a[1] for a in agg.values())
print(tot,tots)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3])]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/tots:5.1f}% samp {k[0]}:{k[1]:>4s} {a[2]}")
