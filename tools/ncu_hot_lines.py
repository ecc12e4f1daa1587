#!/usr/bin/env python
"""Join the per-instruction stall samples of an ncu report (SASS source page) with nvdisasm line info.

usage: ncu_hot_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top_n]
Prints the source lines with the most warp-stall samples (needs -lineinfo at compile time)."""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, kre, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, isrc, isamp, iex = H.index("Address"), H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
samples = []
for r in rows[hdr + 1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    samples.append((int(r[ia], 16), r[isrc], int(r[isamp] or 0), int(r[iex] or 0)))
base = samples[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# per-section address -> line; concatenate sections in the order they appear after the kernel's own section
sec, cur, line = None, None, None
maps = collections.OrderedDict()
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        sec = m.group(1); maps[sec] = {}; line = None; continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and sec:
        maps[sec][int(m.group(1), 16)] = (line, m.group(2))
ksec = next(s for s in maps if re.search(kre, s))
agg, agg_ex = collections.Counter(), collections.Counter()
tot = 0
for addr, src, ns, ex in samples:
    off = addr - base
    ln = maps[ksec].get(off, (None, None))[0]
    agg[ln] += ns; agg_ex[ln] += ex; tot += ns
print(f"kernel section {ksec}: {len(samples)} instructions, {tot} samples")
for ln, ns in agg.most_common(top):
    print(f"{100.0*ns/max(tot,1):6.2f}%  samples {ns:7d}  inst_exec {agg_ex[ln]:10d}  {ln}")
