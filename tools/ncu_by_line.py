"""Aggregate an ncu SASS-level source page by CUDA source line.
usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [top N]
Joins `ncu --page source --csv` (per-SASS-address samples / executed counts) with
`nvdisasm --print-line-info` (address -> file:line, inlined call chain collapsed to the innermost)."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, infn = {}, None, False
for l in sass:
    if l.startswith("\t.section\t.text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if "inlined at" not in l or cur is None:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        # keep the innermost location (first of an inline chain); chain continuation lines contain 'inlined at'
        if "inlined at" in l and l.strip().startswith("//## File") and cur is not None:
            pass
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
base = None
tot = [0, 0, 0]
for r in rows[2:]:
    try:
        a = int(r[ix["Address"]], 16)
    except Exception:
        continue
    base = a if base is None else base
    off = a - base
    loc = addr2line.get(off, (("?", 0), ""))[0]
    s = int(r[ix["# Samples"]] or 0); ie = int(r[ix["Instructions Executed"]] or 0); te = int(r[ix["Thread Instructions Executed"]] or 0)
    g = agg[loc]; g[0] += s; g[1] += ie; g[2] += te; g[3] += 1
    tot[0] += s; tot[1] += ie; tot[2] += te
print(f"total samples {tot[0]}, warp-inst {tot[1]:.3e}, thread-inst {tot[2]:.3e}, avg active lanes {tot[2]/max(tot[1],1):.2f}")
print(f"{'file:line':32s} {'samples%':>8s} {'winst%':>7s} {'lanes':>6s} {'#sass':>5s}")
for loc, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{loc[0]+':'+str(loc[1]):32s} {100*g[0]/tot[0]:8.2f} {100*g[1]/max(tot[1],1):7.2f} {g[2]/max(g[1],1):6.2f} {g[3]:5d}")
