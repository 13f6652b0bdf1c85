"""Aggregate an ncu SASS-level source page by PHASE = the function the kernel body called
(inline chains from `nvdisasm -gi`, collapsed to the two outermost frames).
usage: ncu_by_phase.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [depth=2]
Columns: share of PC samples, share of warp instructions, share of thread instructions, lanes = thread/warp."""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1:4]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "raytracing-practice_b200", "csrc")

# function start lines per source file (regex on definitions; a function extends to the next start)
fn_starts = {}
for f in os.listdir(CSRC):
    starts = []
    for n, l in enumerate(open(os.path.join(CSRC, f), errors="replace"), 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:__global__|__device__)[^;(]*?\b([A-Za-z_]\w*)\s*\(", l)
        if m and not l.strip().endswith(";"):
            starts.append((n, m.group(1)))
    fn_starts[f] = starts


def fn_of(file, line):
    s = fn_starts.get(file)
    if not s:
        return file
    i = bisect.bisect_right([x[0] for x in s], line) - 1
    return s[i][1] if i >= 0 else file


tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
addr2chain, chain, pending, infn = {}, [], [], False
for l in sass:
    if l.startswith("\t.section\t.text."):
        infn = kern in l
        continue
    if not infn:
        continue
    if l.lstrip().startswith("//## File"):
        pending.append([(os.path.basename(a), int(b)) for a, b in re.findall(r'"([^"]+)", line (\d+)', l)][0])
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        if pending:
            chain, pending = pending, []
        addr2chain[int(m.group(1), 16)] = chain  # innermost first
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for r in rows[2:]:
    try:
        a = int(r[ix["Address"]], 16)
    except Exception:
        continue
    base = a if base is None else base
    ch = addr2chain.get(a - base, [])
    names = []
    for file, line in reversed(ch):  # outermost first
        n = fn_of(file, line)
        if file not in fn_starts:
            break
        if not names or names[-1] != n:
            names.append(n)
    label = " > ".join(names[:depth]) if names else "?"
    s = int(r[ix["# Samples"]] or 0); ie = int(r[ix["Instructions Executed"]] or 0); te = int(r[ix["Thread Instructions Executed"]] or 0)
    g = agg[label]; g[0] += s; g[1] += ie; g[2] += te
    tot[0] += s; tot[1] += ie; tot[2] += te
print(f"total samples {tot[0]}, warp-inst {tot[1]:.4e}, thread-inst {tot[2]:.4e}, avg active lanes {tot[2]/max(tot[1],1):.2f}")
print(f"{'phase':52s} {'samples%':>8s} {'winst%':>7s} {'tinst%':>7s} {'lanes':>6s}")
for label, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if g[1] == 0 and g[0] == 0:
        continue
    print(f"{label:52s} {100*g[0]/max(tot[0],1):8.2f} {100*g[1]/max(tot[1],1):7.2f} {100*g[2]/max(tot[2],1):7.2f} {g[2]/max(g[1],1):6.2f}")
