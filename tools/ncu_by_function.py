"""PC samples / executed instructions of one kernel by SOURCE FUNCTION of rt_device.cuh (line ranges taken from the file as it
is now: run it on a capture of the current build).   usage: ncu_by_function.py <report.ncu-rep> <lib.so> <kernel substring>"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, lib, kern = sys.argv[1:4]
src = open(os.path.join(ROOT, "raytracing-practice_b200", "csrc", "rt_device.cuh")).read().splitlines()
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:__device__|template|static|inline).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
    if l.startswith("__device__") and m:
        starts.append((i, m.group(1)))
starts.sort()
def func_of(line):
    name = "rt_device.cuh:top"
    for ln, n in starts:
        if ln <= line:
            name = n
        else:
            break
    return name
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, lib, kern, "100000"], capture_output=True, text=True).stdout
agg = {}
for l in out.splitlines():
    m = re.match(r"(\S+):(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\d+)", l)
    if not m:
        continue
    f, ln, sp, wi, la = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    key = func_of(ln) if f == "rt_device.cuh" else f
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += sp; a[1] += wi; a[2] += wi * la
print(f"{'function':32s} {'samples%':>8s} {'winst%':>7s} {'lanes':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{k:32s} {v[0]:8.2f} {v[1]:7.2f} {v[2] / max(v[1], 1e-9):6.2f}")
