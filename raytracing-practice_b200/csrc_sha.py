"""sha256 of the CODE of csrc/ (comments and whitespace removed): ties figures that can only come from an ncu capture
(DRAM traffic, issue-slot utilisation, active lanes) to the kernels they were measured on.  `tools/ncu_summary.py` writes it
into profiles/latest_ncu.json; `bench.py` quotes those figures only while the hash still matches — any token of any kernel
source changing invalidates them, rewording a comment does not."""
import glob
import hashlib
import os
import re


def _code_only(src):
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            i = n if j < 0 else j + 2
            out.append(" ")
        elif c == '"' or c == "'":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1])
            i = j + 1
        else:
            out.append(c)
            i += 1
    return re.sub(r"\s+", " ", "".join(out)).strip()


def csrc_sha(root=None):
    root = root or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(root, "raytracing-practice_b200", "csrc", "*"))):
        if f.endswith((".cu", ".cuh", ".h", ".hpp")):
            h.update(os.path.basename(f).encode())
            h.update(_code_only(open(f, encoding="utf-8").read()).encode())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(csrc_sha())
