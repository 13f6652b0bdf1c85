"""Executed warp instructions of one kernel by SASS opcode and by pipe (from `ncu --page source --csv`).
usage: ncu_by_opcode.py <report.ncu-rep> [top N]
Pipes as in B300_MICROARCH.md: fma = FFMA/FMUL/FADD/IMAD/HFMA2/DFMA-class, alu = IADD3/LOP3/SHF/PRMT/FMNMX/ISETP/FSETP/SEL/LEA/MOV-class, ..."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2")
ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "FMNMX", "FMNMX3", "ISETP", "FSETP", "SEL", "FSEL", "LEA", "MOV", "PLOP3", "VIMNMX", "VIADD", "IABS", "FLO", "POPC", "I2FP", "FCHK", "CS2R", "BMSK", "SGXT", "VABSDIFF", "VIMNMX3", "IMNMX")
XU = ("MUFU", "F2F", "F2I", "I2F", "FRND", "F2FP")
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
LSU = ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "ATOMS", "ATOMG", "RED", "ATOM", "LDC", "LDSM")
CTRL = ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "BREAK", "WARPSYNC", "NOP", "BAR", "VOTE", "VOTEU", "SHFL", "S2R", "S2UR", "R2UR", "UMOV", "ULDC", "ULOP3", "UIADD3", "USEL", "UISETP", "BMOV", "MATCH", "REDUX", "YIELD", "DEPBAR", "ERRBAR", "MEMBAR", "CCTL", "UFLO", "UPOPC", "ULEA", "USHF", "UIMAD", "UPLOP3", "UPRMT", "R2P", "P2R")
def pipe(op):
    b = op.split(".")[0]
    for name, ops in (("fma", FMA), ("alu", ALU), ("xu", XU), ("fp64", FP64), ("lsu", LSU), ("ctrl/uniform", CTRL)):
        if b in ops:
            return name
    return "other"
byop, bypipe = collections.Counter(), collections.Counter()
lanes = collections.Counter()
tot = 0
for r in rows[2:]:
    try:
        ie = int(r[ix["Instructions Executed"]] or 0); te = int(r[ix["Thread Instructions Executed"]] or 0)
    except Exception:
        continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    b = op.split(".")[0]
    byop[b] += ie; lanes[b] += te; bypipe[pipe(op)] += ie; tot += ie
print(f"warp instructions executed: {tot:.4e}")
print("by pipe: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in bypipe.most_common()))
print(f"{'opcode':10s} {'winst%':>7s} {'lanes':>6s} pipe")
for b, v in byop.most_common(top):
    print(f"{b:10s} {100 * v / tot:7.2f} {lanes[b] / max(v, 1):6.2f} {pipe(b)}")
