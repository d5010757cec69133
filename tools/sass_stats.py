"""Static SASS statistics of one kernel of librtk_b200.so: instruction count, opcode-class mix (which pipe),
spills.  A proxy for the dynamic instruction count when no GPU is at hand.
usage: python tools/sass_stats.py [substring of the demangled kernel name] [library]"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
pat = sys.argv[1] if len(sys.argv) > 1 else "k_trace<(int)2, (int)1, (bool)0, (bool)0, (bool)1>"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "..", "rtk_b200", "librtk_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = subprocess.run(["cu++filt"], input=out, capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", dem)
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HMUL2", "HADD2")
ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "FMNMX", "FMNMX3", "SEL", "FSEL", "ISETP", "FSETP", "MOV", "LEA", "VIMNMX", "VIMNMX3", "IMNMX", "PLOP3", "P2R", "R2P", "POPC", "FLO", "BREV", "BMSK", "SGXT", "IABS", "FCHK", "VABSDIFF", "CS2R", "FSET", "I2FP", "VIADD")
LSU = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "ATOMS", "ATOMG", "RED", "LD", "ST", "LDSM", "LDC", "ULDC", "LDGDEPBAR", "DEPBAR")
XU = ("MUFU", "F2I", "I2F", "F2F", "FRND")
CTL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "VOTE", "SHFL", "BAR", "NANOSLEEP", "CALL", "RET", "BRX", "YIELD", "NOP", "S2R", "S2UR", "REDUX", "MATCH", "VOTEU", "R2UR", "ELECT")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for ln in b.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", ln)
        if m:
            ops[m.group(1)] += 1
    total = sum(ops.values())
    cls = collections.Counter()
    for o, c in ops.items():
        k = "fma" if o in FMA else "alu" if o in ALU else "lsu" if o in LSU else "xu" if o in XU else "ctl" if o in CTL else "other"
        cls[k] += c
    print(name.strip())
    print("  instructions %d | %s" % (total, "  ".join("%s %d" % kv for kv in sorted(cls.items(), key=lambda kv: -kv[1]))))
    print("  top: " + "  ".join("%s %d" % kv for kv in ops.most_common(24)))
    print("  spills: STL %d LDL %d   S2R %d" % (ops["STL"], ops["LDL"], ops["S2R"]))
