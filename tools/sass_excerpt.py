"""profiles/r2_k_trace.sass: the memory, collective and special-function instructions of the shipped traversal
kernels, from the library as built (cuobjdump -sass | cu++filt).  usage: python tools/sass_excerpt.py > profiles/r2_k_trace.sass"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "rtk_b200", "librtk_b200.so")
KERNELS = ("void k_resolve<(bool)0>", "k_push_rows(", "void k_trace<(int)2, (int)1, (bool)0, (bool)0, (bool)1>")
KEEP = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS", "LDGDEPBAR", "DEPBAR", "ATOM", "RED", "SHFL", "VOTE", "MATCH", "REDUX", "CREDUX",
        "BAR", "WARPSYNC", "MUFU", "S2R", "POPC", "FLO", "CCTL", "MEMBAR", "ERRBAR", "NANOSLEEP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
dem = subprocess.run(["cu++filt"], input=out, capture_output=True, text=True).stdout
print("# cuobjdump -sass rtk_b200/librtk_b200.so | cu++filt, excerpt (final kernels of round 2; tools/sass_excerpt.py)")
print("# sm_100a, nvcc 12.9, -O3 -lineinfo; kernel k_trace<2 lanes/ray, provable culling, no stats, closest hit, pair-distributed leaves>")
for b in re.split(r"\n\s*Function : ", dem)[1:]:
    name = b.split("\n", 1)[0].strip()
    if not any(k in name for k in KERNELS):
        continue
    lines = [ln for ln in b.splitlines() if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln)]
    print("\nFunction : " + name)
    print("  (%d SASS instructions; memory, collective and special-function instructions only)" % len(lines))
    for ln in lines:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and m.group(1).startswith(KEEP):
            print(ln.rstrip())
