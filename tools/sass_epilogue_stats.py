"""Offline check of the conv epilogue code generation: for every LDTM (tcgen05.ld) in k_conv_tc<EG>, the instruction mix up to the next
UTMASTG / LDTM (one 32-column chunk of the fast path).  Usage: python tools/sass_epilogue_stats.py [EG]"""
import collections
import re
import subprocess
import sys

eg = sys.argv[1] if len(sys.argv) > 1 else "2"
sass = subprocess.run(["cuobjdump", "-sass", "cartoonsegmentation_b200/_build/tc_conv.o"], capture_output=True, text=True).stdout
fn, lines = False, []
for ln in sass.splitlines():
    if "Function :" in ln:
        fn = f"k_conv_tcILi{eg}" in ln
    elif fn:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            lines.append(m.group(1).strip())
idx = [i for i, l in enumerate(lines) if "LDTM" in l]
for n, i in enumerate(idx):
    end = next((j for j in range(i + 1, min(len(lines), i + 900)) if "UTMASTG" in lines[j] or "LDTM" in lines[j]), None)
    if end is None or "LDTM" in lines[end]:
        continue
    ops = collections.Counter()
    for l in lines[i:end]:
        op = l.split()[1] if l.startswith("@") else l.split()[0]
        ops[op.split(".")[0] + (".MOV" if ".MOV" in op else "")] += 1
    tot = sum(ops.values())
    print(f"LDTM #{n}: {tot} instrs to UTMASTG:", ", ".join(f"{k} {v}" for k, v in ops.most_common(12)))
