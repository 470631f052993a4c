"""Count the Blackwell-specific SASS mnemonics per kernel of the built library: `python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.md`.
Mnemonics (B200_PROFILING.md): UTCHMMA/UTCQMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA tensor load/store, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, FFMA2 = packed fp32 FMA, REDG = global reductions (red.global), USETMAXREG = setmaxnreg, UCGABAR = cluster barrier,
HMMA = mma.sync (legacy path)."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cartoonsegmentation_b200", "libcsb200.so")
MN = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA2", "HMMA", "REDG", "ATOMG", "MUFU", "USETMAXREG", "UCGABAR_ARV"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_n"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                per[cur][k] += 1


def short(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"^void ", "", r).replace("(anonymous namespace)::", "").replace("csb::", "")
    r = re.sub(r"\(.*$", "", r)
    return r[:70]


tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"# SASS summary of `{os.path.basename(lib)}` (cuobjdump -sass, CUDA 12.9)\n")
print(f"cubin architectures: {', '.join(arch)}; {len(per)} kernels, {tot['_n']} SASS instructions.\n")
print("Totals: " + ", ".join(f"`{k}` x{tot[k]}" for k in MN if tot[k]) + "\n")
cols = [k for k in MN if tot[k] and k not in ("MUFU", "REDG", "ATOMG")]
print("Kernels that use tcgen05 / TMEM / TMA (one row per template instance):\n")
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for n, c in per.items():
    if any(c[k] for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM")):
        print(f"| `{short(n)}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
print("\nOther kernels with packed fp32 FMA (`FFMA2`) or global reductions:\n")
print("| kernel | instr | FFMA2 | REDG | ATOMG | MUFU |")
print("|---|---|---|---|---|---|")
for n, c in per.items():
    if not any(c[k] for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM")) and (c["FFMA2"] or c["REDG"] or c["ATOMG"]):
        print(f"| `{short(n)}` | {c['_n']} | {c['FFMA2'] or ''} | {c['REDG'] or ''} | {c['ATOMG'] or ''} | {c['MUFU'] or ''} |")
