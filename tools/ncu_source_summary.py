"""Summarise an `ncu -i rep --page source --csv` dump (optionally .gz): executed warp instructions by opcode, and the warp-stall samples by reason
and by the opcode they were taken on.  python tools/ncu_source_summary.py gpurun_out/x_source.csv.gz [top]"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
fh = gzip.open(path, "rt", errors="replace") if path.endswith(".gz") else open(path, errors="replace")
rows = list(csv.reader(fh))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
by_op_inst, by_op_samp, by_reason = collections.Counter(), collections.Counter(), collections.Counter()
tot_inst = tot_samp = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)?)", r[ix["Source"]])
    op = m.group(1) if m else "?"
    op = op.split(".")[0] if not op.startswith(("MUFU", "UTC", "UTMA", "SYNCS", "LDTM", "LDG", "STS", "LDS")) else op
    n = int(float(r[ix["Instructions Executed"]] or 0)); sm = int(float(r[ix["# Samples"]] or 0))
    by_op_inst[op] += n; by_op_samp[op] += sm; tot_inst += n; tot_samp += sm
    for c in stall_cols:
        v = r[ix[c]]
        if v not in ("", "0"):
            by_reason[c] += int(float(v))
print(f"{path}: {tot_inst} warp instructions executed, {tot_samp} stall samples\n")
print("| opcode | warp instr | share | stall samples | share |\n|---|---:|---:|---:|---:|")
for op, n in by_op_inst.most_common(top):
    print(f"| {op} | {n} | {100 * n / tot_inst:.1f}% | {by_op_samp[op]} | {100 * by_op_samp[op] / max(1, tot_samp):.1f}% |")
print("\n| stall reason (all samples) | samples | share |\n|---|---:|---:|")
tr = sum(by_reason.values())
for c, n in by_reason.most_common(10):
    print(f"| {c} | {n} | {100 * n / max(1, tr):.1f}% |")
