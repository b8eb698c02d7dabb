"""Summarise `ncu --page source --csv` output: per-kernel stall-sample totals grouped by instruction class
and the top sampled instructions.  Usage: python tools/ncu_source_summary.py report.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; rx = sys.argv[2] if len(sys.argv) > 2 else "."
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', out)
for blk in blocks[1:]:
    lines = blk.splitlines()
    name = lines[0].strip('",')
    if not re.search(rx, name):
        continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]; rows = [r for r in rows[1:] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ci["# Samples"]] or 0) for r in rows)
    print(f"=== {name[:90]}  total samples {tot}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ci[h]] or 0) for r in rows) for h in stall_cols}
    print("  stall totals:", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(tot, 1) >= 1))
    cls = {}
    for r in rows:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        cls[op] = cls.get(op, 0) + int(r[ci["# Samples"]] or 0)
    print("  by opcode:", ", ".join(f"{k}={v * 100 // max(tot, 1)}%" for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:12]))
    top = sorted(rows, key=lambda r: -int(r[ci["# Samples"]] or 0))[:14]
    for r in top:
        st = sorted(((h[6:], int(r[ci[h]] or 0)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"   {int(r[ci['# Samples']]):7d} {r[ci['Address']][-5:]} {r[ci['Source']][:70]:70s} {st}")
