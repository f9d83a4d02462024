"""Top stall sites of an `ncu --set full --import-source on` report: python scripts/ncu_hot.py report.ncu-rep [N]
Prints the N SASS instructions with the most warp-stall samples (all samples), with their dominant stall reasons."""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[hdr_i:]))))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ci["Warp Stall Sampling (All Samples)"]])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print(f"# {path}: {tot} stall samples over {len(data)} instructions")
for s, r in sorted(data, key=lambda t: -t[0])[:top]:
    reasons = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    rs = ", ".join(f"{n}:{v}" for v, n in reasons if v)
    print(f"{100*s/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {rs}")
