"""Aggregate the source page of an ncu report (ncu -i X.ncu-rep --page source --csv) between the CTA barriers
of the channel-bank kernel's tile loop: instructions executed, stall samples and top stall reasons per phase."""
import csv, sys, collections

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
seg, segs = 0, collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    s = segs.setdefault(seg, dict(n=0, inst=0, samples=0, st=collections.Counter(), ops=collections.Counter(), first=r[col["Address"]], top=[]))
    s["n"] += 1
    s["inst"] += int(float(r[col["Instructions Executed"]] or 0))
    smp = int(float(r[col["# Samples"]] or 0))
    s["samples"] += smp
    for c in stall_cols:
        v = r[col[c]]
        if v:
            s["st"][c[6:]] += int(float(v))
    s["ops"][src.split()[0] if not src.startswith("@") else src.split()[1]] += 1
    s["top"].append((smp, src.strip()[:70]))
    if "BAR.SYNC" in src or "SYNCS.PHASECHK" in src:
        seg += 1
tot = sum(s["samples"] for s in segs.values())
tin = sum(s["inst"] for s in segs.values())
print(f"total samples {tot}, instructions executed {tin}")
for k, s in segs.items():
    top = ", ".join(f"{n} {v}" for n, v in s["st"].most_common(6))
    print(f"seg {k:2d} @{s['first']:>6s}: {s['n']:4d} instr  exec {s['inst']:>10d} ({100*s['inst']/max(tin,1):4.1f} %)  samples {s['samples']:6d} ({100*s['samples']/max(tot,1):4.1f} %)  {top}")
    if len(sys.argv) > 2:
        for smp, src in sorted(s["top"], reverse=True)[:4]:
            print(f"          {smp:6d}  {src}")
