"""Stall hot spots of one kernel from an `ncu --set full --import-source on` report (run here, no GPU needed):
    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep profiles/ncu_<kernel>_hotspots.txt [launch-index]
Warp-stall samples per SASS instruction: the top instructions, and the totals per opcode and per stall reason."""
import collections, csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": next(csv.reader([line]))[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(line)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = blocks[which]
rows = list(csv.reader(io.StringIO("\n".join(b["rows"]))))
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in rows)
per_op, per_reason = collections.Counter(), collections.Counter()
items = []
for r in rows:
    n = int(r[ix["# Samples"]])
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    per_op[op.split(".")[0]] += n
    rs = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]])}
    for s, v in rs.items():
        per_reason[s] += v
    if n:
        items.append((n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), rs))
items.sort(key=lambda x: -x[0])
L = [f"warp-stall samples of {b['name']} ({rep}, launch {which}; ncu --set full --import-source on); {tot} samples, {len(rows)} SASS instructions",
     "", "per stall reason: " + ", ".join(f"{k[6:]} {v} ({100 * v / tot:.1f} %)" for k, v in per_reason.most_common(12)),
     "", "per opcode: " + ", ".join(f"{k} {v} ({100 * v / tot:.1f} %)" for k, v in per_op.most_common(16)),
     "", "samples  share  warp-instr  SASS                                                    top stall reasons"]
for n, src, ex, rs in items[:40]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    L.append(f"{n:7d} {100 * n / tot:5.1f} % {ex:10d}  {src[:54]:54s}  [{top}]")
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L[:14]))
