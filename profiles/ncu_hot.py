"""Summarise an `ncu --page source --csv` dump: warp-stall samples per SASS instruction, grouped into
address ranges so that loops show up as regions.  usage: python profiles/ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# a dump may hold several kernels ("Kernel Name" line + header + instructions each): take the first one
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(ends) > 1:
    rows = rows[:ends[1]]
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = int(r[ci["# Samples"]])
    st = max(stalls, key=lambda h: int(r[ci[h]]))
    data.append((int(r[ci["Address"]], 16), r[ci["Source"]].strip(), s, int(r[ci["Instructions Executed"]]), st))
tot = sum(d[2] for d in data)
base = data[0][0]
print("total samples", tot, "instructions", len(data))
print("---- hottest instructions")
for a, src, s, ex, st in sorted(data, key=lambda t: -t[2])[:top]:
    print(f"{s:8d} {100.0*s/tot:5.1f}%  +{a-base:05x}  exec {ex:10d}  {st:24s} {src[:70]}")
print("---- samples per 0x400-byte region")
reg = {}
for a, src, s, ex, st in data:
    reg.setdefault((a - base) // 0x400, [0, 0])
    reg[(a - base) // 0x400][0] += s
    reg[(a - base) // 0x400][1] += ex
for k in sorted(reg):
    if reg[k][0] > 0.01 * tot:
        print(f"+{k*0x400:05x}: {100.0*reg[k][0]/tot:5.1f}%  warp-instr {reg[k][1]}")
