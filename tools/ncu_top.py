"""Summarise an `ncu --page source --csv` dump: per kernel, top instructions by stall samples with stall reasons."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
sections = []
name = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        name = r[1]
    if r and r[0] == "Address":
        sections.append([name, r, []])
    elif sections and r and r[0] not in ("Kernel Name",):
        sections[-1][2].append(r)
def num(s):
    try: return int(float(s or 0))
    except ValueError: return 0
for name, hdr, data in sections:
    si = hdr.index("# Samples"); src = hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in data if len(r) > si]
    tot = sum(num(r[si]) for r in data) or 1
    print("==", name[:90], "total samples", tot)
    agg = {hdr[i]: sum(num(r[i]) for r in data) for i in stall_cols}
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -num(r[si]))[:topn]:
        reasons = sorted([(num(r[i]), hdr[i][6:]) for i in stall_cols], reverse=True)[:3]
        print("%6d %5.1f%%  %-72s %s" % (num(r[si]), 100.0 * num(r[si]) / tot, r[src][:72], reasons))
