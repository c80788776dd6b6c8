"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass`: the SASS instructions with the most stall samples and
the opcode mix (executed warp instructions per opcode class).  usage: ncu_sass_summary.py source.csv [top=40]"""
import csv
import sys
from collections import Counter


def main():
    fn = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(fn)))
    hdr = None
    recs = []
    for r in rows:
        if not r:
            continue
        if hdr is None and ("Source" in r or "# Samples" in " ".join(r) or "Sampling Data (All)" in " ".join(r)):
            hdr = r
            continue
        if hdr is not None and len(r) == len(hdr):
            recs.append(dict(zip(hdr, r)))
    if not recs:
        print("no source rows found; header candidates:", rows[:3])
        return
    def num(d, *keys):
        for k in keys:
            if k in d:
                try:
                    return float(d[k].replace(",", "") or 0)
                except ValueError:
                    return 0.0
        return 0.0
    skey = next((k for k in hdr if k.startswith("# Samples") or k.startswith("Sampling Data (All)")), None)
    ikey = next((k for k in hdr if k.startswith("Instructions Executed") or k == "# Instructions Executed"), None)
    src = "Source" if "Source" in hdr else hdr[1]
    tot_s = sum(num(d, skey) for d in recs) or 1.0
    tot_i = sum(num(d, ikey) for d in recs) or 1.0
    print(f"rows {len(recs)}  stall samples {tot_s:.0f}  warp instructions executed {tot_i:.0f}")
    print("--- top instructions by stall samples")
    for d in sorted(recs, key=lambda d: -num(d, skey))[:top]:
        print(f"{100 * num(d, skey) / tot_s:6.2f}%  inst {num(d, ikey):12.0f}  {d.get(src, '')[:110]}")
    mix = Counter()
    for d in recs:
        op = (d.get(src, "").strip().split() or ["?"])
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else op[0]
        mix[op.split(".")[0]] += num(d, ikey)
    print("--- opcode mix (executed warp instructions)")
    for op, c in mix.most_common(30):
        print(f"{100 * c / tot_i:6.2f}%  {op}")


if __name__ == "__main__":
    main()
