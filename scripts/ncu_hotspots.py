"""Summarise the SASS page of an ncu report (made with --set full --import-source on) per kernel launch:
stall-reason totals, instruction mix by opcode, and the hottest SASS instructions by stall samples.
usage: ncu -i rep.ncu-rep --page source --csv > sass.csv ; python scripts/ncu_hotspots.py sass.csv [top_n]"""
import csv
import sys
from collections import Counter


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and len(r) > 5:
            cur["rows"].append(r)
    for k in kernels:
        h = {n: i for i, n in enumerate(k["hdr"])}
        def col(r, n):
            try:
                return float(r[h[n]] or 0)
            except ValueError:
                return 0.0
        tot_s = sum(col(r, "# Samples") for r in k["rows"])
        tot_i = sum(col(r, "Instructions Executed") for r in k["rows"])
        print(f"== {k['name'][:60]}  samples {tot_s:.0f}  warp instructions {tot_i:.0f}  SASS lines {len(k['rows'])}")
        stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        print("  stall share: " + ", ".join(f"{n[6:]} {100 * sum(col(r, n) for r in k['rows']) / max(tot_s, 1):.1f}%"
                                          for n in sorted(stalls, key=lambda n: -sum(col(r, n) for r in k["rows"]))[:8]))
        mix, smp = Counter(), Counter()
        for r in k["rows"]:
            op = r[h["Source"]].split()
            op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
            op = op.split(".")[0]
            mix[op] += col(r, "Instructions Executed")
            smp[op] += col(r, "# Samples")
        print("  instruction mix: " + ", ".join(f"{o} {100 * c / max(tot_i, 1):.1f}%" for o, c in mix.most_common(14)))
        print("  samples by opcode: " + ", ".join(f"{o} {100 * c / max(tot_s, 1):.1f}%" for o, c in smp.most_common(10)))
        hot = sorted(k["rows"], key=lambda r: -col(r, "# Samples"))[:top]
        for r in hot:
            top_stall = max(stalls, key=lambda n: col(r, n))
            print(f"    {col(r, '# Samples'):6.0f} smp  {col(r, 'Instructions Executed'):10.0f} exec  {top_stall[6:]:14s} {r[h['Source']][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
