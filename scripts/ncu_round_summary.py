"""Tables for profiles/ncu_<round>_summary.md from the two captures of a round:
   python scripts/ncu_round_summary.py profiles/launches_r2.csv gpurun_out/r2s2_full.ncu-rep
prints (1) per-pass kernel shares of the launch list (stage-split pass = one stream group, value pass = 4 groups, e2e pass = one
group), (2) the --set full rows of the last complete steady-state frame step, (3) issue-active and stall shares of its kernels."""
import csv
import subprocess
import sys
from collections import defaultdict


def launch_tables(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    seq = []
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        v = float(r[hdr.index("Metric Value")].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}[r[hdr.index("Metric Unit")]]
        seq.append((r[hdr.index("Kernel Name")].split("(")[0], v, r[hdr.index("Grid Size")]))

    def gz(g):
        return int(g.strip("()").split(",")[2])
    zs = sorted({gz(g) for n, v, g in seq if n == "k_pyr_down"})
    small = zs[0]
    first = next(i for i, (n, v, g) in enumerate(seq) if n.startswith("k_pyr") and gz(g) == small)
    last = max(i for i, (n, v, g) in enumerate(seq) if n == "k_right_post_pack" and gz(g) == 1 and i > first and
               any(m == "k_lk_track" and gz2 == "(100, %d, 1)" % small for m, _, gz2 in seq[max(0, i - 3):i])) + 1
    passes = [("Stage-split / roofline pass (one stream group, every launch covers all streams)", seq[:first]),
              ("Value pass (%d streams per launch: the stream groups; live they overlap)" % small, seq[first:last]),
              ("E2E pass (one group; level 0 arrives by DMA, no k_pyr_level0)", seq[last:])]
    print("%d launches in all: %s\n" % (len(seq), ", ".join(sorted({n for n, _, _ in seq}))))
    for title, S in passes:
        d = defaultdict(lambda: [0, 0.0])
        for n, v, g in S:
            d[n][0] += 1
            d[n][1] += v
        tot = sum(v[1] for v in d.values())
        print(f"{title}: {len(S)} launches, {tot:.0f} us serialised\n")
        print("| kernel | launches | sum us | avg us | share |\n|---|---|---|---|---|")
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
            print(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[1] / tot:.3f} |")
        print()


def full_table(rep):
    out = subprocess.run([sys.executable, "scripts/ncu_summary.py", rep], capture_output=True, text=True).stdout.splitlines()
    hdr, rows = out[:2], out[2:]
    names = [r.split("|")[1].strip() for r in rows]
    t = None
    for c in reversed([i for i in range(3, len(rows) - 2) if names[i] == "k_lk_track" and names[i + 1] == "k_gftt_response"]):
        nxt = [i for i in range(c + 3, len(rows)) if names[i] == "k_lk_track"]
        if nxt and names[c - 3:c] == ["k_pyr_down"] * 3:
            t, s = c, nxt[0]
            break
    print("\n".join(hdr))
    print("\n".join(rows[t - 3:s + 1]))


if __name__ == "__main__":
    launch_tables(sys.argv[1])
    full_table(sys.argv[2])
