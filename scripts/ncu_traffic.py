"""profiles/traffic.json from an `ncu --set full` capture of one step of bench.py (one stream group), stamped with the sha of
the kernel sources it was captured from so that bench.py refuses to quote it for other kernels.

usage: python scripts/ncu_traffic.py gpurun_out/<capture>.ncu-rep [profiles/traffic.json]

Stage mapping (launch order inside one frame step, `dvfe_tracker::submit`): the k_pyr_* launches before the first k_lk_track
are the left pyramid, those after it the right pyramid (stage "pyramid" = both); the first k_lk_track is the temporal call, the
second the stereo call."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import kernel_sources_sha
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "traffic.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                      "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                      "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
                                      "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
                                      "sm__warps_active.avg.pct_of_peak_sustained_active")}
    units = rows[1]

    def num(r, k):
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    # one entry per captured launch, in capture order
    L = [(r[col["Kernel Name"]].split("(")[0], r) for r in rows[2:] if len(r) > col["gpu__time_duration.sum"]]
    names = [n for n, _ in L]

    def bytes_of(i):
        return num(L[i][1], "dram__bytes_read.sum") + num(L[i][1], "dram__bytes_write.sum")

    def info(i):
        r = L[i][1]
        return {"sm_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                "l1tex_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                "warp_inst": num(r, "smsp__inst_executed.sum"), "time_us": num(r, "gpu__time_duration.sum"),
                "issue_active_pct": num(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}

    # the steady-state frame step: the temporal LK call is the k_lk_track followed by k_gftt_response (k_compact_tracked is not
    # captured), the stereo call the one right after the right camera's k_pyr_down launches; take the last complete step
    stages, sm_l1 = {}, {}
    t_idx = [i for i in range(len(L) - 2) if names[i] == "k_lk_track" and names[i + 1] == "k_gftt_response" and names[i + 2] == "k_gftt_select"]
    if not t_idx:
        raise SystemExit("no steady-state step (k_lk_track, k_gftt_response, k_gftt_select) in the capture")
    t = sidx = None
    for cand in reversed(t_idx):              # the last step the capture holds completely
        nxt = [i for i in range(cand + 3, len(L)) if names[i] == "k_lk_track"]
        if nxt and cand >= 3 and names[cand - 3:cand] == ["k_pyr_down"] * 3 and names[nxt[0] - 3:nxt[0]] == ["k_pyr_down"] * 3:
            t, sidx = cand, nxt[0]
            break
    if t is None:
        raise SystemExit("the capture does not hold one whole steady-state step")
    stages["lk_temporal"], sm_l1["lk_temporal"] = bytes_of(t), info(t)
    stages["gftt_response"], sm_l1["gftt_response"] = bytes_of(t + 1), info(t + 1)
    stages["gftt_select"], sm_l1["gftt_select"] = bytes_of(t + 2), info(t + 2)
    stages["lk_stereo"], sm_l1["lk_stereo"] = bytes_of(sidx), info(sidx)
    stages["pyramid"] = sum(bytes_of(i) for i in list(range(t - 3, t)) + list(range(sidx - 3, sidx)))
    stages["_sm_l1"] = sm_l1
    stages["_kernel_sources_sha"] = kernel_sources_sha()
    stages["_source"] = ("%s (ncu --set full --clock-control none, one step of bench.py --groups 1, 64 streams): "
                         "dram__bytes_read.sum + dram__bytes_write.sum per launch; pyramid = the k_pyr_down launches of both cameras"
                         % os.path.basename(rep))
    json.dump(stages, open(out, "w"), indent=1)
    print(json.dumps(stages, indent=1))


if __name__ == "__main__":
    main()
