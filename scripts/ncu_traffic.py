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
                                      "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum")}
    units = rows[1]

    def num(r, k):
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    stages, sm_l1, n_lk = {}, {}, 0
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        if name == "k_lk_track":
            stage = "lk_temporal" if n_lk == 0 else "lk_stereo"
            n_lk += 1
            if n_lk > 2:
                break
        elif name.startswith("k_pyr"):
            stage = "pyramid"
        elif name.startswith("k_gftt_response"):
            stage = "gftt_response"
        elif name.startswith("k_gftt_select"):
            stage = "gftt_select"
        else:
            continue
        stages[stage] = stages.get(stage, 0.0) + num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
        if stage != "pyramid":
            sm_l1[stage] = {"sm_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                            "l1tex_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                            "warp_inst": num(r, "smsp__inst_executed.sum")}
    stages["_sm_l1"] = sm_l1
    stages["_kernel_sources_sha"] = kernel_sources_sha()
    stages["_source"] = ("%s (ncu --set full --clock-control none, one step of bench.py --groups 1, 64 streams): "
                         "dram__bytes_read.sum + dram__bytes_write.sum per launch; pyramid = the k_pyr_down launches of both cameras"
                         % os.path.basename(rep))
    json.dump(stages, open(out, "w"), indent=1)
    print(json.dumps(stages, indent=1))


if __name__ == "__main__":
    main()
