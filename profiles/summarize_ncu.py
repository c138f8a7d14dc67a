#!/usr/bin/env python3
"""Turn one `ncu --set full` report of a bench.py step into the files committed under profiles/:
    <tag>_<kernel>_kernel_ncu_summary.txt   the details page of each of the step's kernels (text)
    traffic.json                            DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum),
                                            read by bench.py for roofline.traffic
usage: python profiles/summarize_ncu.py gpurun_out/<report>.ncu-rep <tag> ["<how the report was captured>"]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
KERNELS = (("collide", "b2p_collide_kernel"), ("assemble", "b2p_assemble_kernel"), ("sor", "b2p_sor"),
           ("integrate", "b2p_integrate_kernel"))


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    how = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = num(r[idx[name]])
        u = units[idx[name]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
                 "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return v * scale if v is not None else None

    out = {"source": f"ncu --set full --clock-control none, report {os.path.basename(rep)} ({tag}); {how}", "kernels": {}}
    seen = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for key, pat in KERNELS:
            if pat in name and key not in seen:
                seen[key] = r[idx["ID"]]
                out["kernels"][key] = {
                    "kernel": name,
                    "dram_read_bytes": val(r, "dram__bytes_read.sum"),
                    "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                    "duration_s_under_ncu": val(r, "gpu__time_duration.sum"),
                    "registers": num(r[idx["launch__registers_per_thread"]]),
                    "grid": num(r[idx["launch__grid_size"]]), "block": num(r[idx["launch__block_size"]]),
                    "warps_active_per_scheduler": num(r[idx["smsp__warps_active.avg.per_cycle_active"]]),
                    "warps_eligible_per_scheduler": num(r[idx["smsp__warps_eligible.avg.per_cycle_active"]]),
                    "issue_active_pct": num(r[idx["smsp__issue_active.avg.pct"]]) if "smsp__issue_active.avg.pct" in idx else None,
                    "active_lanes_per_warp_inst": num(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                    "warp_instructions": num(r[idx["smsp__inst_executed.sum"]]),
                }
    k = out["kernels"]
    out["step_kernel_dram_bytes_per_launch"] = sum(k[n]["dram_read_bytes"] + k[n]["dram_write_bytes"]
                                                   for n in ("assemble", "sor", "integrate") if n in k)
    if "collide" in k:
        out["collide_dram_bytes_per_launch"] = k["collide"]["dram_read_bytes"] + k["collide"]["dram_write_bytes"]
    json.dump(out, open(os.path.join(ROOT, "traffic.json"), "w"), indent=1)
    for key, kid in seen.items():
        txt = subprocess.run(["ncu", "-i", rep, "--page", "details", "--kernel-id", f"::regex:{dict(KERNELS)[key]}:1"],
                             capture_output=True, text=True).stdout
        if not txt.strip():
            txt = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        # keep metric tables, drop the advisory prose
        keep = []
        for line in txt.splitlines():
            s = line.strip()
            if not s or s.startswith(("OPT", "INF", "---", "WRN")):
                continue
            keep.append(line.rstrip())
        open(os.path.join(ROOT, f"{tag}_{key}_kernel_ncu_summary.txt"), "w").write("\n".join(keep) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
