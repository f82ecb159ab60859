#!/usr/bin/env python
"""Condense an `ncu --set full` report of one rendered frame into the small JSON bench.py loads for `roofline.traffic`.

    ncu --set full --clock-control none --kernel-name regex:k_anyhit --launch-skip 12 -c 12 -o rep \
        python tools/gpu_one_frame.py c4_open 3840 2160 2 > frame.log
    python tools/ncu_capture_to_json.py rep.ncu-rep k_anyhit frame.log profiles/r02_ncu_k_anyhit_c4_open.json

The kernel's launches of the captured frame are summed (DRAM bytes read + written, duration, warp instructions) and divided by
the rays those launches traversed, which the frame log states (`ao traversed` + shadow rays are not separated by ncu: the
figure is per any-hit ray of the frame).  Numbers under ncu are cold-cache and serialised: they give TRAFFIC and instruction
counts, never a speed."""
import csv
import io
import json
import re
import subprocess
import sys


def main():
    rep, kernel, log, out = sys.argv[1:5]
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(cols)], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    name_i = head.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    launches = []
    for r in rows[2:]:
        if kernel not in r[name_i]:
            continue
        d = {}
        for c in cols:
            if c in head:
                i = head.index(c)
                try:
                    d[c] = float(r[i]) * scale.get(units[i], 1.0)
                except ValueError:
                    pass
        launches.append(d)
    text = open(log).read()
    m = re.findall(r"rays (\d+) \(ao (\d+) traversed (\d+)\)", text)
    sh = re.findall(r"shadow traversed (\d+)", text)
    if not m:
        raise SystemExit("no frame line in %s" % log)
    ao_trav = int(m[-1][2])
    shadow_trav = int(sh[-1]) if sh else 0
    rays = ao_trav + shadow_trav
    tot = lambda c: sum(l.get(c, 0.0) for l in launches)
    dram = tot("dram__bytes_read.sum") + tot("dram__bytes_write.sum")
    inst = tot("smsp__inst_executed.sum")
    res = {
        "source": "ncu --set full --clock-control none, %d launches of %s in one frame (%s); tools/ncu_capture_to_json.py" % (len(launches), kernel, " ".join(sys.argv[1:4])),
        "kernel": kernel, "launches": len(launches), "rays": rays, "ao_rays_traversed": ao_trav, "shadow_rays_traversed": shadow_trav,
        "dram_bytes": dram, "dram_bytes_per_ray": dram / rays if rays else None,
        "counters": {
            "duration_ms_under_ncu": tot("gpu__time_duration.sum"),
            "warp_instructions": inst, "warp_instructions_per_ray": inst / rays if rays else None,
            "active_lanes_per_instruction": sum(l.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * l.get("smsp__inst_executed.sum", 0) for l in launches) / inst if inst else None,
            "issue_active_pct": sum(l.get("smsp__issue_active.avg.pct", 0) * l.get("gpu__time_duration.sum", 0) for l in launches) / max(tot("gpu__time_duration.sum"), 1e-9),
            "l1_hit_pct": sum(l.get("l1tex__t_sector_hit_rate.pct", 0) * l.get("gpu__time_duration.sum", 0) for l in launches) / max(tot("gpu__time_duration.sum"), 1e-9),
            "l2_hit_pct": sum(l.get("lts__t_sector_hit_rate.pct", 0) * l.get("gpu__time_duration.sum", 0) for l in launches) / max(tot("gpu__time_duration.sum"), 1e-9),
            "warps_active_pct": sum(l.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) * l.get("gpu__time_duration.sum", 0) for l in launches) / max(tot("gpu__time_duration.sum"), 1e-9),
            "l2_bytes": tot("lts__t_bytes.sum"), "l1_bytes": tot("l1tex__t_bytes.sum"),
            "registers_per_thread": launches[0].get("launch__registers_per_thread") if launches else None,
        },
        "per_launch": launches,
    }
    res["counters"] = {k: v for k, v in res["counters"].items() if v}      # (metrics this ncu version does not have read as 0)
    with open(out, "w") as f:
        f.write(json.dumps(res) + "\n")          # one line, like the bench lines next to it
    print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}, indent=1))


if __name__ == "__main__":
    main()
