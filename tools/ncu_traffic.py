#!/usr/bin/env python3
"""Read `ncu --set full` captures of the solve kernel and write profiles/r02_traffic.json + one CSV summary per capture.

Usage (here, no GPU needed):  tools/ncu_traffic.py ecdsa=gpurun_out/r02_ecdsa.ncu-rep tiled16=gpurun_out/r02_tiled16.ncu-rep

bench.py reads `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of one k_solve launch) from the JSON,
so the figure in the bench line is the one of the committed capture and changes when the capture does.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum(\.per_second)?|gpu__time_duration\.sum|lts__t_sectors(_op_(read|write|atom|red))?\.sum|"
    r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|l1tex__t_sectors_pipe_lsu_mem_global_op_ld\.sum|"
    r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.per_cycle_active|"
    r"launch__(registers_per_thread|block_size|grid_size|shared_mem_per_block_dynamic)|sm__cycles_elapsed\.max|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_(barrier|long_scoreboard|short_scoreboard|lg_throttle|membar|wait)_per_issue_active\.ratio|"
    r"smsp__inst_executed\.sum|dram__cycles_active\.avg\.pct_of_peak_sustained_elapsed)$")


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    return [dict(zip(names, r)) for r in rows[2:]], dict(zip(names, units))


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * m[unit]


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    jpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    table = json.load(open(jpath)) if os.path.exists(jpath) else {}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        tag = "r02"
        if "@" in key:
            key, tag = key.split("@", 1)
        launches, units = raw_page(rep)
        solve = [l for l in launches if "k_solve" in l.get("Kernel Name", "")]
        if not solve:
            print("no k_solve launch in", rep)
            continue
        l = solve[-1]
        rd = to_bytes(l["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
        wr = to_bytes(l["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
        summary = os.path.join("profiles", "%s_k_solve_%s_ncu.csv" % (tag, key))
        with open(os.path.join(ROOT, summary), "w") as f:
            f.write("# ncu --set full --clock-control none, last k_solve launch of %s (%s)\n" % (os.path.basename(rep), l["Kernel Name"]))
            f.write("metric,unit,value\n")
            for n in sorted(l):
                if KEEP.match(n):
                    f.write("%s,%s,%s\n" % (n, units.get(n, ""), l[n]))
        table[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                      "ncu_ms": float(l["gpu__time_duration.sum"].replace(",", "")) *
                      {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units["gpu__time_duration.sum"]],
                      "source": summary, "grid": l.get("launch__grid_size"), "block": l.get("launch__block_size"),
                      "registers": l.get("launch__registers_per_thread")}
        print(key, json.dumps(table[key]))
    json.dump(table, open(jpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
