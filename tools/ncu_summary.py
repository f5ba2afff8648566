#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_name.md ["title"] [--traffic profiles/rNN_traffic.json]

--traffic also writes dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel (what bench.py reports as
roofline.traffic).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "fp64 tensor %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    argv = list(sys.argv)
    traffic_out = None
    if "--traffic" in argv:
        k = argv.index("--traffic")
        traffic_out = argv[k + 1]
        del argv[k:k + 2]
    rep, out = argv[1], argv[2]
    title = argv[3] if len(argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"source: `{rep}` (`ncu --set full --clock-control none`), one row per captured launch", "",
             "| kernel | " + " | ".join(n for _, n in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for m, _ in METRICS:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("n/a")
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic_out:
        import json

        ks = {}
        for r in data:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[m]].replace(",", "")) * UNIT_BYTES.get(units[idx[m]], 1.0)
            e = ks.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches_captured": 0})
            e["dram_bytes_per_launch"] += tot
            e["launches_captured"] += 1
        for e in ks.values():
            e["dram_bytes_per_launch"] /= e["launches_captured"]
        json.dump({"source": f"{rep} (ncu --set full --clock-control none, tools/profile_once.py, B=256 chains 382x84): "
                             "dram__bytes_read.sum + dram__bytes_write.sum per launch", "kernels": ks}, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main()
