"""Read an `ncu --set full` report here (no GPU needed) and write the per-launch DRAM traffic of the selected kernel as
JSON for bench.py (`roofline.traffic`), plus a markdown table of the usual metrics.

  python tools/ncu_extract.py gpurun_out/r02_ka_full.ncu-rep profiles/r02_ka_traffic.json [kernel-substring]
"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
           "sm__cycles_elapsed.max.per_second", "smsp__inst_executed.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and pat not in name:
            continue
        d = {"kernel": name[:120], "grid": r[col["Grid Size"]] if "Grid Size" in col else None}
        for m in METRICS:
            if m in col:
                d[m] = (r[col[m]], units[col[m]])
        launches.append(d)
    if not launches:
        raise SystemExit("no matching launches")

    def to_bytes(v):
        val, unit = v
        f = float(val.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    big = max(launches, key=lambda d: float(d["gpu__time_duration.sum"][0].replace(",", "")))
    traffic = to_bytes(big["dram__bytes_read.sum"]) + to_bytes(big["dram__bytes_write.sum"])
    js = {"kernel": big["kernel"], "launch": "largest captured launch (by duration)", "source": rep.split("/")[-1],
          "dram_bytes_read": to_bytes(big["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(big["dram__bytes_write.sum"]),
          "dram_bytes": traffic, "duration_us_under_ncu": float(big["gpu__time_duration.sum"][0].replace(",", "")),
          "all_launches": [{k: (v if isinstance(v, (str, type(None))) else f"{v[0]} {v[1]}") for k, v in d.items()} for d in launches]}
    json.dump(js, open(out, "w"), indent=1)
    print(json.dumps({k: js[k] for k in ("kernel", "dram_bytes", "duration_us_under_ncu")}))


if __name__ == "__main__":
    main()
