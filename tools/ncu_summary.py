"""Reads the ncu artefacts of a round (gpurun_out/<tag>_*) here, without a GPU, and writes the summaries that are committed under
profiles/:
  <tag>_ncu_kernels.md         one row per `--set full` capture: duration, DRAM bytes and % of peak, L2 hit rate, occupancy,
                               registers, issue-active %, top stall reasons (warps per issue)
  <tag>_c2_ncu_launches_summary.csv   per-kernel sums of the launch list (time, DRAM read/write bytes, launches)
  whole_step_dram.json         dram__bytes_read.sum + dram__bytes_write.sum over EVERY kernel of one build step
usage: python tools/ncu_summary.py <tag> [out_dir]"""
import csv
import glob
import gzip
import io
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_dir = sys.argv[2] if len(sys.argv) > 2 else "profiles"
src = "gpurun_out"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def raw_page(path):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    return {n: (u, v) for n, u, v in zip(names, units, vals)}


def num(m, key, scale_units=True):
    u, v = m.get(key, ("", ""))
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(u, 1.0) if scale_units else x


def kernel_rows():
    rows = []
    for path in sorted(glob.glob(os.path.join(src, f"{tag}_k_*_full.ncu-rep"))):
        m = raw_page(path)
        name = m.get("Kernel Name", ("", "?"))[1].split("(")[0]
        stalls = sorted(((float(v[1]), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "selected" not in k and v[1] not in ("", "n/a")), reverse=True)[:3]
        dur_us = num(m, "gpu__time_duration.sum")
        rd, wr = num(m, "dram__bytes_read.sum") or 0.0, num(m, "dram__bytes_write.sum") or 0.0
        rows.append({
            "kernel": name, "grid": m.get("launch__grid_size", ("", ""))[1], "block": m.get("launch__block_size", ("", ""))[1], "duration_us": dur_us,
            "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6, "dram_GBs": (rd + wr) / (dur_us * 1e-6) / 1e9 if dur_us else None,
            "dram_pct": num(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False), "sm_pct": num(m, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
            "l2_hit_pct": num(m, "lts__t_sector_hit_rate.pct", False), "occupancy_pct": num(m, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "regs": num(m, "launch__registers_per_thread", False), "issue_active_pct": num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "stalls": ", ".join(f"{n} {v:.1f}" for v, n in stalls), "file": os.path.basename(path),
        })
    return rows


def launch_summary():
    path = os.path.join(src, f"{tag}_c2_ncu_launches.csv.gz")
    if not os.path.exists(path):
        return None
    per = {}
    with gzip.open(path, "rt") as f:
        rows = csv.reader(f)
        hdr = None
        for r in rows:
            if r and r[0] == "ID":
                hdr = {n: i for i, n in enumerate(r)}
                continue
            if hdr is None or len(r) < len(hdr):
                continue
            name = re.sub(r"\(.*", "", r[hdr["Kernel Name"]])
            metric, unit, value = r[hdr["Metric Name"]], r[hdr["Metric Unit"]], float(r[hdr["Metric Value"]].replace(",", ""))
            e = per.setdefault(name, {"launches": 0, "us": 0.0, "read": 0.0, "write": 0.0})
            if metric == "gpu__time_duration.sum":
                e["us"] += value * UNIT.get(unit, 1.0)
                e["launches"] += 1
            elif metric == "dram__bytes_read.sum":
                e["read"] += value * UNIT.get(unit, 1.0)
            elif metric == "dram__bytes_write.sum":
                e["write"] += value * UNIT.get(unit, 1.0)
    return per


def main():
    os.makedirs(out_dir, exist_ok=True)
    rows = kernel_rows()
    if rows:
        with open(os.path.join(out_dir, f"{tag}_ncu_kernels.md"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, first (level-0, largest) launch of each kernel in one C2 build ({tag})\n\n")
            f.write("DRAM % is `gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed` (hardware peak); GB/s = (read + write) / duration, to be read against the measured copy peak in MEASURED_PEAKS.json. "
                    "Stalls are `smsp__average_warps_issue_stalled_*_per_issue_active` (warps waiting per issued instruction), top three.\n\n")
            f.write("| kernel | grid x block | us | DRAM read MB | write MB | GB/s | DRAM % | SM % | L2 hit % | occupancy % | regs | issue active % | top stalls |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| `{r['kernel']}` | {r['grid']} x {r['block']} | {r['duration_us']:.1f} | {r['dram_read_MB']:.1f} | {r['dram_write_MB']:.1f} | {r['dram_GBs']:.0f} | {r['dram_pct']:.1f} | {r['sm_pct']:.1f} | "
                        f"{r['l2_hit_pct']:.1f} | {r['occupancy_pct']:.1f} | {int(r['regs'])} | {r['issue_active_pct']:.1f} | {r['stalls']} |\n")
        print(open(os.path.join(out_dir, f"{tag}_ncu_kernels.md")).read())
    per = launch_summary()
    if per:
        total_us = sum(e["us"] for e in per.values())
        total_bytes = sum(e["read"] + e["write"] for e in per.values())
        with open(os.path.join(out_dir, f"{tag}_c2_ncu_launches_summary.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share", "dram_read_bytes", "dram_write_bytes"])
            for name, e in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
                w.writerow([name, e["launches"], f"{e['us']:.1f}", f"{e['us'] / total_us:.4f}", int(e["read"]), int(e["write"])])
        path = os.path.join(out_dir, "whole_step_dram.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data["C2"] = {"bytes": int(total_bytes), "read_bytes": int(sum(e["read"] for e in per.values())), "write_bytes": int(sum(e["write"] for e in per.values())), "kernel_launches": int(sum(e["launches"] for e in per.values())),
                      "kernel_time_us_under_ncu": round(total_us, 1), "bytes_per_input_triangle": round(total_bytes / 9999392, 1),
                      "source": f"profiles/{tag}_c2_ncu_launches.csv.gz: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one full C2 build (first build of the process), kernels only (memsets and copies are not kernels)"}
        json.dump(data, open(path, "w"), indent=1)
        print("whole step:", data["C2"])


if __name__ == "__main__":
    main()
