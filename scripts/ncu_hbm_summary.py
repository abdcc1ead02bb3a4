"""Summarise `ncu -i X.ncu-rep --page raw --csv` of the HBM / L2-side kernels: per launch duration, DRAM bytes and achieved DRAM GB/s,
L2 bytes and achieved L2 GB/s, issue and occupancy figures.  Usage: python scripts/ncu_hbm_summary.py raw.csv [hbm_peak_gbs]"""
import csv, json, os, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6547.0
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


def scale(name, v):   # to bytes / ns
    u = units[col[name]] if name in col else ""
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(u, 1.0)


print(f"# DRAM peak used for the fractions: {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs, else the recipe's fallback)")
print(f"{'kernel':44s} {'grid':>7s} {'us':>7s} {'DRAM MB':>8s} {'DRAM GB/s':>9s} {'frac':>5s} {'L2 MB':>7s} {'L2 GB/s':>8s} {'L2 %pk':>6s} {'L2 hit%':>7s} {'warps%':>6s} {'issue%':>6s}")
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("fb::", "")[:44]
    t_ns = scale("gpu__time_duration.sum", num(r, "gpu__time_duration.sum"))
    dram = scale("dram__bytes_read.sum", num(r, "dram__bytes_read.sum")) + scale("dram__bytes_write.sum", num(r, "dram__bytes_write.sum"))
    l2 = 32.0 * num(r, "lts__t_sectors.sum")                      # L2 sectors of 32 bytes (tex + fabric side)
    l2pct = num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    hit = num(r, "lts__t_sector_hit_rate.pct")
    grid = r[col["Grid Size"]] if "Grid Size" in col else ""
    warps = num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    issue = num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", -1)
    issue = num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", issue)
    print(f"{name:44s} {grid:>7s} {t_ns / 1e3:7.1f} {dram / 1e6:8.1f} {dram / t_ns:9.0f} {dram / t_ns / peak:5.2f} {l2 / 1e6:7.1f} {l2 / t_ns:8.0f} {l2pct:6.1f} {hit:7.1f} {warps:6.1f} {issue:6.1f}")
