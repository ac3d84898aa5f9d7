"""Turns ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r01 gpurun_out/prof_r01_train.ncu-rep gpurun_out/prof_r01_predict.ncu-rep
    python profiles/summarize.py --launches r01 gpurun_out/launches_r01.csv
"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "launch__waves_per_multiprocessor"]


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    if sys.argv[1] == "--launches":
        tag, path = sys.argv[2], sys.argv[3]
        agg = {}
        lines = [l for l in open(path) if l.startswith('"')]
        rd = csv.reader(lines)
        hdr = next(rd)
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rd:
            name = r[ki].split("(")[0].replace("void ", "").replace("fmwr::", "")
            v = float(r[vi].replace(",", ""))
            if r[ui] in ("ns", "nsecond"): v /= 1e3
            elif r[ui] in ("ms", "msecond"): v *= 1e3
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1; a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open("profiles/%s_launches.md" % tag, "w") as f:
            f.write("# %s — launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n\n" % tag)
            f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
            for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (name, n, us, us / n, 100 * us / tot))
        print(open("profiles/%s_launches.md" % tag).read())
        return
    tag, paths = sys.argv[1], sys.argv[2:]
    with open("profiles/%s_ncu_full.csv" % tag, "w", newline="") as f:
        w = csv.writer(f)
        first = True
        for path in paths:
            hdr, units, rows = raw_rows(path)
            idx = [hdr.index(k) for k in KEEP if k in hdr]
            w.writerow(["# report", path])
            w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx]); first = False
            for r in rows:
                w.writerow([r[i] for i in idx])
    print(open("profiles/%s_ncu_full.csv" % tag).read())


if __name__ == "__main__":
    main()
