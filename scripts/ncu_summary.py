"""Per-kernel summary of an `ncu --page raw --csv` export: duration, DRAM bytes, throughputs, occupancy, top stalls.
    python scripts/ncu_summary.py gpurun_out/r2_ncu_stage_raw.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "MB_rd", 1e-6), ("dram__bytes_write.sum", "MB_wr", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1), ("launch__registers_per_thread", "regs", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1), ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
        ("smsp__inst_executed.sum", "Minst", 1e-6)]
STALL = "smsp__average_warps_issue_stalled_"      # ..._per_issue_active.ratio (warp cycles per issued instruction)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    out = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[col["Kernel Name"]]
        def val(n):
            try:
                v = float(r[col[n]].replace(",", ""))
            except Exception:
                return None
            u = units[col[n]]
            if n.endswith("time_duration.sum"):
                v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
            if "bytes" in n:
                v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            return v
        d = {lab: (val(n)*sc if (n in col and val(n) is not None) else None) for n, lab, sc in WANT}
        stalls = sorted(((float(r[i].replace(",", "")), names[i][len(STALL):].split("_per_")[0]) for i, n in enumerate(names)
                         if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")), reverse=True)[:3]
        d["stalls"] = " ".join("%s=%.1f" % (s, v) for v, s in stalls)
        out.setdefault(k, []).append(d)
    print("%-34s %3s %8s %7s %7s %6s %5s %6s %5s %5s %6s %6s %7s  %s" % ("kernel", "n", "us", "MB_rd", "MB_wr", "dram%", "sm%", "fp64%", "occ%", "regs", "issue%", "l2hit%", "Minst", "top stalls (cycles per issued instruction)"))
    for k, ds in out.items():
        def avg(lab):
            v = [d[lab] for d in ds if d[lab] is not None]
            return sum(v)/len(v) if v else float("nan")
        print("%-34s %3d %8.1f %7.1f %7.1f %6.1f %5.1f %6.1f %5.1f %5.0f %6.1f %6.1f %7.2f  %s" % (
            k[:34], len(ds), avg("us"), avg("MB_rd"), avg("MB_wr"), avg("dram%"), avg("sm%"), avg("fp64%"), avg("occ%"), avg("regs"),
            avg("issue%"), avg("l2hit%"), avg("Minst"), ds[0]["stalls"]))


if __name__ == "__main__":
    main(sys.argv[1])
