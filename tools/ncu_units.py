"""Summarise `ncu --set full` raw CSV exports (ncu -i X.ncu-rep --page raw --csv) into profiles/r2_ncu_units.json:
per workload and kernel, what binds it (issue slots / L1 / L2 / HBM as % of each unit's own peak), DRAM and L2 bytes,
SIMD efficiency and occupancy. bench.py reads the k_extend<EXT_MAIN> entry of its workload for `roofline.bound`.

    python tools/ncu_units.py workload=csv_path[:kernel_regex] ...   (merges into the existing JSON)
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_ncu_units.json")
PEAK_HBM = 6545.3  # MEASURED_PEAKS.json of this pool


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def scale(v, unit, kind):
    """ncu picks units per row: normalise bytes to bytes and time to microseconds."""
    if v is None:
        return None
    u = unit.lower()
    if kind == "bytes":
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    if kind == "time":
        return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(u, 1)
    return v


def summarise(path, kernel_re=None):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, kind=None):
        i = col.get(name)
        if i is None:  # section-prefixed duplicates, e.g. "FBSP.TriageCompute.dram__throughput..."
            hits = [j for h, j in col.items() if h.endswith("." + name)]
            i = hits[0] if hits else None
        return None if i is None else scale(fnum(r[i]), units[i], kind)
    out = []
    for r in data:
        name = r[col["Kernel Name"]]
        if kernel_re and not re.search(kernel_re, name):
            continue
        dur = get(r, "gpu__time_duration.sum", "time")
        rd, wr = get(r, "dram__bytes_read.sum", "bytes"), get(r, "dram__bytes_write.sum", "bytes")
        lts = get(r, "lts__t_bytes.sum", "bytes")
        if lts is None and get(r, "lts__t_sectors.sum") is not None:
            lts = 32.0 * get(r, "lts__t_sectors.sum")
        dram_pct = get(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
        if dram_pct is None:  # read + write shares of the DRAM peak
            a, b = get(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed"), get(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed")
            dram_pct = None if a is None or b is None else a + b
        e = {
            "kernel": re.sub(r"^.*?(k_\w+(<[^>]*>)?).*$", r"\1", name),
            "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "duration_us": dur,
            "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_throughput_pct": get(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
            "lts_throughput_pct": get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": dram_pct,
            "dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
            "l2_bytes_per_launch": lts,
            "hbm_gbs": None if not dur or rd is None else (rd + wr) / dur / 1e3,
            "l2_gbs": None if not dur or lts is None else lts / dur / 1e3,
            "hbm_frac_of_measured_peak": None if not dur or rd is None else (rd + wr) / dur / 1e3 / PEAK_HBM,
            "active_lanes_per_inst": get(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "inst_executed": get(r, "smsp__inst_executed.sum"),
            "l1_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct"),
            "registers": get(r, "launch__registers_per_thread"),
            "source": "profiles/" + os.path.basename(path) + " (ncu --set full --clock-control none, one launch, cold caches, serialised)",
        }
        out.append(e)
    return out


def main():
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        path, _, kre = rest.partition(":")
        launches = summarise(path, kre or None)
        if not launches:
            print("no kernel matched in", path)
            continue
        # the longest matching launch represents the kernel (bounce 0 launches dominate)
        best = max(launches, key=lambda e: e["duration_us"] or 0)
        best["launches_in_capture"] = len(launches)
        db[key] = best
        print(key, best["kernel"], "%.1f us" % best["duration_us"], {k: best[k] for k in ("issue_active_pct", "l1tex_throughput_pct", "lts_throughput_pct", "dram_throughput_pct", "active_lanes_per_inst", "warps_active_pct")})
    json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
