"""Condenses `ncu --page raw --csv` exports into (1) ncu_traffic.json -- DRAM bytes per launch of the dominant kernels, what
bench.py reports as `roofline.traffic` -- and (2) a small per-kernel summary CSV for profiles/.

    python tools/ncu_condense.py <saturated_raw.csv> <headline_raw.csv> <out.json> <summary.csv>
"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(unit, 1)


def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {n: (r[i], units[i]) for i, n in enumerate(names)}
        out.append(d)
    return out


def kernel(d):
    return d["Kernel Name"][0].split("(")[0]


def agg(launches):
    by = {}
    for d in launches:
        by.setdefault(kernel(d), []).append(d)
    res = {}
    for k, ds in by.items():
        rd = [to_bytes(*d["dram__bytes_read.sum"]) for d in ds]
        wr = [to_bytes(*d["dram__bytes_write.sum"]) for d in ds]
        ms = [to_ms(*d["gpu__time_duration.sum"]) for d in ds]
        res[k] = {"launches": len(ds), "dram_read_bytes_per_launch": sum(rd) / len(ds), "dram_write_bytes_per_launch": sum(wr) / len(ds),
                  "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(ds), "ms_per_launch_under_ncu": sum(ms) / len(ds),
                  "largest_launch_dram_bytes": max(a + b for a, b in zip(rd, wr))}
    return res


def main():
    sat, head, out_json, out_csv = sys.argv[1:5]
    s, h = load(sat), load(head)
    sa, ha = agg(s), agg(h)
    doc = {"source": "ncu --set full --clock-control none (tools/ncu_traffic.sh); dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "saturated": {"pairs": 131072, "algorithmic_input_bytes": 131072 * 192, "line_table_bytes": 131072 * 29120, "kernels": sa},
           "headline": {"n": 4096, "algorithmic_input_bytes_first_launch": 4096 * 192, "kernels": ha}}
    json.dump(doc, open(out_json, "w"), indent=1, sort_keys=True)
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["capture", "kernel", "launch"] + KEEP)
        for tag, ls in (("saturated", s), ("headline", h)):
            seen = {}
            for d in ls:
                k = kernel(d)
                seen[k] = seen.get(k, 0) + 1
                if seen[k] > 3:
                    continue
                w.writerow([tag, k, seen[k]] + [(d[m][0] + " " + d[m][1]).strip() if m in d else "" for m in KEEP])


if __name__ == "__main__":
    main()
