# coding: utf-8
"""Summarise an `ncu --page raw --csv` export of the synthesis kernel into the small JSON bench.py reads
(profiles/r2_ncu_summary.json).   python scripts/summarize_ncu.py gpurun_out/r2_prof_raw.csv T B > profiles/r2_ncu_summary.json"""
import csv
import json
import sys


def main():
    path, T, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr = i
            break
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    col = {n: (v, u) for n, u, v in zip(names, units, vals)}

    def num(name, default=None):
        if name not in col:
            return default
        try:
            return float(col[name][0].replace(",", ""))
        except ValueError:
            return default

    def to_bytes(name):
        v = num(name)
        if v is None:
            return None
        u = col[name][1].lower()
        mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
        return v * mul

    dur = num("gpu__time_duration.sum")
    du = col.get("gpu__time_duration.sum", ("", "ns"))[1].lower()
    dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(du, 1e-6)
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    stalls = {}
    for n in names:
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
            key = n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            v = num(n)
            if v is not None:
                stalls[key] = round(v, 3)
    stalls = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    out = {
        "kernel": col.get("Kernel Name", ("?",))[0],
        "workload": "BASELINE config 2, B=%d" % B, "T": T, "utts_per_gpu": B,
        "capture": "ncu --set full --clock-control none --import-source on -k regex:wn_persistent -c 1 python scripts/ncu_target.py %d %d" % (T, B),
        "duration_ms": dur_ms,
        "dram_bytes_per_launch": (rd or 0) + (wr or 0),
        "dram_bytes_per_sample": ((rd or 0) + (wr or 0)) / T,
        "dram_read_GBps_during_capture": (rd or 0) / (dur_ms * 1e-3) / 1e9 if dur_ms else None,
        "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
        "dram_throughput_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_throughput_pct_of_peak": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct_of_peak": num("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": num("launch__registers_per_thread"),
        "issue_active_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") if False else num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_stalls_per_issue": stalls,
        "note": "launch durations under ncu are serialised / cold-cache: compare shares, not absolutes",
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
