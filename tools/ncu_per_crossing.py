#!/usr/bin/env python
"""Per-crossing figures of one `ncu --set full` capture: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and FP64
thread instructions (DADD + DMUL + DFMA of the SASS page) divided by the crossings of the captured launch, which the bench
run under ncu reports (counters.pushes of a --steps 1 run = the captured, timed launch).
Usage: tools/ncu_per_crossing.py KEY REP BENCH_LOG [KEY REP BENCH_LOG ...] -> merges into profiles/r02_ncu_per_crossing.json"""
import json, re, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import ncu_summary as ns

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles" / "r02_ncu_per_crossing.json"


def num(s):
    return float(s.replace(",", ""))


def entry(rep, log):
    line = [l for l in Path(log).read_text().splitlines() if l.startswith('{"metric"')][-1]
    b = json.loads(line)
    assert b["steps"] == 1, "the capture must be of a --steps 1 run"
    pushes = b["counters"]["pushes"]
    r = ns.raw(rep)
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    dram = sum(num(r[k][0]) * unit_scale[r[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    scale = 1.0
    rows = ns.source(rep)
    cols = rows[0].keys()
    c_thr = next(c for c in cols if c.startswith("Thread Instructions Executed"))
    c_inst = next(c for c in cols if c.startswith("# Instructions Executed") or c == "Instructions Executed")
    fp64 = dsetp = warp = thr = 0
    for row in rows:
        try:
            t = int(row[c_thr].replace(",", "")); i = int(row[c_inst].replace(",", ""))
        except ValueError:
            continue
        warp += i; thr += t
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", row["Source"])
        op = m.group(1) if m else "?"
        if op in ("DADD", "DMUL", "DFMA"): fp64 += t
        if op == "DSETP": dsetp += t
    t_ms = num(r["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["gpu__time_duration.sum"][1], 1e-6)
    return {"capture": Path(rep).name, "workload": b["config"]["workload"], "crossings_in_capture": pushes,
            "dram_bytes_per_crossing": dram * scale / pushes, "fp64_thread_inst_per_crossing": fp64 / pushes,
            "dsetp_thread_inst_per_crossing": dsetp / pushes, "warp_inst_per_crossing": warp / pushes,
            "lanes_per_warp_inst": thr / max(warp, 1), "kernel_ms_under_ncu": t_ms,
            "l2_hit_pct": num(r["lts__t_sector_hit_rate.pct"][0]) if "lts__t_sector_hit_rate.pct" in r else None}


def main():
    a = sys.argv[1:]
    try:
        tab = json.loads(OUT.read_text())
    except Exception:
        tab = {"_note": "per-crossing figures from one `ncu --set full --clock-control none` capture per kernel AND workload "
                        "(key '<config.workload>:<kernel key>'); see tools/ncu_per_crossing.py.  DRAM bytes = "
                        "dram__bytes_read.sum + dram__bytes_write.sum of the captured launch / its crossings."}
    for k in range(0, len(a), 3):
        tab[a[k]] = entry(a[k + 1], a[k + 2])
        print(a[k], json.dumps(tab[a[k]]))
    OUT.write_text(json.dumps(tab, indent=1))


if __name__ == "__main__":
    main()
