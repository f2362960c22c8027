#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics, stall reasons, and a per-function SASS breakdown
(instructions executed, thread efficiency, stall samples).  Usage: tools/ncu_summary.py REP [--top N]"""
import csv, io, subprocess, sys, re, collections

def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout

def raw(rep):
    out = run([rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}

def source(rep):
    out = run([rep, "--page", "source", "--csv", "--print-source", "sass"])
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or '"Source"' in l)
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

KEYS = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    r = raw(rep)
    for k in KEYS:
        if k in r: print(f"{k:75s} {r[k][0]:>18s} {r[k][1]}")
    st = [(float(v[0].replace(",", "")), k) for k, v in r.items()
          if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and v[0] not in ("", "n/a")]
    for v, k in sorted(st, reverse=True)[:10]:
        print(f"  stalled warps per issue: {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}")
    rows = source(rep)
    cols = rows[0].keys()
    c_inst = next(c for c in cols if c.startswith("# Instructions Executed") or c == "Instructions Executed")
    c_thr = next(c for c in cols if c.startswith("Thread Instructions Executed"))
    c_smp = next(c for c in cols if c.startswith("Warp Stall Sampling (All"))
    c_src = "Source"
    tot_i = tot_t = tot_s = 0
    by_op = collections.defaultdict(lambda: [0, 0, 0])
    for row in rows:
        try:
            i = int(row[c_inst].replace(",", "")); t = int(row[c_thr].replace(",", "")); s = int(row[c_smp].replace(",", ""))
        except ValueError:
            continue
        tot_i += i; tot_t += t; tot_s += s
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", row[c_src])
        op = m.group(1) if m else "?"
        b = by_op[op]; b[0] += i; b[1] += t; b[2] += s
    print(f"SASS: warp-inst {tot_i:.3e}  thread-inst {tot_t:.3e}  lanes/inst {tot_t/max(tot_i,1):.2f}  samples {tot_s}")
    print(f"{'op':12s} {'warp-inst%':>10s} {'lanes':>6s} {'samples%':>9s}")
    for op, (i, t, s) in sorted(by_op.items(), key=lambda kv: -kv[1][2])[:25]:
        print(f"{op:12s} {100*i/tot_i:10.2f} {t/max(i,1):6.1f} {100*s/max(tot_s,1):9.2f}")
    if top:
        print("top SASS lines by samples:")
        idx = sorted(range(len(rows)), key=lambda k: -int((rows[k][c_smp] or "0").replace(",", "") or 0))[:top]
        for k in idx:
            row = rows[k]
            print(f"  {k:6d} {row[c_smp]:>8s} {row[c_inst]:>12s} {row[c_src][:90]}")

if __name__ == "__main__":
    main()
