#!/usr/bin/env python
"""Hot-loop view of one kernel of an .ncu-rep (read here, no GPU): the SASS instructions executed at >= FRAC of the most
executed one, with per-instruction executed counts and sampled stall reasons, plus the op histogram of that hot path and
its cost under the measured issue costs (tools/issue_microbench.cu): FP64 2, IMAD.WIDE 4, MUFU 8, everything else 1.
usage: python tools/ncu_hot.py REP [launch_index] [frac] [--list]"""
import collections, csv, io, re, subprocess, sys

def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    frac = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    body = [r for r in rows[hi + 1:] if len(r) > iE and r[iE].isdigit()]
    # the CSV repeats the listing once per source view: keep the first copy
    seen, uniq = set(), []
    iA = h.index("Address")
    for r in body:
        if r[iA] in seen:
            continue
        seen.add(r[iA]); uniq.append(r)
    mx = max(int(r[iE]) for r in uniq)
    tot = sum(int(r[iE]) for r in uniq)
    hot = [r for r in uniq if int(r[iE]) >= frac * mx]
    ops = collections.Counter()
    st = collections.Counter()
    for r in hot:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
        full = m.group(2)
        op = "IMAD.WIDE" if full.startswith("IMAD.WIDE") else full.split(".")[0]
        ops[op] += int(r[iE]) / mx
        for k in stalls:
            v = r[h.index(k)]
            if v.isdigit():
                st[k] += int(v)
    n = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    cost = n + fp64 + 3 * ops.get("IMAD.WIDE", 0) + 7 * ops.get("MUFU", 0)
    if "--json" in sys.argv:                        # (--json: stdout is the JSON document alone, the summary goes to stderr)
        sys.stdout, real_stdout = sys.stderr, sys.stdout
    print(f"executed warp-instructions (whole kernel): {tot}; most executed instruction: {mx}")
    print(f"hot path (>= {frac} of max): {n:.1f} instructions per trip, FP64 {fp64:.1f}, other {n - fp64:.1f}; "
          f"issue-cost model {cost:.0f} cycles per trip")
    print("ops: " + ", ".join(f"{k} {v:.1f}" for k, v in ops.most_common(40)))
    ts = sum(st.values())
    print("stall samples on the hot path: " + ", ".join(f"{k[6:]} {100.0 * v / ts:.1f}%" for k, v in st.most_common(12)))
    if "--json" in sys.argv:
        # the per-kernel evidence file bench.py reads (profiles/chain_metrics_r2.json)
        import json
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        hd, row = rr[0], rr[2 + idx]
        g = lambda k: row[hd.index(k)] if k in hd else None          # noqa: E731
        fl = lambda k: float(g(k)) if g(k) not in (None, "", "n/a") else None   # noqa: E731
        unit = lambda k: rr[1][hd.index(k)] if k in hd else ""       # noqa: E731

        def to_bytes(k):
            v = fl(k)
            if v is None:
                return None
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit(k), 1.0)
        stl = sorted(((float(row[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                      for i, k in enumerate(hd) if k.startswith("smsp__average_warps_issue_stalled_")
                      and k.endswith("_per_issue_active.ratio") and row[i] not in ("", "n/a")), reverse=True)
        dur = fl("gpu__time_duration.sum")
        out = {"report": rep, "kernel": g("Kernel Name"), "instr_per_pair": round(n, 2), "fp64_instr_per_pair": round(fp64, 2),
               "imad_wide_per_pair": round(ops.get("IMAD.WIDE", 0), 2), "mufu_per_pair": round(ops.get("MUFU", 0), 2),
               "issue_cost_cycles_per_pair": round(cost, 1), "ops_per_pair": {k: round(v, 2) for k, v in ops.most_common()},
               "executed_warp_instructions": tot,
               "fp64_pipe_pct": fl("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "alu_pipe_pct": fl("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "fma_pipe_pct": fl("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "xu_pipe_pct": fl("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": fl("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": fl("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "registers_per_thread": fl("launch__registers_per_thread"),
               "dram_bytes_per_launch": (to_bytes("dram__bytes_read.sum") or 0) + (to_bytes("dram__bytes_write.sum") or 0),
               "duration_ms_under_ncu": dur / 1e3 if dur and unit("gpu__time_duration.sum") in ("us", "usecond") else dur,
               "local_memory_sectors": (fl("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") or 0) + (fl("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum") or 0),
               "top_stalls": {k: round(v, 2) for v, k in stl[:8]}}
        real_stdout.write(json.dumps(out, indent=1) + "\n")
        return
    if "--list" in sys.argv:
        for r in hot:
            print(f"{int(r[iE]) / mx:5.2f} {r[iN]:>5s} {r[iS][:100]}")

main()
