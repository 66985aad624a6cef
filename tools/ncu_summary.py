#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per captured kernel the headline metrics + executed-instruction mix.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch_json]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active"]
STALL_PREFIX, STALL_SUFFIX = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)}
        res.append(d)
    return res


def mix(path, idx):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    iA = hdr.index("Address")
    seen = set()
    for r in rows[hi + 1:]:
        if len(r) <= iE:
            continue
        if r[iA] in seen:          # the CSV lists every instruction once per source view: count each address once
            continue
        seen.add(r[iA])
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
        if not m:
            continue
        full = m.group(2)
        op = full.split(".")[0]
        key = full if op in ("MUFU", "F2F", "I2F", "F2I") else op
        try:
            ops[key] += int(r[iE])
        except ValueError:
            pass
    return ops


def main():
    path = sys.argv[1]
    kernels = raw(path)
    for idx, d in enumerate(kernels):
        name = d.get("Kernel Name", ("?", ""))[0]
        print("=" * 110)
        print(f"[{idx}] {name[:150]}")
        for k in KEYS:
            if k in d:
                print(f"    {k:75s} {d[k][0]:>18s} {d[k][1]}")
        st = sorted(((float(v[0]), k[len(STALL_PREFIX):-len(STALL_SUFFIX)]) for k, v in d.items()
                     if k.startswith(STALL_PREFIX) and k.endswith(STALL_SUFFIX) and v[0] not in ("", "n/a")), reverse=True)
        if st:
            print("    warp stall reasons per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
            print("        " + ", ".join(f"{k} {v:.2f}" for v, k in st if v >= 0.005))
        try:
            ops = mix(path, idx)
            tot = sum(ops.values())
            print(f"    executed warp-instructions: {tot}")
            print("    mix: " + ", ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in ops.most_common(22)))
        except Exception as e:       # noqa: BLE001
            print("    (no source page)", e)


if __name__ == "__main__":
    main()
