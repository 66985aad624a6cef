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
    print(f"executed warp-instructions (whole kernel): {tot}; most executed instruction: {mx}")
    print(f"hot path (>= {frac} of max): {n:.1f} instructions per trip, FP64 {fp64:.1f}, other {n - fp64:.1f}; "
          f"issue-cost model {cost:.0f} cycles per trip")
    print("ops: " + ", ".join(f"{k} {v:.1f}" for k, v in ops.most_common(40)))
    ts = sum(st.values())
    print("stall samples on the hot path: " + ", ".join(f"{k[6:]} {100.0 * v / ts:.1f}%" for k, v in st.most_common(12)))
    if "--list" in sys.argv:
        for r in hot:
            print(f"{int(r[iE]) / mx:5.2f} {r[iN]:>5s} {r[iS][:100]}")

main()
