#!/usr/bin/env python
"""Static view of a kernel's main loop from cuobjdump SASS (no GPU): the largest backward-branch span, instruction count and
opcode histogram.  usage: python tools/sass_loop.py OBJ MANGLED_NAME_SUBSTRING"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", names)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+((?:@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]*)([^;]*);", f)
    addr = [int(a, 16) for a, _, _ in ins]
    # the hot loop: the SHORTEST backward branch whose span holds at least 40 FP64 instructions
    best = None
    for i, (a, op, rest) in enumerate(ins):
        if op.split()[-1].startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m:
                t = int(m.group(1), 16)
                if t < int(a, 16):
                    nf = sum(1 for (b, o, _) in ins if t <= int(b, 16) <= int(a, 16) and o.split()[-1][:4] in ("DFMA", "DMUL", "DADD"))
                    if nf >= 40 and (best is None or int(a, 16) - t < best[0]):
                        best = (int(a, 16) - t, t, int(a, 16))
    if best is None:
        continue
    body = [(a, op) for (a, op, _) in ins if best[1] <= int(a, 16) <= best[2]]
    ops = collections.Counter(op.split()[-1].split(".")[0] if not op.split()[-1].startswith("IMAD.WIDE") else "IMAD.WIDE" for _, op in body)
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(name[:110])
    print(f"  loop 0x{best[1]:x}..0x{best[2]:x}: {len(body)} instructions (static count of the innermost FP64 loop; side paths inside the span included), FP64 {fp64}")
    print("  " + ", ".join(f"{k} {v}" for k, v in ops.most_common(30)))
