"""Static issue-cycle estimate of the innermost loops of a kernel: for every backward branch, sum the stall fields of the
loop body and count FP32 packed instructions.  cycles/FP2 = 2.0 means a single warp could keep the FP32 pipe busy.
Usage: python scripts/sass_loop_cost.py lib.so kernel_substring"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, lines, ins = None, out.split("\n"), {}
i = 0
while i < len(lines):
    l = lines[i]
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", l)
    if m and cur and pat in cur:
        w1 = int(re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1]).group(1), 16)
        ins.setdefault(cur, []).append((int(m.group(1), 16), m.group(2).strip(), (w1 >> 41) & 15))
        i += 1
    i += 1
for fn, L in ins.items():
    print(fn[:100], "instructions:", len(L))
    for a, t, st in L:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in L if tgt <= x[0] <= a]
            fp2 = sum(1 for x in body if re.match(r"(@\S+\s+)?(FADD2|FFMA2|FMUL2)\b", x[1]))
            fp1 = sum(1 for x in body if re.match(r"(@\S+\s+)?(FADD|FFMA|FMUL)\b", x[1]))
            cyc = sum(max(x[2], 1) for x in body)
            if fp2 + fp1 > 16:
                print(f"   loop {tgt:05x}-{a:05x}: {len(body)} instr, {fp2} packed + {fp1} scalar FP, stall-sum {cyc} cycles"
                      f" -> {cyc / max(fp2 + 0.5 * fp1, 1):.2f} cycles per packed FP (2.0 = pipe-bound from one warp)")
