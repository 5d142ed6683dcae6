"""Dump SASS of one kernel with decoded scheduling control (stall count, yield, barriers) -- sm_90+/sm_100 encoding:
high 64-bit word bits 41..44 stall, 45 yield, 46..48 write-barrier, 49..51 read-barrier, 52..57 wait mask.
Usage: python scripts/sass_sched.py lib.so kernel_substring [start_addr_hex end_addr_hex]"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, lines = None, out.split("\n")
i = 0
while i < len(lines):
    l = lines[i]
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", l)
    if m and cur and pat in cur:
        addr, txt, w0 = int(m.group(1), 16), m.group(2), int(m.group(3), 16)
        m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        w1 = int(m2.group(1), 16)
        stall, yld = (w1 >> 41) & 15, (w1 >> 45) & 1
        wb, rb, wm = (w1 >> 46) & 7, (w1 >> 49) & 7, (w1 >> 52) & 63
        if lo <= addr <= hi:
            print(f"{addr:05x} st={stall:2d} y={yld} wb={wb if wb != 7 else '-'} rb={rb if rb != 7 else '-'} wait={wm:06b}  {txt}")
        i += 1
    i += 1
