"""Per-region stall breakdown of one kernel from an .ncu-rep source page.
Usage: python scripts/ncu_hot.py file.ncu-rep kernel_regex [top_n]
Prints total samples by stall reason, the share of samples inside the FP-dense loop bodies (instructions that are
FADD2/FFMA2/LDS.64/LDC inside backward-branch ranges) vs. outside, and the top-N instructions by samples."""
import csv
import io
import re
import subprocess
import sys


def main():
    path, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    blocks = re.split(r'(?m)^"Kernel Name",', out)
    for blk in blocks[1:2]:
        lines = blk.split("\n")
        print("kernel:", lines[0][:110])
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: 0 for h in reasons}
        ins = []
        for r in data:
            s = int(r[idx["# Samples"]] or 0)
            ex = int(r[idx["Instructions Executed"]] or 0)
            for h in reasons:
                tot[h] += int(r[idx[h]] or 0)
            ins.append((s, ex, r[idx["Source"]].strip(), {h: int(r[idx[h]] or 0) for h in reasons}))
        total = sum(x[0] for x in ins)
        print("total samples", total, "instructions", sum(x[1] for x in ins))
        for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
            print(f"   {h:24s} {v:8d} {100.0 * v / max(total, 1):5.1f}%")
        fp = [x for x in ins if re.match(r"(@\S+\s+)?(FADD2|FFMA2|FMUL2|FADD|FFMA|FMUL)\b", x[2])]
        print("FP-instruction samples: %.1f%% ; executed FP warp-instr %d of %d (%.1f%%)" % (
            100.0 * sum(x[0] for x in fp) / max(total, 1), sum(x[1] for x in fp), sum(x[1] for x in ins),
            100.0 * sum(x[1] for x in fp) / max(sum(x[1] for x in ins), 1)))
        print("top instructions by samples:")
        for s, ex, src, st in sorted(ins, key=lambda x: -x[0])[:top]:
            main_r = sorted(st.items(), key=lambda kv: -kv[1])[:2]
            print(f"   {s:7d} ({100.0 * s / max(total, 1):4.1f}%) ex={ex:9d}  {src[:60]:60s} {main_r}")


if __name__ == "__main__":
    main()
