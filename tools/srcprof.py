"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: top source lines by samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
print("total samples", tot, "total warp-instr", sum(l[4] for l in lines))
for f, ln, src, smp, inst in sorted(lines, key=lambda l: -l[3])[:top_n]:
    print(f"{f}:{ln:4d} samples={smp:5d} ({100*smp/max(tot,1):4.1f}%) inst={inst:7d}  {src[:100]}")
