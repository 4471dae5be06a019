#!/usr/bin/env python3
"""Per source line: executed warp-instructions, average active threads, stall samples (from the SASS rows of an .ncu-rep)."""
import csv, subprocess, sys
from collections import defaultdict
def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = None; cur_line = None; src = {}
    agg = defaultdict(lambda: [0, 0, 0])
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] in ("Function Name", "Line No", "File Name"): continue
        if r[0].isdigit():
            cur_line = (cur_file, int(r[0])); src[cur_line] = r[1].strip()[:90]; continue
        if r[0] == "" and len(r) > 9 and cur_line:
            try:
                agg[cur_line][0] += int(r[7]); agg[cur_line][1] += int(r[8]); agg[cur_line][2] += int(r[6] or 0)
            except ValueError: pass
    tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values()); tthr = sum(v[1] for v in agg.values())
    print(f"total warp-instructions {tot:,}; avg active threads {tthr/max(tot,1):.1f}; samples {tots:,}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{v[0]/tot*100:5.1f}% inst  {v[1]/max(v[0],1):5.1f} thr  {v[2]/max(tots,1)*100:5.1f}% smp  {k[0]}:{k[1]:<4d} {src[k]}")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
