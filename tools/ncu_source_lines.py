#!/usr/bin/env python3
"""Per-source-line executed-instruction counts and stall samples from an .ncu-rep (needs -lineinfo)."""
import csv, subprocess, sys
def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file = None; lines = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] in ("Function Name", "Line No", "File Name"): continue
        if r[0].isdigit() and len(r) > 8 and r[7] not in ("-", ""):
            try: lines.append((cur_file, int(r[0]), r[1].strip()[:100], int(r[7]), int(r[6] or 0), float(r[10] or 0)))
            except ValueError: pass
    tot = sum(l[3] for l in lines); tots = sum(l[4] for l in lines)
    print(f"total warp-instructions attributed: {tot:,}  samples: {tots:,}")
    for f, ln, src, n, smp, thr in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{n/tot*100:5.1f}% inst {smp/max(tots,1)*100:5.1f}% smp thr={thr:4.1f} {f}:{ln:<4d} {src}")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
