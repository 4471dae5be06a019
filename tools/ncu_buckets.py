#!/usr/bin/env python3
"""Executed warp-instructions of an .ncu-rep bucketed by source-line ranges: file:lo-hi=name ..."""
import csv, subprocess, sys
def main(path, specs):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = None; lines = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0].isdigit() and len(r) > 8 and r[7] not in ("-", ""):
            try: lines.append((cur, int(r[0]), int(r[7]), int(r[6] or 0)))
            except ValueError: pass
    tot = sum(l[2] for l in lines); tots = sum(l[3] for l in lines)
    buckets = []
    for sp in specs:
        rng, name = sp.split("=")
        f, lh = rng.split(":"); lo, hi = map(int, lh.split("-"))
        buckets.append((f, lo, hi, name))
    acc = {b[3]: [0, 0] for b in buckets}; acc["other"] = [0, 0]
    for f, ln, n, smp in lines:
        for bf, lo, hi, name in buckets:
            if f == bf and lo <= ln <= hi:
                acc[name][0] += n; acc[name][1] += smp; break
        else:
            acc["other"][0] += n; acc["other"][1] += smp
    print(f"total warp-instructions {tot:,}; samples {tots:,}")
    for k, (n, smp) in acc.items():
        print(f"{k:28s} {n/tot*100:5.1f}% inst  {smp/max(tots,1)*100:5.1f}% samples")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
