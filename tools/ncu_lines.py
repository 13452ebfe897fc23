#!/usr/bin/env python
"""Per-CUDA-source-line issue-slot shares from an ncu report captured with --import-source on.
usage: tools/ncu_lines.py rep.ncu-rep [min_pct]  (uses `ncu --page source --print-source cuda,sass`)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, agg = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                agg.append((fname, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]),
                            int(d["Thread Instructions Executed"]), int(d["# Samples"]), int(d["stall_barrier"]),
                            int(d["stall_wait"]), int(d["stall_short_sb"]), int(d["stall_long_sb"]), int(d["stall_math"])))
            except (ValueError, KeyError):
                pass
    tot = sum(a[3] for a in agg)
    tsm = sum(a[5] for a in agg)
    print(f"total warp instr {tot:.4g}, samples {tsm}")
    print(f"{'file:line':28s} {'%slots':>6} {'thr':>5} {'%smp':>5} {'bar':>5} {'wait':>5} {'ssb':>5} {'lsb':>5} {'math':>5}  source")
    for a in sorted(agg, key=lambda a: (a[0], a[1])):
        pct = 100.0 * a[3] / tot
        if pct < min_pct:
            continue
        s = max(1, tsm)
        print(f"{a[0][:22] + ':' + str(a[1]):28s} {pct:6.2f} {a[4] / max(1, a[3]):5.1f} {100 * a[5] / s:5.2f} "
              f"{100 * a[6] / s:5.2f} {100 * a[7] / s:5.2f} {100 * a[8] / s:5.2f} {100 * a[9] / s:5.2f} {100 * a[10] / s:5.2f}  {a[2]}")


if __name__ == "__main__":
    main()
