#!/usr/bin/env python
"""Group the SASS listing of an ncu source page (--print-source sass) into runs of instructions with the
same execution count (≈ basic blocks) and print, per run: #instr, warp-level executions, average active
threads, FP64-pipe share and stall samples.  Shows where the issue slots of a kernel go.
usage: ncu -i rep --page source --csv --print-source sass > x.csv; tools/ncu_sass_blocks.py x.csv [min_pct]"""
import csv
import sys

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")


def main():
    path = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
    ins = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            ins.append((r[0], r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]),
                        int(r[ci["Thread Instructions Executed"]]), int(r[ci["# Samples"]])))
        except ValueError:
            pass
    total = sum(x[2] for x in ins)
    total_fp64 = sum(x[2] for x in ins if x[1].split()[0].lstrip("@!P0123456789 ").startswith(FP64) or any(f in x[1].split()[0:2][-1] for f in FP64))
    total_samples = sum(x[4] for x in ins)
    print(f"total warp instr {total:.4g}; thread-instr/warp-instr {sum(x[3] for x in ins) / total:.2f}; stall samples {total_samples}")
    blocks, cur = [], None
    for idx, (addr, src, ex, tex, smp) in enumerate(ins):
        op = [t for t in src.split() if not t.startswith("@")][0]
        is64 = op.split(".")[0] in FP64
        if cur is None or ex != cur["ex"]:
            cur = dict(start=idx, addr=addr, ex=ex, n=0, tex=0, smp=0, n64=0, ops={})
            blocks.append(cur)
        cur["n"] += 1
        cur["tex"] += tex
        cur["smp"] += smp
        cur["n64"] += is64
        cur["ops"][op.split(".")[0]] = cur["ops"].get(op.split(".")[0], 0) + 1
    fp64_total = sum(b["n64"] * b["ex"] for b in blocks)
    print(f"FP64-pipe warp instr {fp64_total:.4g} ({100 * fp64_total / total:.1f} % of issue slots)")
    print(f"{'idx':>5} {'n':>4} {'exec(M)':>9} {'%slots':>6} {'thr':>5} {'fp64':>4} {'%stall':>6}  top ops")
    for b in blocks:
        share = 100.0 * b["n"] * b["ex"] / total
        if share < min_pct:
            continue
        thr = b["tex"] / max(1, b["n"] * b["ex"])
        ops = sorted(b["ops"].items(), key=lambda kv: -kv[1])[:6]
        print(f"{b['start']:5d} {b['n']:4d} {b['ex'] / 1e6:9.1f} {share:6.2f} {thr:5.1f} {b['n64']:4d} {100.0 * b['smp'] / max(1, total_samples):6.2f}  "
              + " ".join(f"{k}:{v}" for k, v in ops))


if __name__ == "__main__":
    main()
