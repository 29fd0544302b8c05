"""Per-instruction executed counts / stall samples from an ncu report (source page, SASS view).

    python tools/ncu_sass_hot.py REPORT.ncu-rep KERNEL_SUBSTR [top]
Prints, for the first launch whose name contains KERNEL_SUBSTR, the executed-instruction total, the share of SYNCS
(mbarrier try_wait) / spin instructions, an opcode histogram weighted by executions, and the top stall-sample rows.
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and row:
            cur["rows"].append(row)
    blk = next(b for b in blocks if sub in b["name"])
    hdr = blk["rows"][0]
    ix = {h: i for i, h in enumerate(hdr)}
    rows = blk["rows"][1:]
    tot = sum(int(r[ix["Instructions Executed"]]) for r in rows)
    samp = sum(int(r[ix["# Samples"]]) for r in rows)
    print(blk["name"])
    print(f"warp instructions executed {tot}, stall samples {samp}, static SASS {len(rows)}")
    hist, shist = collections.Counter(), collections.Counter()
    for r in rows:
        src = r[ix["Source"]].strip()
        parts = src.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        hist[op] += int(r[ix["Instructions Executed"]])
        shist[op] += int(r[ix["# Samples"]])
    print("opcode: executed share | stall-sample share")
    for op, n in hist.most_common(22):
        print(f"  {op:10s} {100.0 * n / tot:6.2f}%   {100.0 * shist[op] / max(1, samp):6.2f}%")
    print("top stall rows:")
    for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        print(f"  {int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
