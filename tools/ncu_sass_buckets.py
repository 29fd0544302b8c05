import csv, io, subprocess, sys, collections
rep, sub = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
blk = next(b for b in blocks if sub in b["name"])
hdr = blk["rows"][0]; ix = {h: i for i, h in enumerate(hdr)}
rows = blk["rows"][1:]
tot = sum(int(r[ix["# Samples"]]) for r in rows)
print("total samples", tot)
for b0 in range(0, len(rows), B):
    seg = rows[b0:b0 + B]
    smp = sum(int(r[ix["# Samples"]]) for r in seg)
    ex = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    ops = collections.Counter()
    for r in seg:
        p = r[ix["Source"]].split()
        op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
        ops[op] += 1
    top = max(seg, key=lambda r: int(r[ix["# Samples"]]))
    print(f"{b0:5d}  samples {smp:5d} ({100*smp/tot:5.1f}%)  exec/instr {ex//max(1,len(seg)):7d}  {' '.join(f'{k}:{v}' for k,v in ops.most_common(4)):40s} | top {int(top[ix['# Samples']])} {top[ix['Source']].strip()[:50]}")
