#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel in an .ncu-rep (read on the CPU box).

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX CUBIN [min_pct]

ncu's source page gives per-SASS-instruction counters; nvdisasm -g gives the line of every SASS instruction
of the same cubin; the two are joined by instruction order."""
import csv, io, re, subprocess, sys
rep, kre, cubin = sys.argv[1:4]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if not rows:
    print("no kernel matching", kre, "in", rep); sys.exit(1)
# several captured launches of the same kernel: keep the first table only
for _i in range(1, len(rows)):
    if rows[_i] and rows[_i][0] == "Kernel Name":
        rows = rows[:_i]
        break
kname = rows[0][1]
hdr = rows[1]
ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sass = [(r[isrc].strip(), int(r[ie]), int(r[isamp]), [int(r[i]) for i in stall_cols]) for r in rows[2:] if len(r) > ie]
# mangled name of the kernel
m = re.search(r"(\w+)<", kname) or re.search(r"(\w+)\(", kname)
base = m.group(1).split("::")[-1]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split by function
funcs, cur, name = {}, None, None
for line in dis.splitlines():
    mm = re.match(r"\s*\.text\.(\S+):", line)
    if mm:
        name = mm.group(1); cur = []; funcs[name] = cur; ln = None; continue
    if cur is None: continue
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if mm: ln = (mm.group(1).split("/")[-1], int(mm.group(2))); continue
    mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if mm: cur.append((mm.group(1).strip(), ln))
cands = [k for k in funcs if base in k and len(funcs[k]) == len(sass)]
if not cands:
    print("no function with", len(sass), "instructions; have", {k: len(v) for k, v in funcs.items() if base in k}); sys.exit(1)
f = funcs[cands[0]]
tot = sum(s[1] for s in sass); tots = sum(s[2] for s in sass)
per = {}
for (txt, n, smp, st), (t2, ln) in zip(sass, f):
    d = per.setdefault(ln, [0, 0, [0] * len(stall_cols)])
    d[0] += n; d[1] += smp
    for i, v in enumerate(st): d[2][i] += v
print(f"{kname}: {tot} warp instructions, {tots} samples, {len(sass)} SASS instructions")
srcs = {}
for ln, (n, smp, st) in sorted(per.items(), key=lambda kv: (kv[0] or ("", 0))):
    if n / tot * 100 < min_pct and smp / max(1, tots) * 100 < min_pct: continue
    top = sorted(zip(st, [hdr[i][6:] for i in stall_cols]), reverse=True)[:2]
    text = ""
    if ln:
        try:
            if ln[0] not in srcs:
                import glob
                p = glob.glob(f"/root/repo/wild-gaussians_b200/csrc/{ln[0]}")
                srcs[ln[0]] = open(p[0]).read().splitlines() if p else []
            text = srcs[ln[0]][ln[1] - 1].strip()[:90]
        except Exception: pass
    print(f"{(ln[0] + ':' + str(ln[1])) if ln else '?':28s} inst {n / tot * 100:5.1f}%  samp {smp / max(1, tots) * 100:5.1f}%  {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}  | {text}")
