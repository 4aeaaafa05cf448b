"""Per-source-line instruction counts and stall samples of one kernel of an .ncu-rep (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py rep.ncu-rep <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():   # a source line with its aggregated counters
        ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        data.append((int(r[ie] or 0), int(r[smp] or 0), f"{fname}:{r[0]}", r[1].strip()[:110]))
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"kernel {kern}: {tot} warp instructions, {ts} stall samples")
for n, s, loc, text in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{n:10d} {100 * n / max(tot, 1):5.1f}%  samples {100 * s / max(ts, 1):5.1f}%  {loc}: {text}")
