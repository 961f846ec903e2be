"""Warp instructions / stall samples of one kernel of an .ncu-rep between source markers (developer tool).
    python tools/ncu_marks.py report.ncu-rep kernel_substring file.cu tiles 'name=pattern' ..."""
import collections, csv, subprocess, sys
rep, kern, path, tiles = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = hdr = fn = None
agg, smp = collections.Counter(), collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "Function Name":
        fn = r[1]
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and kern in (fn or ""):
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        agg[(cur, int(r[0]))] += n
        try:
            smp[(cur, int(r[0]))] += int(d.get("# Samples", "0") or 0)
        except ValueError:
            pass
tot, stot = sum(agg.values()), sum(smp.values())
lines = open(path).read().split("\n")
base = path.split("/")[-1]
marks = []
for m in sys.argv[5:]:
    name, pat = m.split("=", 1)
    marks.append((name, next(i + 1 for i, l in enumerate(lines) if pat in l)))
marks.append(("end", len(lines) + 1))
print(f"{kern}: {tot} warp instructions ({tot / tiles:.0f} per tile), {stot} samples")
def rng(a, b):
    return (sum(n for (c, l), n in agg.items() if c == base and a <= l < b), sum(n for (c, l), n in smp.items() if c == base and a <= l < b))
i, s_ = rng(1, marks[0][1])
print(f"{'(before)':14s} {i / tot * 100:5.1f}% inst ({i / tiles:7.0f}/tile) {s_ / stot * 100:5.1f}% samples")
for (n1, a), (n2, b) in zip(marks, marks[1:]):
    i, s_ = rng(a, b)
    print(f"{n1:14s} {i / tot * 100:5.1f}% inst ({i / tiles:7.0f}/tile) {s_ / stot * 100:5.1f}% samples")
o = sum(n for (c, l), n in agg.items() if c != base)
print(f"{'other files':14s} {o / tot * 100:5.1f}% inst ({o / tiles:7.0f}/tile)")
