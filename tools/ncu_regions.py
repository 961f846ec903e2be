"""Per-source-line warp instructions / stall samples of one kernel of an .ncu-rep (--import-source on).
    python tools/ncu_regions.py report.ncu-rep kernel_substring [N]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = hdr = fn = None
agg, smp, src = collections.Counter(), collections.Counter(), {}
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
        k = (cur, int(r[0]))
        agg[k] += n
        try:
            smp[k] += int(d.get("# Samples", "0") or 0)
        except ValueError:
            pass
        src[k] = r[1]
tot, stot = sum(agg.values()) or 1, sum(smp.values()) or 1
print(f"{kern}: warp instructions {tot}, samples {stot}")
for k, n in agg.most_common(top):
    print(f"{n / tot * 100:5.1f}% inst {smp[k] / stot * 100:5.1f}% smp  {k[0]}:{k[1]:<4d} {src[k].strip()[:105]}")
