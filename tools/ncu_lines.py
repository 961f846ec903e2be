"""Top source lines of an .ncu-rep (captured with --import-source on) by warp instructions executed.
    python tools/ncu_lines.py report.ncu-rep [N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur, hdr = None, None
agg, samp, src = collections.Counter(), collections.Counter(), {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0] != "Function Name" and hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        key = (cur, int(r[0]))
        agg[key] += n
        try:
            samp[key] += int(d.get("# Samples", "0") or 0)
        except ValueError:
            pass
        src[key] = r[1]
tot, stot = sum(agg.values()) or 1, sum(samp.values()) or 1
print(f"total warp instructions {tot}, samples {stot}")
for key, n in agg.most_common(top):
    print(f"{n / tot * 100:5.1f}% inst {samp[key] / stot * 100:5.1f}% smp  {key[0]}:{key[1]:<4d} {src[key].strip()[:100]}")
