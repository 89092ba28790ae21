"""Per-source-line instruction/stall breakdown from an .ncu-rep captured with --import-source on (-lineinfo build):
   python profiles/top_lines.py gpurun_out/prof.ncu-rep [N]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, topn=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[2] == "-":  # a source line row (SASS rows carry an address)
            d = dict(zip(hdr[4:], r[4:]))
            try:
                ie = int(d.get("Instructions Executed", "0") or 0)
                smp = int(d.get("# Samples", "0") or 0)
            except ValueError:
                continue
            data.append((cur, int(r[0]), r[1].strip()[:100], ie, smp))
    tot = sum(x[3] for x in data) or 1
    tots = sum(x[4] for x in data) or 1
    print("total warp instructions", tot, "samples", tots)
    pf = defaultdict(lambda: [0, 0])
    for f, l, s, ie, smp in data:
        pf[f][0] += ie
        pf[f][1] += smp
    for f, v in sorted(pf.items(), key=lambda kv: -kv[1][0]):
        print("%-18s %5.1f%% inst %5.1f%% samples" % (f, 100 * v[0] / tot, 100 * v[1] / tots))
    print("--- top lines by instructions executed (inst%, stall-sample%)")
    for f, l, s, ie, smp in sorted(data, key=lambda x: -x[3])[:topn]:
        print("%-16s %4d %5.1f%% %5.1f%%  %s" % (f, l, 100 * ie / tot, 100 * smp / tots, s))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
