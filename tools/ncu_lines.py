#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]

Prints the source lines with the most stall samples and executed warp instructions.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        d = dict(zip(range(len(hdr)), r))
        try:
            lines.append((int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        except ValueError:
            pass
    tot_s = sum(l[2] for l in lines) or 1
    tot_i = sum(l[3] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    for ln, src, smp, ins in sorted(lines, key=lambda l: -l[2])[:top]:
        print("%5d  %5.1f%% smp  %5.1f%% inst  %s" % (ln, 100.0 * smp / tot_s, 100.0 * ins / tot_i, src[:110]))


if __name__ == "__main__":
    main()
