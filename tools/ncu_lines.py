"""Per source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].strip().isdigit():
            si = hdr["# Samples"]
            if len(r) > si and r[si].isdigit() and int(r[si]) > 0:
                why = sorted(((int(r[i]), h[6:]) for h, i in hdr.items()
                              if h.startswith("stall_") and "(" not in h and i < len(r) and r[i].isdigit()
                              and int(r[i]) > 0), reverse=True)[:3]
                lines.append((int(r[si]), fpath, int(r[0]), r[1].strip()[:90], r[hdr["Instructions Executed"]], why))
    tot = sum(l[0] for l in lines)
    print("total samples", tot)
    for s, f, ln, src, ex, why in sorted(lines, reverse=True)[:top]:
        print("%5d %5.1f%% %-20s:%-4d exec=%-8s %-90s %s" % (s, 100.0 * s / tot, f, ln, ex, src, why))


if __name__ == "__main__":
    main()
