"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py report.ncu-rep kernel_regex [top]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
            try:
                ie = int(r[hdr.index("Instructions Executed")])
                te = int(r[hdr.index("Thread Instructions Executed")])
                smp = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            lines.append((ie, te, smp, fname, r[0], r[1].strip()[:110]))
    tot = sum(l[0] for l in lines) or 1
    tots = sum(l[2] for l in lines) or 1
    print(f"total warp instructions attributed: {tot}, samples {tots}")
    for ie, te, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100.0 * ie / tot:5.1f}% inst {100.0 * smp / tots:5.1f}% samp  lanes {te / max(ie, 1):4.1f}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()
