"""Per-source-line stall samples / instruction counts of one kernel from an .ncu-rep (needs -lineinfo and --import-source on):
    python tools/ncu_source_lines.py REPORT.ncu-rep [TOP]"""
import csv, subprocess, sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    fname, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":   # a source line (its SASS rows carry an address)
            lines.append((fname, r))
    sa, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lsb = [i for i, h in enumerate(hdr) if h in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_lg", "stall_mio")]
    tot = sum(int(r[sa] or 0) for _, r in lines) or 1
    toti = sum(int(r[ie] or 0) for _, r in lines) or 1
    print("total samples", tot, "warp instructions", toti)
    for f, r in sorted(lines, key=lambda x: -int(x[1][sa] or 0))[:top]:
        st = " ".join(f"{hdr[i][6:]}={r[i]}" for i in lsb if r[i] not in ("0", ""))
        print(f"{int(r[sa] or 0) * 100 / tot:5.1f}% smp {int(r[ie] or 0) * 100 / toti:5.1f}% ins {f}:{r[0]:>4s} | {r[1].strip()[:100]} | {st}")


main()
