"""Split a kernel's SASS (ncu --page source --csv) at BAR.SYNC instructions and report, per phase,
the share of executed warp instructions, lane utilisation and sampled stall share.
usage: python profiles/sass_phases.py <report.ncu-rep> <kernel regex>"""
import csv
import subprocess
import sys


def main(rep, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[1]
    ci, cs, ct, cp = h.index("Instructions Executed"), h.index("Source"), h.index("Thread Instructions Executed"), h.index("# Samples")
    data = []
    for r in rows[2:]:
        try:
            data.append((r[cs].strip(), int(r[ci]), int(r[ct]), int(r[cp])))
        except Exception:
            pass
    tot = sum(d[1] for d in data)
    tots = sum(d[3] for d in data)
    print(f"{rows[0][1]}: {tot} warp instructions (summed over SASS lines), {len(data)} SASS lines")
    phase, acc, first = 0, [0, 0, 0, 0], 0
    for k, (src, n, t, s) in enumerate(data):
        acc[0] += n; acc[1] += t; acc[2] += s; acc[3] += 1
        if "BAR.SYNC" in src or k == len(data) - 1:
            print(f"phase {phase:2d} sass[{first:4d}..{k:4d}] {100 * acc[0] / tot:5.1f}% of instr, "
                  f"{acc[1] / max(1, acc[0]):4.1f} thr/inst, {100 * acc[2] / max(1, tots):5.1f}% of samples, {acc[3]} lines")
            phase += 1; acc = [0, 0, 0, 0]; first = k + 1


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
