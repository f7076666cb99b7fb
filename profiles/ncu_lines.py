"""Per-SOURCE-line executed instructions and stall samples of one kernel, straight from an ncu report taken with
--import-source on (the CUDA <-> SASS correlation of the report itself, so the library need not be the same build).
usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import collections
import csv
import subprocess
import sys


def main(rep, kern, top=45):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    fname, h = None, None
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            h = r
            il, isrc, isamp, iinst = 0, 1, h.index("# Samples"), h.index("Instructions Executed")
            continue
        if h is None or len(r) <= iinst:
            continue
        if r[il].isdigit():          # a CUDA source line row; following rows with empty line no are its SASS
            cur = (fname, int(r[il]))
            text[cur] = r[isrc].strip()
            continue
        if r[il] == "" and r[isamp].isdigit() and r[iinst].isdigit():
            inst[cur] += int(r[iinst])
            samp[cur] += int(r[isamp])
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"# {kern}: {ti} warp instructions, {ts} stall samples; share of instructions / of samples per source line")
    print("file,line,inst_share,sample_share,source")
    for k, c in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:top]:
        print(f"{k[0]},{k[1]},{c / ti:.4f},{samp[k] / max(ts, 1):.4f},\"{text.get(k, '')[:110]}\"")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)
