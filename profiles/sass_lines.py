"""Executed warp instructions per SOURCE line of one kernel: joins `nvdisasm -g` (line info of the cubin
inside the built library) with the per-SASS-instruction counts of an ncu report (same build!).
usage: python profiles/sass_lines.py <report.ncu-rep> <kernel regex> <mangled-name substring> [top N]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flamegpu2-prisoners-dilemma-abm_b200")


def sass_with_lines(mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(PKG, "libpdabm.so")], cwd=tmp, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(text) if l.startswith(".text.") and mangled in l and l.rstrip().endswith(":")][0]
    cur, seq = None, []
    for l in text[start + 1:]:
        if l.startswith("//---") and ".text." in l:
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append(cur)
    return seq


def main(rep, kern, mangled, top=40):
    seq = sass_with_lines(mangled)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[1]
    ci = h.index("Instructions Executed")
    counts = [int(r[ci]) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
    if len(counts) == 2 * len(seq):  # ncu lists the function twice
        counts = counts[:len(seq)]
    assert len(counts) == len(seq), f"ncu has {len(counts)} SASS lines, the cubin {len(seq)}: different builds?"
    tot = sum(counts)
    by = collections.Counter()
    for k, c in zip(seq, counts):
        by[k or ("?", 0)] += c
    src = {}
    print("share_pct,file,line,source")
    for (f, ln), c in by.most_common(top):
        if f not in src:
            path = os.path.join(PKG, "csrc", f)
            src[f] = open(path).read().split("\n") if os.path.exists(path) else []
        txt = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
        print(f'{100 * c / tot:.2f},{f},{ln},"{txt[:100].replace(chr(34), chr(39))}"')


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
