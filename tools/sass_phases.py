"""Executed warp-instructions of a kernel by OUTERMOST source line of a given file (inlining resolved), from an ncu
source page (SASS view, csv) joined with `nvdisasm -gi -c` of the same cubin; optionally summed over line ranges.
    python tools/sass_phases.py src.csv dis_gi.txt k_step_physics physics_lanes.cuh [name:lo-hi ...]
"""
import collections
import csv
import re
import sys


def main():
    src, dis, kern, fname = sys.argv[1:5]
    ranges = []
    for a in sys.argv[5:]:
        n, r = a.split(":")
        lo, hi = r.split("-")
        ranges.append((n, int(lo), int(hi)))
    rows = list(csv.reader(open(src)))
    H, D = rows[1], rows[2:]
    ia = H.index("Instructions Executed")
    counts = [int(r[ia]) for r in D]
    lines = open(dis).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l)
    frames, per, fresh = [], [], True  # an annotation block replaces the frames; instructions without one inherit them
    for l in lines[start + 1:]:
        if l.startswith("//--------------------- "):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                frames, fresh = [], False
            frames.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            fresh = True
            mine = [f for f in frames if f[0] == fname]
            per_other = frames[-1] if frames else ("?", 0)
            per.append(mine[-1][1] if mine else ("other", per_other[0], per_other[1]))
    assert len(per) == len(counts), (len(per), len(counts))
    tot = sum(counts)
    agg = collections.Counter()
    for p, c in zip(per, counts):
        agg[p] += c
    if ranges:
        out = collections.Counter()
        for p, c in agg.items():
            if isinstance(p, tuple):
                out[f"({p[1]})"] += c
                continue
            for n, lo, hi in ranges:
                if lo <= p <= hi:
                    out[n] += c
                    break
            else:
                out["unassigned"] += c
        for n, c in out.most_common():
            print(f"{c:11d} {c / tot * 100:5.1f}%  {n}")
    else:
        for p, c in agg.most_common(60):
            print(f"{c:11d} {c / tot * 100:5.1f}%  {p}")


if __name__ == "__main__":
    main()
