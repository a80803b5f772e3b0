"""Join an ncu source page (SASS view, csv) with nvdisasm -g line info: executed warp-instructions per source line.
    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libdyros_b200.so; nvdisasm -g -c physics_kernels.sm_100a.cubin > dis.txt
    python tools/sass_lines.py src.csv dis.txt k_step_physics [top]
"""
import collections
import csv
import re
import sys


def main():
    src, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    rows = list(csv.reader(open(src)))
    H, D = rows[1], rows[2:]
    ia, isrc = H.index("Instructions Executed"), H.index("Source")
    counts = [int(r[ia]) for r in D]
    # nvdisasm: find the function section, collect (file, line) per instruction in order
    lines = open(dis).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.") and kern in l)
    cur, per = ("?", 0), []
    for l in lines[start + 1:]:
        if l.startswith("//--------------------- "):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            per.append(cur)
    print(f"# {len(per)} instructions in nvdisasm, {len(counts)} in the ncu page", file=sys.stderr)
    n = min(len(per), len(counts))
    agg = collections.Counter()
    for i in range(n):
        agg[per[i]] += counts[i]
    tot = sum(counts)
    byfile = collections.Counter()
    for (f, ln), v in agg.items():
        byfile[f] += v
    print("total", tot, dict(byfile))
    for (f, ln), v in agg.most_common(top):
        print(f"{v:10d} {v / tot * 100:5.1f}%  {f}:{ln}")


if __name__ == "__main__":
    main()
