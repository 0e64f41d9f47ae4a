"""Per-opcode and per-instruction warp-stall samples of `ncu --set full --import-source on` captures.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv ; python profiles/r02/stall_table.py src.csv [...] -> markdown"""
import csv
import sys
from collections import defaultdict


def tables(path):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    seen = set()
    for k, a in enumerate(starts):
        name = rows[a][1].split("(nw::")[0].replace("void nw::", "")
        if name in seen:
            continue
        seen.add(name)
        b = starts[k + 1] if k + 1 < len(starts) else len(rows)
        h = rows[a + 1]
        seg = [r for r in rows[a + 2:b] if len(r) > 10]
        iS, iI = h.index("# Samples"), h.index("Instructions Executed")
        stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[iS] or 0) for r in seg) or 1
        print("## `%s`  (%d stall samples, %s)\n" % (name, tot, path.split("/")[-1]))
        agg = defaultdict(int)
        for r in seg:
            for c in stall:
                agg[c] += int(r[h.index(c)] or 0)
        print("stall reasons: " + ", ".join("%s %.0f %%" % (c[6:], 100.0 * v / tot) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]) + "\n")
        byop = defaultdict(lambda: [0, 0])
        for r in seg:
            t = r[1].split()
            op = t[0] if not t[0].startswith("@") else t[1]
            byop[op][0] += int(r[iS] or 0)
            byop[op][1] += int(r[iI] or 0)
        print("| opcode | samples | share | warp instructions |\n|---|---|---|---|")
        for op, (s, i) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:10]:
            print("| `%s` | %d | %.1f %% | %d |" % (op, s, 100.0 * s / tot, i))
        print("\n| samples | executed | instruction | top stall reasons |\n|---|---|---|---|")
        for r in sorted(seg, key=lambda r: -int(r[iS] or 0))[:10]:
            st = sorted(((c[6:], int(r[h.index(c)] or 0)) for c in stall), key=lambda kv: -kv[1])[:2]
            print("| %s | %s | `%s` | %s |" % (r[iS], r[iI], " ".join(r[1].split())[:70], ", ".join("%s %d" % kv for kv in st if kv[1])))
        print()


if __name__ == "__main__":
    print("# r02 - where the warps wait: stall samples per opcode / instruction (final build, `profiles/r02/measure_final.sh` captures)\n")
    for p in sys.argv[1:]:
        tables(p)
