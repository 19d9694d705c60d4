"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import csv, collections, re, sys

def main(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mv, mn, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]); name = re.sub(r"^void ", "", name)
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k[:70], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    txt = "\n".join(lines) + "\n\ntotal %.1f us over %d launches (ncu serialises launches and runs them cold-cache: compare shares, not absolutes)\n" % (tot, sum(v[0] for v in agg.values()))
    open(out, "w").write(txt) if out else None
    print(txt)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
