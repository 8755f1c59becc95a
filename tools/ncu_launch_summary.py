"""Per-kernel totals / shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`).

    python tools/ncu_launch_summary.py gpurun_out/launches.csv profiles/rNN_ncu_launch_summary.csv "comment" ...
"""
import collections
import csv
import re
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    comments = sys.argv[3:]
    tot = collections.defaultdict(float)
    cnt = collections.defaultdict(int)
    n = 0
    with open(src, newline="") as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    head = rows[0]
    ik, iv, im = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("b2v::", "")
        tot[name] += float(r[iv].replace(",", "")) / 1e3
        cnt[name] += 1
        n += 1
    total = sum(tot.values())
    with open(out, "w", newline="") as f:
        for c in comments:
            f.write("# " + c + "\n")
        f.write(f"# total {total / 1e3:.1f} ms over {n} launches\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k in sorted(tot, key=lambda k: -tot[k]):
            w.writerow([k, cnt[k], f"{tot[k]:.1f}", f"{tot[k] / total:.4f}"])
    print(open(out).read())


if __name__ == "__main__":
    main()
