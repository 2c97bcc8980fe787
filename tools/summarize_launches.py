"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/one_step.py`)
into the per-kernel summary committed under profiles/: the LAST n launches are the measured step (n = the `launches_in_measured_step`
line tools/one_step.py prints).  usage: python tools/summarize_launches.py launches.csv n > summary.csv"""
import csv
import re
import sys


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    for r in rd:
        if r[im] == "gpu__time_duration.sum":
            rows.append((r[ik], float(r[iv].replace(",", ""))))
    step = rows[-n:]
    total = sum(t for _, t in step)
    agg = {}
    for name, t in step:
        short = re.sub(r"\(.*$", "", name).strip()
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1; a[1] += t
    print(f"# one C2 bench step ({n} launches, tools/one_step.py under ncu --metrics gpu__time_duration.sum --clock-control none, TNQS_NO_GRAPH=1); total {total / 1e6:.1f} ms")
    print("kernel,launches,total_us,share")
    for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{cnt},{t / 1e3:.1f},{t / total:.4f}")


if __name__ == "__main__":
    main()
