"""Summarise ncu output brought back from the GPU box into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg")


def short(name):
    name = re.sub(r"vmb::<unnamed>::|<unnamed>::|void ", "", name)
    return re.sub(r"\((const|float|long|int|void|vmb|CUtensorMap|HeadDev|unsigned|__nv).*", "", name)[:90]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms of device time (ncu: cold cache, serialised — compare shares)")
    print(f"{'share':>7} {'n':>5} {'avg us':>10}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[0]:5d} {v[1] / v[0]:10.1f}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: one row per captured launch (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print(f"\n## {short(r[cols['Kernel Name']])}  grid {r[cols['Grid Size']]} block {r[cols['Block Size']]}")
        for m in FULL_METRICS:
            if m in cols:
                print(f"  {m:70s} {r[cols[m]]:>18s} {units[cols[m]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
