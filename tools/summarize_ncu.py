#!/usr/bin/env python
"""Turn the ncu CSVs of tools/gpu_round.sh into the markdown summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01b_launches.csv [--last-step N] > profiles/..._launches.md
    python tools/summarize_ncu.py full gpurun_out/r01b_full_raw.csv > profiles/..._ncu_full.md
    python tools/summarize_ncu.py traffic gpurun_out/r02s_full_raw.csv <samples of that launch> <source label>
        -> updates profiles/traffic.json (DRAM bytes per audio sample of each captured kernel; bench.py's roofline.traffic)
"""
import collections
import csv
import re
import sys


def read_csv(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("m3::", "").replace("(anonymous namespace)::", "")


def launches(path, last_n=None):
    rows = read_csv(path)
    if last_n:
        rows = rows[-last_n:]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e6
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {total:.3f} | |")


FULL_COLS = [
    ("gpu__time_duration.sum", "time ms", 1e-6, "ns"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe %", 1, None),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma active %", 1, None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM busy %", 1, None),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1, None),
    ("launch__registers_per_thread", "regs", 1, None),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB", 1e-3, "byte"),
    ("dram__bytes_read.sum", "DRAM rd MB", None, None),
    ("dram__bytes_write.sum", "DRAM wr MB", None, None),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1, None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1, None),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem %", 1, None),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %", 1, None),
]


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    u = (unit or "").lower().split("/")[0]
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return x * mult


def full(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    print("| kernel | grid | " + " | ".join(c[1] for c in FULL_COLS if c[0] in idx) + " |")
    print("|---|---|" + "---|" * sum(1 for c in FULL_COLS if c[0] in idx))
    for row in rd:
        if not row:
            continue
        cells = []
        for name, _, scale, _ in FULL_COLS:
            if name not in idx:
                continue
            v, u = row[idx[name]], units[idx[name]]
            try:
                if "bytes" in name and scale is None:
                    cells.append(f"{to_bytes(v, u) / 1e6:.1f}")
                elif name.startswith("launch__shared"):
                    cells.append(f"{to_bytes(v, u) / 1e3:.1f}")
                elif name == "gpu__time_duration.sum":
                    x = float(v.replace(",", ""))
                    ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6) * x
                    cells.append(f"{ms:.3f}")
                else:
                    cells.append(f"{float(v.replace(',', '')):.1f}")
            except ValueError:
                cells.append(v)
        print(f"| {short(row[idx['Kernel Name']])} | {row[idx['Grid Size']]} | " + " | ".join(cells) + " |")


def traffic(path, samples, source):
    import json
    import pathlib
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    out_path = pathlib.Path(__file__).resolve().parent.parent / "profiles" / "traffic.json"
    table = json.loads(out_path.read_text()) if out_path.exists() else {}
    for row in rd:
        if not row:
            continue
        name = re.sub(r"<.*$", "", short(row[idx["Kernel Name"]]))
        rdb = to_bytes(row[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wrb = to_bytes(row[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        table[name] = {"dram_bytes_per_sample": (rdb + wrb) / samples, "dram_read_bytes": rdb, "dram_write_bytes": wrb,
                       "samples_in_launch": samples, "source": source}
        print(name, table[name])
    out_path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "traffic":
        traffic(path, float(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    if mode == "launches":
        n = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] == "--last-step" else None
        launches(path, n)
    else:
        full(path)
