#!/usr/bin/env python
"""Summarise ncu artefacts into profiles/: a launch list CSV -> per-kernel share table; a --set full .ncu-rep ->
key metrics (duration, DRAM traffic, tensor pipe, issue, occupancy) + the top stall sites.

    python tools/summarize_ncu.py launches <csv> <out.md>
    python tools/summarize_ncu.py full <rep> <out.md>
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# launch list summary of `{path}`\n\nncu `--metrics gpu__time_duration.sum --clock-control none` "
                f"(cold-cache, serialised: compare shares, not absolutes). Total {tot / 1e6:.2f} ms over "
                f"{sum(v[0] for v in agg.values())} launches.\n\n| share | launches | avg us | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write(f"| {v[1] / tot * 100:.1f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k[:110]}` |\n")


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_xu.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## `{name[:100]}` (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")
        srows = list(csv.reader(io.StringIO(src)))
        h = None
        for i, r in enumerate(srows[:5]):
            if "Source" in r and "# Samples" in r:
                h = i
        if h is not None:
            hd = srows[h]
            si, ii = hd.index("Warp Stall Sampling (All Samples)"), hd.index("Instructions Executed")
            seen, data = set(), []
            for r in srows[h + 1:]:
                if len(r) <= si or not r[0].startswith("0x"):
                    continue
                if r[0] in seen:
                    break
                seen.add(r[0])
                data.append(r)
            num = lambda x: int(x) if x.isdigit() else 0  # noqa: E731
            tot = sum(num(r[si]) for r in data) or 1
            f.write(f"## top stall sites (first launch, {tot} samples, {len(data)} SASS instructions)\n\n| samples | executed | SASS |\n|---|---|---|\n")
            for r in sorted(data, key=lambda r: -num(r[si]))[:20]:
                f.write(f"| {num(r[si]) / tot * 100:.1f}% | {num(r[ii])} | `{r[1][:90]}` |\n")
            ops = collections.Counter()
            for r in data:
                op = r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]
                if any(t in op for t in ("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "MUFU")):
                    ops[op.split(".")[0]] += 1
            f.write("\nBlackwell-native SASS present: " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
