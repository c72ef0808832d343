#!/usr/bin/env python3
"""Turns ncu outputs (gpurun_out/) into the committed summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md> [title]
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md> [title]
  python tools/summarize_ncu.py facts <report.ncu-rep> <out.json> <key>=<points> [...]
      per-launch facts bench.py cites (profiles/r2_ncu_facts.json): for every <key> (e.g. k_msm_accumulate_g1) the
      first kernel whose name matches the key's pattern gives dram bytes, fmaheavy %, duration; <points> = the
      number of base points that launch processed
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("b200::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write("# %s\n\nSource: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are "
                 "cold-cache and serialised - compare SHARES, not absolutes).\n\n" % (title, path))
        fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:110], n, t, 100 * t / tot))
        fh.write("\nTotal kernel time: %.3f ms over %d launches.\n" % (tot, sum(a[0] for a in agg.values())))


def _raw_rows(path):
    """rows of `ncu --page raw --csv`: from a report, or from a CSV already exported on the GPU box (reports with
    --import-source exceed what gpurun copies back)"""
    if path.endswith(".csv"):
        return list(csv.reader(open(path)))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(raw.splitlines()))


def full(path, out, title):
    rows = _raw_rows(path)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as fh:
        fh.write("# %s\n\nSource: `%s` (ncu --set full --clock-control none).\n\n" % (title, path))
        for r in rows[2:]:
            fh.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[idx["Kernel Name"]].split("(")[0][:120])
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "n/a"):
                    fh.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            fh.write("\n")


FACT_PATTERNS = {"k_msm_accumulate_g1": ("k_msm_accumulate<", "FpT<"), "k_msm_accumulate_g2": ("k_msm_accumulate<", "Fp2T<"),
                 "k_ntt_pass": ("k_ntt_pass<", ""), "k_ntt_fused": ("k_ntt_fused", "")}


def facts(path, out, specs):
    import json
    import os
    rows = _raw_rows(path)
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    res = json.load(open(out)) if os.path.exists(out) else {}

    def num(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except (KeyError, ValueError):
            return None

    units = rows[1]

    def to_bytes(r, k):
        v = num(r, k)
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[idx[k]], 1)

    for spec in specs:
        key, pts = spec.split("=")
        a, b = FACT_PATTERNS[key]
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            if a in name and b in name.split("(")[0]:
                rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
                res[key] = {"points": int(pts), "dram_bytes": (rd or 0) + (wr or 0), "dram_read": rd, "dram_write": wr,
                            "fmaheavy_pct": num(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                            "registers": num(r, "launch__registers_per_thread"), "source": os.path.basename(path),
                            "kernel": name.split("(")[0][:100]}
                break
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if sys.argv[1] == "facts":
        facts(sys.argv[2], sys.argv[3], sys.argv[4:])
        sys.exit(0)
    mode, path, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else path
    (launches if mode == "launches" else full)(path, out, title)
