#!/usr/bin/env python
"""Collect the bench lines of the multi-GPU / other-config sessions under gpurun_out/ into profiles/r2_range_split.{json,md}
and profiles/r2_other_configs.{json,md}.  Pure bookkeeping: no number is computed here except ratios of reported times."""
import json
import os
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def line(name):
    try:
        return json.loads(open(os.path.join(G, name)).read().strip().splitlines()[-1])
    except Exception:
        return None


def brief(d):
    k = d["config"].get("key") or {}
    return {"n_gpus": d["n_gpus"], "proofs_per_s": round(d["value"], 3), "ms_per_proof": round(d["ms_per_step"] / d["config"].get("proofs_per_step_per_gpu", 1) if d["scaling"] == "weak" and "proofs_per_step_per_gpu" in d["config"] else d["ms_per_step"], 2),
            "e2e_proofs_per_s": round(d["e2e"]["value"], 3), "scaling": d["scaling"], "steps": d["steps"], "warmup": d["warmup"],
            "wire_window_bits": k.get("wire_window_bits"), "z_window_bits": k.get("z_window_bits"),
            "table_gb_per_gpu": round(k.get("table_gb_per_gpu", k.get("table_gb", 0)), 2), "register_s": round(k.get("register_s", 0), 2),
            "sm_mhz": (d.get("clocks") or {}).get("sm_mhz"), "reasons": (d.get("clocks") or {}).get("reasons")}


def timeline(name):
    try:
        d = json.load(open(os.path.join(G, name)))
    except Exception:
        return None
    agg = defaultdict(lambda: [0.0, 1e9, 0.0, 0])
    for r in d["records"]:
        a = agg[r["phase"]]
        a[0] += r["end_ms"] - r["start_ms"]
        a[1] = min(a[1], r["start_ms"])
        a[2] = max(a[2], r["end_ms"])
        a[3] += 1
    return {"wall_ms": round(d["wall_ms"], 2),
            "phases": {k: {"launch_groups": a[3], "busy_ms": round(a[0], 2), "first_start_ms": round(a[1], 2), "last_end_ms": round(a[2], 2)}
                       for k, a in sorted(agg.items(), key=lambda kv: kv[1][1])}}


def main():
    sessions = [
        ("r2b", "4-GPU box; model window on every slice, equal slices", {"agg": [("r2b_agg_n1.json", "r2b_agg_tl_n1.json"), ("r2b_agg_n2.json", "r2b_agg_tl_n2.json"), ("r2b_agg_n4.json", "r2b_agg_tl_n4.json")], "agg_unsharded_quotient": [("r2b_agg_n4_unsharded.json", None)], "st": [("r2b_st_n4.json", "r2b_st_tl_n4.json")]}),
        ("r2e", "8-GPU box; same build as r2b", {"agg": [("r2e_agg_n8.json", "r2e_agg_tl_n8.json")], "st": [("r2e_st_n8.json", "r2e_st_tl_n8.json")]}),
        ("r2g", "4- and 8-GPU boxes; wire window two bits below the model on slices too (rejected), lighter rank 0", {"agg": [("r2g_agg_n2.json", None), ("r2g_agg_n4.json", "r2g_agg_tl_n4.json"), ("r2g_agg_n8.json", "r2g_agg_tl_n8.json")], "st": [("r2g_st_n4.json", None), ("r2g_st_n8.json", None)]}),
        ("r2k", "2-GPU box; FINAL build (model window on slices, lighter rank 0)", {"agg": [("r2k_agg_n2.json", None)]}),
    ]
    out = {"note": "range-split = ONE proof split by point range over N GPUs (bench.py --config aggregator|statetransition --gpus N); "
                   "times are device-timed, max over ranks; N = 1 is the same code path on one GPU", "sessions": []}
    md = ["# Round 2 - one proof split over N GPUs (range split)", "",
          "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --config aggregator|statetransition --steps 4 --warmup 2`",
          "(aggregator = BW6-761, n = m = 2^22; statetransition = BN254, n = m = 2^24, incl. the blob KZG flow).  Source lines: `gpurun_out/r2[begk]_*.json`.", ""]
    for tag, what, groups in sessions:
        s = {"session": tag, "what": what, "runs": {}}
        md += ["## %s - %s" % (tag, what), "", "| config | N | ms / proof | proofs/s | e2e proofs/s | wire c | Z c | tables GB/GPU | key setup s |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
        for cfg, files in groups.items():
            rows = []
            for f, tl in files:
                d = line(f)
                if not d:
                    continue
                b = brief(d)
                b["source"] = "gpurun_out/" + f
                if tl:
                    t = timeline(tl)
                    if t:
                        b["rank0_timeline"] = t
                rows.append(b)
                md.append("| %s | %d | %.1f | %.2f | %.2f | %s | %s | %.1f | %.1f |" % (cfg, b["n_gpus"], d["ms_per_step"], b["proofs_per_s"], b["e2e_proofs_per_s"], b["wire_window_bits"], b["z_window_bits"], b["table_gb_per_gpu"], b["register_s"]))
            s["runs"][cfg] = rows
        md.append("")
        out["sessions"].append(s)
    # speed-ups of the accepted policy
    n1 = line("r2b_agg_n1.json")["ms_per_step"]
    acc = {"N=2 (r2k, final build)": line("r2k_agg_n2.json")["ms_per_step"], "N=2 (r2b)": line("r2b_agg_n2.json")["ms_per_step"], "N=4 (r2b)": line("r2b_agg_n4.json")["ms_per_step"],
           "N=8 (r2e)": line("r2e_agg_n8.json")["ms_per_step"], "N=8 (r2g)": line("r2g_agg_n8.json")["ms_per_step"]}
    out["aggregator_speedup_vs_n1"] = {k: round(n1 / v, 2) for k, v in acc.items()}
    md += ["## Aggregator proof latency against N = 1 (%.1f ms)" % n1, ""] + ["* %s: %.1f ms -> %.2fx" % (k, v, n1 / v) for k, v in acc.items()] + [""]
    md += ["## Rank 0 timelines (busy ms per phase; phases overlap on different streams)", ""]
    for f in ["r2b_agg_tl_n1.json", "r2b_agg_tl_n2.json", "r2b_agg_tl_n4.json", "r2g_agg_tl_n8.json", "r2b_st_tl_n4.json", "r2e_st_tl_n8.json"]:
        t = timeline(f)
        if not t:
            continue
        md += ["### %s (wall %.1f ms)" % (f, t["wall_ms"]), "", "| phase | groups | busy ms | first start | last end |", "|---|---:|---:|---:|---:|"]
        md += ["| %s | %d | %.2f | %.2f | %.2f |" % (k, v["launch_groups"], v["busy_ms"], v["first_start_ms"], v["last_end_ms"]) for k, v in t["phases"].items()] + [""]
    json.dump(out, open(os.path.join(P, "r2_range_split.json"), "w"), indent=1)
    open(os.path.join(P, "r2_range_split.md"), "w").write("\n".join(md) + "\n")

    # ---- other single-GPU configurations
    oc = {}
    md = ["# Round 2 - other configurations on one B200", "", "| run | metric | value | unit | ms / step | e2e | notes |", "|---|---|---:|---|---:|---:|---|"]
    for tag, f, note in [("voteverifier (headline, final build)", "r2l_bench_n1.json", "4 proofs per step, 4 in flight"),
                         ("voteverifier, own G2 window (B200_G2_WINDOW=model)", "r2i_bench_g2own.json", "A/B of the G2 window policy; single-proof latency 58.4 ms vs 67 ms"),
                         ("voteverifier, shared G2 window", "r2i_bench_g2w18.json", "same box, back to back with the line above"),
                         ("statetransition BN254 2^24 (+ blob flow)", "r2f_st_n1.json", "2 slots"),
                         ("aggregator BW6-761 2^22", "r2f_agg_n1.json", ""),
                         ("blob KZG (commitment + opening)", "r2f_blob.json", ""),
                         ("CPU reference arm (voteverifier, 16 threads)", "r2f_reference.json", "bench.py --impl reference")]:
        d = line(f)
        if not d:
            continue
        k = d["config"].get("key") or {}
        oc[tag] = {"source": "gpurun_out/" + f, "metric": d["metric"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"], "key": k,
                   "clocks": d.get("clocks"), "cpu_baseline": d.get("cpu_baseline"), "note": note}
        extra = note
        if k:
            extra += " tables %.1f GB, wire c=%s, Z c=%s" % (k.get("table_gb", k.get("table_gb_per_gpu", 0)), k.get("wire_window_bits"), k.get("z_window_bits"))
        md.append("| %s | %s | %.4g | %s | %.1f | %.4g | %s |" % (tag, d["metric"], d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], extra.strip()))
    json.dump(oc, open(os.path.join(P, "r2_other_configs.json"), "w"), indent=1)
    open(os.path.join(P, "r2_other_configs.md"), "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
