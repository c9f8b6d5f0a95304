#!/usr/bin/env python
"""Condense one `ncu --set full` report into the numbers DESIGN.md / bench.py quote.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rXX_ncu_<kernel>.json [launch index in the report]
(runs `ncu -i ... --page raw --csv` and `--page source --csv --print-source sass`)"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "sm__cycles_elapsed.avg": "cycles_elapsed",
    "sm__cycles_active.avg": "cycles_active",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_pct_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_pct_active",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_kb",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "lts__t_sectors.sum": "l2_sectors", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum": "ldgsts_smem_wavefronts",
    "smsp__inst_executed_op_ldgsts.sum": "ldgsts_instructions",
}
STALLS = "smsp__average_warps_issue_stalled_"


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return float(v) * mul if mul else float(v)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u, v = r[0], r[1], r[2 + which]
    res = {"report": rep, "kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else None, "stalls_per_issue": {}}
    for n, un, val in zip(h, u, v):
        if n in KEYS:
            x = to_bytes(val, un) if n.startswith("dram__bytes") else float(val.replace(",", ""))
            if n == "gpu__time_duration.sum":
                x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(un, 1.0)
            res[KEYS[n]] = x
        elif n.startswith(STALLS) and n.endswith("_per_issue_active.ratio"):
            x = float(val)
            if x >= 0.05:
                res["stalls_per_issue"][n[len(STALLS):-len("_per_issue_active.ratio")]] = round(x, 3)
    if "dram_read" in res:
        res["dram_bytes"] = res["dram_read"] + res.get("dram_write", 0.0)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    s = list(csv.reader(io.StringIO(src)))
    # one section per launch (a report with several launches repeats "Kernel Name" / header rows; ncu prints each twice)
    heads = [i for i, x in enumerate(s) if x and x[0] == "Kernel Name"]
    if len(heads) > 1:
        per = 2 if len(heads) >= 2 * (len(r) - 2) else 1
        a = heads[min(per * which, len(heads) - 1)]
        nxt = [i for i in heads if i > a]
        s = s[a:(nxt[0] if nxt else len(s))]
    if len(s) > 2:
        hh = s[1]
        c, ie = hh.index("Warp Stall Sampling (All Samples)"), hh.index("Instructions Executed")
        rows = [(int(x[c] or 0), int(x[ie] or 0), x[1].strip()) for x in s[2:] if len(x) > ie]
        tot = max(1, sum(a for a, _, _ in rows))
        res["sass_instructions"] = len(rows)
        mix = {}
        for _, b, t in rows:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            op = op.split(".")[0]
            mix[op] = mix.get(op, 0) + b
        res["executed_mix_top"] = dict(sorted(mix.items(), key=lambda kv: -kv[1])[:12])
        res["top_stall_sites"] = [{"pct": round(100.0 * a / tot, 1), "executed": b, "sass": t[:80]}
                                  for a, b, t in sorted(rows, key=lambda x: -x[0])[:8]]
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in res if k not in ("top_stall_sites", "executed_mix_top")}, indent=1))


if __name__ == "__main__":
    main()
