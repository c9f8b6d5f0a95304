"""Sum an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel over the LAST force step in it.

A step starts at a `tree_keys_kernel` launch; the rows from the last one to the end of the file are one step.
Times are serialised and cold-cache (ncu), so it is each kernel's SHARE of the step that is compared with the
bench line's phases, not the absolute numbers.  usage: python tools/launch_summary.py launches.csv [out.json]"""
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = re.sub(r"^cb200::", "", name)
    return name


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        rows.append((short(r["Kernel Name"]), ns))
    starts = [i for i, (k, _) in enumerate(rows) if k.startswith("tree_keys_kernel")]
    step = rows[starts[-1]:] if starts else rows
    per = {}
    for k, ns in step:
        e = per.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += ns
    total = sum(ns for _, ns in step)
    out = {"source": path, "launches_in_file": len(rows), "steps_in_file": len(starts), "launches_in_step": len(step),
           "step_sum_ms": total * 1e-6,
           "kernels": [{"kernel": k, "launches": c, "ms": ns * 1e-6, "share": ns / total}
                       for k, (c, ns) in sorted(per.items(), key=lambda kv: -kv[1][1])]}
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    for e in out["kernels"][:14]:
        print(f"{e['ms']:9.3f} ms {100 * e['share']:5.1f}%  x{e['launches']:<4d} {e['kernel'][:90]}")
    print(f"{out['step_sum_ms']:9.3f} ms  sum of {len(step)} launches")


if __name__ == "__main__":
    main()
