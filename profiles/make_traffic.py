#!/usr/bin/env python
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of ONE step, from a summary written
by profiles/summarize_ncu.py: the `roofline.traffic` source of bench.py.
Usage: python profiles/make_traffic.py profiles/X_full_ncu_summary.txt > profiles/X_traffic.json
The capture holds two or more steps (warm-up + timed ...); the last complete one is kept."""
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path):
    entries, cur = [], None
    for line in open(path):
        m = re.match(r"\[(\d+)\] (.*)", line)
        if m:
            name = re.sub(r"^void ", "", m.group(2)).split("(")[0]
            cur = {"name": name}
            entries.append(cur)
            continue
        m = re.match(r"\s+(dram read|dram write|duration)\s+([\d.]+)\s*(\S*)", line)
        if m and cur is not None:
            key, val, unit = m.group(1), float(m.group(2)), m.group(3)
            cur[key] = val * UNIT.get(unit, 1.0) if key != "duration" else f"{val} {unit}"
    firsts = [i for i, e in enumerate(entries) if e["name"].startswith(("cap_analysis", "udgrade", "ring_analysis"))]
    steps = [entries[a:b] for a, b in zip(firsts, firsts[1:] + [len(entries)])] or [entries]
    full = max(len(x) for x in steps)
    step = [x for x in steps if len(x) == full][-1]          # the last COMPLETE step of the capture
    out, seen = {}, {}
    for e in step:
        k = seen.get(e["name"], 0) + 1
        seen[e["name"]] = k
        key = e["name"] if k == 1 else f"{e['name']}#{k}"
        out[key] = {"traffic_bytes": e.get("dram read", 0.0) + e.get("dram write", 0.0), "dram_read": e.get("dram read", 0.0),
                    "dram_write": e.get("dram write", 0.0), "duration": e.get("duration")}
    json.dump({"source": f"ncu --set full --clock-control none, bench.py --steps 1 --warmup 1 at cfg4 ({path}); per launch",
               "kernels": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
