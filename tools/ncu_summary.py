"""Turn ncu output into the markdown / json summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r01m_launches.csv  > profiles/r01m_launches_summary.md
    python tools/ncu_summary.py full gpurun_out/r01m_full.ncu-rep profiles/r01m   (writes _ncu_full_summary.md, _traffic.json)

`launches` reads the csv of `ncu --metrics gpu__time_duration.sum --csv --log-file ...`;
`full` reads a `--set full` report through `ncu -i ... --page raw --csv`.
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def _rows(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = _rows(open(path).read())
    head = rows[0]
    ik, iv, iu = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    ig, ib = head.index("Grid Size"), head.index("Block Size")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[head.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else (v * 1000.0 if r[iu] in ("ms", "msecond") else v)
        key = "%s grid=%s block=%s" % (r[ik].split("(")[0], r[ig], r[ib])
        a = agg.setdefault(key, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    print("%d launches, %.1f us in total (cold-cache, serialised: compare SHARES)\n" % (n, tot))
    print("| us/launch | share | launches | kernel |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("| %.1f | %.1f%% | %d | `%s` |" % (a[0] / a[1], 100 * a[0] / tot, a[1], k))


def full(rep, prefix):
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = _rows(text)
    head, units = rows[0], rows[1]
    ik = head.index("Kernel Name")
    seen, out, traffic = set(), [], {}
    stall = [i for i, h in enumerate(head) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[ik]
        if name in seen:
            continue
        seen.add(name)
        out.append("## `%s`\n\n| metric | value | unit |\n|---|---|---|" % name[:100])
        vals = {}
        for m in KEEP:
            if m in head:
                i = head.index(m)
                out.append("| %s | %s | %s |" % (m, r[i], units[i]))
                vals[m] = (r[i], units[i])
        st = sorted(((float(r[i].replace(",", "") or 0), head[i].split("stalled_")[1].split("_per_issue")[0]) for i in stall), reverse=True)[:5]
        out.append("| top stall reasons (warps per issue) | %s | |\n" % ", ".join("%s=%.2f" % (n, v) for v, n in st))

        def to_bytes(m):
            v, u = vals.get(m, ("0", "byte"))
            f = float(v.replace(",", "") or 0)
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        def num(m):
            v, _ = vals.get(m, ("0", ""))
            try:
                return float(v.replace(",", "") or 0)
            except ValueError:
                return 0.0
        dur, du = vals.get("gpu__time_duration.sum", ("0", "us"))
        dur = float(dur.replace(",", "") or 0) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(du, 1.0)
        traffic[name.split("(")[0].replace("void ", "").strip()] = {
            "kernel": name[:80], "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
            "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
            "warp_instructions": num("smsp__inst_executed.sum"),
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "registers": num("launch__registers_per_thread"), "ncu_duration_us": dur}
    open(prefix + "_ncu_full_summary.md", "w").write(
        "# `ncu --set full --clock-control none --import-source on`, first captured launch of each kernel\n\nSource: %s (not committed)\n\n" % rep + "\n".join(out))
    json.dump({"source": prefix + "_ncu_full_summary.md", "kernels": traffic}, open(prefix + "_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3])
