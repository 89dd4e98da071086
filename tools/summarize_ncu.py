#!/usr/bin/env python
"""Turn ncu CSV exports (launch list + `--page raw` of a `--set full` capture) into the markdown
summaries kept under profiles/.   usage: summarize_ncu.py [--traffic-json OUT.json] TITLE launches.csv raw.csv [raw2.csv ...]

--traffic-json also writes {short kernel name: {"bytes_per_launch": dram read + write of the first captured
launch, "duration_ms", "source": raw csv + kernel}} -- the file bench.py's roofline `traffic` keys cite."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def launch_table(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("second", "s").replace("usecond", "us"), None)
        if scale is None:
            scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[ui], 1e-6)
        n, t = agg.get(r[ki], (0, 0.0))
        agg[r[ki]] = (n + 1, t + v * scale)
    tot = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k[:70], n, t, 100 * t / tot))
    return "\n".join(out)


def raw_tables(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out, seen = [], set()
    for r in rows[2:]:
        name = r[ki]
        if name in seen:
            continue
        seen.add(name)
        out.append("### `%s`\n" % name[:90])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append("- %s = %s %s" % (m, r[i], units[i]))
        out.append("")
    return "\n".join(out)


def _to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def _to_ms(v, unit):
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
                                         "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[unit]


def traffic_json(raws):
    out = {}
    for path in raws:
        rows = list(csv.reader(open(path, errors="replace")))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        ir, iw, it = (hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        for r in rows[2:]:
            short = re.sub(r"^void\s+", "", r[ki])
            short = re.split(r"[<(]", short)[0].split("::")[-1]
            if short in out and out[short]["_file"] == path:
                continue
            if short in out:                    # same kernel captured by another command: key it by file
                short = "%s@%s" % (short, os.path.basename(path))
                if short in out:
                    continue
            out[short] = {"bytes_per_launch": _to_bytes(r[ir], units[ir]) + _to_bytes(r[iw], units[iw]),
                          "dram_read_bytes": _to_bytes(r[ir], units[ir]), "dram_write_bytes": _to_bytes(r[iw], units[iw]),
                          "duration_ms_under_ncu": _to_ms(r[it], units[it]), "_file": path,
                          "source": "ncu --set full --clock-control none, first captured launch of `%s` (%s)" % (
                              r[ki][:80], os.path.basename(path))}
    for v in out.values():
        del v["_file"]
    return out


if __name__ == "__main__":
    if sys.argv[1] == "--traffic-json":
        tj = sys.argv[2]
        del sys.argv[1:3]
        json.dump(traffic_json(sys.argv[3:]), open(tj, "w"), indent=1)
    title, launches, raws = sys.argv[1], sys.argv[2], sys.argv[3:]
    print("# %s\n" % title)
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), share of device time\n")
    print(launch_table(launches))
    print("\n## `ncu --set full --clock-control none`, first captured launch of each kernel\n")
    for r in raws:
        print(raw_tables(r))
