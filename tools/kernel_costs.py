#!/usr/bin/env python
"""tools/kernel_costs.py NCU_SUMMARY.txt POSITIONS OUT.json -- turn the ncu summary of one playout launch
(tools/ncu_summary.py) into the tracked figures bench.py quotes next to its roofline: executed thread
instructions per position and the measured pipe utilisations."""
import json
import re
import sys

text = open(sys.argv[1]).read().split("\n== ")[0]
positions = float(sys.argv[2])


def val(name):
    m = re.search(r"^\s+%s\s+([0-9.]+)" % re.escape(name), text, flags=re.M)
    return float(m.group(1)) if m else None


warp_inst = val("smsp__inst_executed.sum")
lanes = val("smsp__thread_inst_executed_per_inst_executed.ratio")
dur = val("gpu__time_duration.sum")
unit = re.search(r"gpu__time_duration.sum\s+[0-9.]+\s+(\w+)", text).group(1)
dur_ms = dur * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[unit]
out = {
    "kernel": text.splitlines()[0].replace("== ", "")[:80],
    "source": sys.argv[1],
    "positions_per_launch": positions,
    "duration_ms_under_ncu": dur_ms,
    "positions_per_ms_under_ncu": positions / dur_ms,
    "warp_inst_per_launch": warp_inst,
    "active_lanes_per_inst": lanes,
    "thread_inst_per_position": warp_inst * lanes / positions,
    "warp_inst_per_position_warp": warp_inst * 32 / positions,
    "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fmaheavy_pipe_pct": val("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "xu_pipe_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "issue_slot_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_bytes_per_launch": (val("dram__bytes_read.sum") or 0) * 1e6 + (val("dram__bytes_write.sum") or 0) * 1e9,
}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
