#!/usr/bin/env python
"""tools/measure_peaks.py [OUT.json] -- the integer roofline denominators of this GPU, measured with csrc/peak.cu
(the same kernels bench.py runs live): ALU pipe alone (LOP3/SHF) and ALU + FMA pipes co-issuing (LOP3 + IMAD)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from subproc_b200 import ops

dev = torch.device("cuda:0")
p = torch.cuda.get_device_properties(dev)
alu = max(ops.int32_peak(dev) for _ in range(3))
dual = max(ops.int32_peak(dev, dual=True) for _ in range(3))
clock_hz = 1.965e9
out = {
    "gpu": p.name, "sms": p.multi_processor_count,
    "alu_pipe_lane_ops_per_s": alu, "alu_plus_fma_lane_ops_per_s": dual,
    "alu_pipe_lanes_per_clk_per_sm": alu / p.multi_processor_count / clock_hz,
    "alu_plus_fma_lanes_per_clk_per_sm": dual / p.multi_processor_count / clock_hz,
    "clock_hz_assumed": clock_hz,
    "how": "othello_int32_peak_kernel / othello_int32_dual_peak_kernel (csrc/peak.cu): 8 blocks/SM x 256 threads x 4096 "
           "rounds of 32 independent lane-ops, best of 5 launches, CUDA events; theoretical ALU pipe = 64 lanes/clk/SM",
}
s = json.dumps(out, indent=1)
print(s)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(s + "\n")
