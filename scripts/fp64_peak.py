#!/usr/bin/env python
"""FP64 peak of the box: cuBLAS DGEMM 8192^3 through torch.matmul (plumbing only), burst = best of 10,
sustained = back-to-back launches for >= 4 s; SM clock sampled alongside.  Prints one JSON line.
BASELINE.md section 2 asks for this number next to every FP64 / DMMA fraction."""
import json
import subprocess
import sys
import time

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    flop = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(1, int(secs / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / reps
    print(json.dumps({"what": f"cuBLAS DGEMM {n}^3 via torch.matmul(float64)", "burst_tflops": flop / (best * 1e-3) / 1e12,
                      "sustained_tflops": flop / (sus * 1e-3) / 1e12, "sustained_reps": reps,
                      "sm_mhz,max,power_w under load": clk}))


if __name__ == "__main__":
    main()
