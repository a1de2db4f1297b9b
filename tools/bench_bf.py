#!/usr/bin/env python
"""BASELINE configs C1/C2 timing: BruteForceSearcher (f32) and ScalarQuantizedBruteForceSearcher (int8) batched
search on synthetic Gaussian data, tensor-core ranking path vs the CUDA-core path (SCANN_BF_NO_TC=1).
Prints one JSON line per case; tensor fraction is quoted on 2*nq*n*dim flops of the FILTER pass against
MEASURED_PEAKS.json bf16_tflops (burst)."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=1_000_000)
    p.add_argument("--dim", type=int, default=128)
    p.add_argument("--nq", type=int, default=10_000)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--legacy", action="store_true", help="also time the CUDA-core path")
    a = p.parse_args()
    import torch
    pkg = importlib.import_module("scann-rust_b200")
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    x = torch.randn((a.n, a.dim), generator=g, device=dev)
    g.manual_seed(123)
    q = torch.randn((a.nq, a.dim), generator=g, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))

    def run(name, make):
        s = make()
        for _ in range(2):
            s.search_batched(q, a.k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ids, d, c = s.search_batched(q, a.k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        tf = 2.0 * a.nq * a.n * a.dim / (ms / 1e3) / 1e12
        out = {"case": name, "n": a.n, "dim": a.dim, "nq": a.nq, "k": a.k, "ms_per_batch": ms,
               "queries_per_s": a.nq / (ms / 1e3), "tflops_equiv": tf, "frac_of_measured_bf16_peak": tf / peak_tf,
               "path_stats": s.path_stats()}
        print(json.dumps(out), flush=True)
        s.close()
        return ids

    M = pkg.DistanceMeasure
    modes = [("tc", "0")] + ([("cuda-core", "1")] if a.legacy else [])
    codes, cal = pkg.scalar_quantize(x)
    for tag, env in modes:
        os.environ["SCANN_BF_NO_TC"] = env
        run(f"bf_f32_dot[{tag}]", lambda: pkg.BruteForceSearcher(x, M.DotProduct))
        run(f"bf_f32_sql2[{tag}]", lambda: pkg.BruteForceSearcher(x, M.SquaredL2))
        run(f"sq8_dot[{tag}]", lambda: pkg.ScalarQuantizedBruteForceSearcher.from_quantized(
            codes, float(cal[2]), M.DotProduct))


if __name__ == "__main__":
    main()
