#!/usr/bin/env python3
"""Per-kernel device times of one workload (CUDA events around every launch, csa_gpu_profile_*), the table the
roofline numbers in DESIGN.md come from.   python profiles/kbench.py mammals 480 [sets32 192 ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csa_b200.api import RotationFinder
from csa_b200.workloads import workload_batch

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6549.1


def main():
    args = sys.argv[1:] or ["mammals", "480"]
    out = {}
    for name, nsets in zip(args[::2], args[1::2]):
        batch = workload_batch(name, int(nsets), seed=1000)
        rf = RotationFinder(device=0)
        if os.environ.get("KBENCH_MODE"):
            rf.debug_rounds(int(os.environ["KBENCH_MODE"]))  # csa_gpu_debug_rounds: force one of the equivalent suffix-array paths
        rf.upload(batch)
        for _ in range(3):
            rf.run()
        ms = []
        for _ in range(5):
            rf.run()
            ms.append(rf.timings()[0])
        best = min(ms, key=lambda m: m[5])
        rf.profile_enable(True)
        rf.run()
        rows = sorted(rf.profile(), key=lambda r: -r[2])
        rf.profile_enable(False)
        tot = sum(r[2] for r in rows)
        print(f"== {name} x {nsets}: {batch.nbases} bases, whole run {best[5]:.3f} ms = {batch.nbases / best[5] / 1e6:.3f} Gbases/s; stages {[round(x, 2) for x in best]}; profiled sum {tot:.2f} ms")
        for k, n, t, by in rows[:16]:
            print("  %-18s n=%-4d %8.3f ms %5.1f%% %8.1f GB/s  frac %.3f" % (k, n, t, 100 * t / tot, by / t / 1e6 if t > 0 else 0, by / t / 1e6 / PEAK if t > 0 else 0))
        out[name] = {"nsets": int(nsets), "bases": batch.nbases, "ms": best[5], "stages": best, "kernels": [(k, n, round(t, 4), by) for k, n, t, by in rows]}
        rf.close()
    if os.environ.get("KBENCH_JSON"):
        json.dump(out, open(os.environ["KBENCH_JSON"], "w"))


if __name__ == "__main__":
    main()
