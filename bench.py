#!/usr/bin/env python3
"""bench.py -- the rotation-finding hot path of fjdf/CSA (`./CSA R`) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mammals|sets32|variants256|bacterial]
                    [--sets S] [--impl reference] [--only-headline]

The headline (metric/value/e2e/roofline) is BASELINE.json's configs[1]; the same line carries, under "workloads", the other
configs (configs[3] sets of 32, configs[2] 256 variants, configs[4] 16 x 5 Mb -- with N > 1 its suffix-array buckets
sharded over the ranks, strong scaling) with value / ms_per_step / e2e / dominant kernel each, and under "single_set" what ONE
call on one Mammals-shaped set costs.

A step = one pass of the whole path (suffix array, LCP, common blocks, block order, chaining,
rotations) over one batch of S independent synthetic sequence sets per GPU.
  value : circular bases/s with the batch already resident in HBM (csa_gpu_batch_run only)
  e2e   : the same through the C ABI with HOST buffers: csa_gpu_batch_upload_flat + run + download
          per step, host->device and device->host copies inside the timed region; two contexts fed by
          two host threads (--e2e-contexts), so one batch's copy runs under the other batch's kernels
          (the runs themselves take turns)
  roofline : the kernel with the largest share of device time, timed with CUDA events on the launch
          stream in a separate profiled pass (csa_gpu_profile_*), against MEASURED_PEAKS.json
  cpu_baseline : the UNMODIFIED reference binary (oracle/_ref/CSA_ref, compiled from the reference's
          own sources) on a bounded sample of the same sets, on this box's host cores
Sets are independent, so N GPUs shard them with no data-path collective ("weak": S sets per GPU).
`--impl reference` times only the reference's CPU implementation (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "circular bases/sec"
UNIT = "bases/s"
DEFAULT_SETS = {"mammals": 480, "sets32": 192, "variants256": 24, "bacterial": 1}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---- clocks during the timed region ---------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons while the timed steps run: NVML polled every ~10 ms from a thread
    (the timed region of a short run is shorter than one `nvidia-smi -lms` period; polling faster than
    this steals time from the launching thread when eight ranks share the host)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, cuda_index):
        self.samples, self.stop_flag, self.thread, self.h, self.nv = [], False, None, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(cuda_index)
            try:
                bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
        except Exception as e:  # no NVML: say so in the line
            self.err = repr(e)

    def _poll(self):
        nv, h = self.nv, self.h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(get_reasons(h)),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.010)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self, t0, t1):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        rows = [r for r in self.samples if t0 <= r[0] <= t1] or self.samples
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": statistics.median(r[1] for r in rows),
                "sm_max_mhz": self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                "power_w_max": max(r[3] for r in rows), "samples": len(rows),
                "reasons": [name for bit, name in self.REASONS.items() if bits & bit]}


# ---- the reference's CPU implementation ---------------------------------------------------------------
def cpu_reference_rate(sets, cores):
    """Times the reference on `sets` (lists of bytes), `cores` processes side by side.
    oracle/_ref/CSA_ref = the unmodified reference compiled from its own sources ("reference");
    without it, oracle/_build/csa_oracle = the restatement ("port")."""
    ref = os.path.join(ROOT, "oracle", "_ref", "CSA_ref")
    port = os.path.join(ROOT, "oracle", "_build", "csa_oracle")
    if os.path.exists(ref):
        binary, kind = ref, "reference"
    elif os.path.exists(port):
        binary, kind = port, "port"
    else:
        return None
    tmp = tempfile.mkdtemp(prefix="csa_cpu_")
    try:
        dirs = []
        for i, seqs in enumerate(sets):
            d = os.path.join(tmp, str(i))
            os.mkdir(d)
            with open(os.path.join(d, "in.fa"), "wb") as f:
                for k, s in enumerate(seqs):
                    f.write(b">s%d\n" % k + s + b"\n")
            dirs.append(d)
        def one(d):
            return subprocess.run([binary, "R", "in.fa"], cwd=d, stdin=subprocess.DEVNULL,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            rcs = list(ex.map(one, dirs))
        dt = time.perf_counter() - t0
        bases = sum(len(s) for seqs in sets for s in seqs)
        return {"value": bases / dt, "unit": UNIT, "cores": cores, "kind": kind, "seconds": dt,
                "failed_sets": sum(1 for r in rcs if r != 0),
                "sample": f"{len(sets)} sets ({bases} bases) of the same workload through `{os.path.basename(binary)} R` "
                          f"(whole CLI: FASTA load, generalized cyclic suffix tree, analyzeTree, output files), "
                          f"{cores} processes side by side"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mammals", choices=list(DEFAULT_SETS))
    ap.add_argument("--sets", type=int, default=0, help="sets per GPU per step")
    ap.add_argument("--cpu-sets", type=int, default=0, help="sets of the CPU sample (default 12 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the other BASELINE configs and the single-set latency")
    ap.add_argument("--e2e-contexts", type=int, default=2, help="contexts (host threads) feeding the GPU in the e2e loop")
    a = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    nsets = a.sets or DEFAULT_SETS[a.workload]
    from csa_b200.workloads import WORKLOADS, batch_sets, workload_batch
    what = WORKLOADS[a.workload][5]
    cores = len(os.sched_getaffinity(0))

    if a.impl == "reference":
        if rank != 0:
            return 0
        per_step = a.cpu_sets or max(cores, min(nsets, cores * 4))
        batch = workload_batch(a.workload, per_step, seed=1000)
        sets = batch_sets(batch)
        for _ in range(max(0, min(a.warmup, 1))):
            cpu_reference_rate(sets[:cores], cores)
        t, bases, r = 0.0, 0, None
        for _ in range(a.steps):
            r = cpu_reference_rate(sets, cores)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/CSA_ref and oracle/_build/csa_oracle not built"}))
                return 0
            t += r["seconds"]
            bases += batch.nbases
        v = bases / t
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": what, "sets_per_step": per_step, "bases_per_step": batch.nbases},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": r["sample"]},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0

    # rank 0 prints ONE line on stdout: anything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from csa_b200.api import RotationFinder
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    e2e_contexts = a.e2e_contexts

    def measure(workload, nsets, steps, warmup):
        """one workload: resident value, e2e through the C ABI with host buffers, per-kernel profile"""
        # One bacterial-scale set does not split into independent sets: with N > 1 its suffix-array stage is sharded by
        # buckets (csa_gpu_shard_*, csa_b200/shard.py): every rank holds the same set, total work is fixed ("strong").
        buckets = workload == "bacterial" and world > 1
        batch = workload_batch(workload, nsets, seed=1000 + (0 if buckets else rank))  # every rank its own sets
        rf = RotationFinder(device=local)
        if os.environ.get("CSA_BENCH_MODE"):  # (experiments: csa_gpu_debug_rounds, one of the equivalent suffix-array paths forced)
            rf.debug_rounds(int(os.environ["CSA_BENCH_MODE"]))
        stream = torch.cuda.current_stream()
        rf.set_stream(stream.cuda_stream)
        if buckets:
            from csa_b200.shard import run_bucket_sharded
            rf_run = lambda: run_bucket_sharded(rf, rank, world, dist)
        else:
            rf_run = rf.run

        # ---- value: inputs resident in HBM, csa_gpu_batch_run only ----
        rf.upload(batch)
        for _ in range(warmup):
            rf_run()
        clocks = ClockSampler(local)
        barrier()
        clocks.start()
        t_wall0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches = 0
        stage_ms = [0.0] * 6
        for _ in range(steps):
            rf_run()
            ms, l = rf.timings()
            launches += l
            stage_ms = [x + y for x, y in zip(stage_ms, ms)]
        e1.record(stream)
        barrier()
        t_wall1 = time.perf_counter()
        dev_ms = allmax(e0.elapsed_time(e1))
        clk = clocks.stop(t_wall0, t_wall1)
        total_bases = (batch.nbases if buckets else allsum(batch.nbases)) * steps
        value = total_bases / (dev_ms / 1e3)
        rot, info = rf.download()
        ok_sets = sum(1 for i in info if i.status == 0)

        # ---- e2e: host buffers in, rotations out, through the C ABI ----
        pinned = rf.pin(batch)  # "from pinned host memory": the copy engine reads the caller's buffer in place
        # Independent sets: two contexts on the GPU, each fed by its own host thread through the same three C-ABI calls
        # (upload -> run -> download, every step its own copies), so that one batch's host->device copy runs under the
        # other batch's kernels -- what a service in front of the library does.  One large set sharded over ranks: one
        # context, the steps one after the other.
        nctx = 1 if buckets else max(1, e2e_contexts)
        ctxs = [rf]
        for _ in range(nctx - 1):
            r2 = RotationFinder(device=local)
            s2 = torch.cuda.Stream()
            r2.set_stream(s2.cuda_stream)
            r2._stream_keepalive = s2
            ctxs.append(r2)
        for c2 in ctxs:
            for _ in range(max(1, warmup // 2)):
                c2.upload(batch); (rf_run() if c2 is rf else c2.run()); c2.download()
        last = [None] * nctx
        run_turn = threading.Lock()  # one batch's kernels at a time (two batches' kernels side by side only evict each other's
                                     # packed text from L2: 14.7 ms a step against 13.7); the OTHER context's copies run beside them
        def feed(i, nsteps):
            torch.cuda.set_device(local)
            for _ in range(nsteps):
                ctxs[i].upload(batch)
                with run_turn:
                    rf_run() if (buckets and i == 0) else ctxs[i].run()
                last[i] = ctxs[i].download()[0]
        share = [steps // nctx + (1 if i < steps % nctx else 0) for i in range(nctx)]
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        if nctx == 1:
            feed(0, steps)
        else:
            th = [threading.Thread(target=feed, args=(i, share[i])) for i in range(nctx)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()  # every context's stream is idle: the event below closes the whole region
        e3.record(stream)
        barrier()
        e2e_ms = allmax(e2.elapsed_time(e3))
        rf.unpin(pinned)
        for i in range(nctx):
            if share[i]:
                assert np.array_equal(rot, last[i])
        for c2 in ctxs[1:]:
            c2.close()
        if buckets:  # the bucket-sharded run against the same set on this GPU alone: rotations and the whole block list
            rf.upload(batch); rf_run()
            sharded_blocks = rf.blocks()
            rf.upload(batch); rf.run()
            rot3, _ = rf.download()
            assert np.array_equal(rot, rot3), "bucket-sharded rotations differ from the single-GPU run"
            alone = rf.blocks()
            assert all(np.array_equal(x, y) for x, y in zip(sharded_blocks[0], alone[0])) and np.array_equal(sharded_blocks[1], alone[1]), \
                "bucket-sharded blocks differ from the single-GPU run"
        h2d = batch.nbases + 4 * (2 * batch.nseqs + 4 * batch.nsets + 8) + 8 * (batch.nseqs + 1)
        d2h = 4 * batch.nseqs + 3 * 4 * batch.nsets
        e2e_value = total_bases / (e2e_ms / 1e3)

        # ---- roofline: a separate profiled pass, CUDA events around every launch ----
        rf.profile_enable(True)
        rf_run()
        rows = rf.profile()
        rf.profile_enable(False)
        roofline, kernels = None, []
        if rows:
            tot = sum(r[2] for r in rows)
            rows.sort(key=lambda r: -r[2])
            if os.environ.get("CSA_BENCH_ALL_KERNELS"):
                for name, n, ms, by in rows:
                    print("  %-18s n=%-4d %8.3f ms %5.1f%% %8.1f GB/s" % (name, n, ms, 100 * ms / tot, by / ms / 1e6 if ms > 0 else 0), file=sys.stderr)
            for name, n, ms, by in rows[:12]:
                kernels.append({"kernel": name, "launches": n, "ms": round(ms, 3), "share": round(ms / tot, 4),
                                "algorithmic_GBps": round(by / ms / 1e6, 1) if ms > 0 else None,
                                "frac": round(by / ms / 1e6 / peak, 3) if ms > 0 else None})
            name, n, ms, by = rows[0]
            ach = by / ms / 1e6
            # DRAM bytes per launch from the committed ncu capture of this kernel (profiles/traffic.json holds
            # dram__bytes_read+write per suffix), scaled to this run's launch size
            traffic, traffic_src = None, None
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
                if tr and abs(batch.nbases - tr["items"]) <= 0.02 * tr["items"]:  # measured at this launch size, not scaled across the L2 cliff
                    traffic, traffic_src = tr["dram_bytes_per_item"] * batch.nbases, tr["source"]
            except (OSError, ValueError, KeyError):
                pass
            roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": n, "avg_launch_ms": ms / n,
                        "algorithmic_bytes_per_launch": by / n, "share_of_step": ms / tot}


        res = {"value": value, "ms_per_step": dev_ms / steps, "e2e_value": e2e_value, "e2e_ms_per_step": e2e_ms / steps, "h2d": h2d, "d2h": d2h,
               "nctx": nctx, "launches": launches, "clk": clk, "roofline": roofline, "kernels": kernels, "ok_sets": ok_sets,
               "stage_ms": [round(x / steps, 3) for x in stage_ms], "buckets": buckets, "batch": batch, "steps": steps,
               "shard_path": (sys.modules["csa_b200.shard"].last_path if buckets else None)}
        rf.close()
        return res

    head = measure(a.workload, nsets, a.steps, a.warmup)
    batch = head["batch"]

    # ---- every other BASELINE config, fewer steps each (parity of their rotations is the -m gpu tests' business) ----
    others = {}
    if not a.only_headline:
        k_other, w_other = max(3, a.steps // 4), max(3, min(a.warmup, 3))
        for name in DEFAULT_SETS:
            if name == a.workload:
                continue
            r = measure(name, DEFAULT_SETS[name], k_other, w_other)
            top = r["roofline"] or {}
            others[name] = {"workload": WORKLOADS[name][5], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                            "scaling": "strong" if r["buckets"] else "weak", "sets_per_gpu_per_step": DEFAULT_SETS[name],
                            "bases_per_gpu_per_step": r["batch"].nbases, "sets_ok": r["ok_sets"],
                            "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["e2e_ms_per_step"],
                                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
                            "gpu_launches": r["launches"], "stage_ms_per_step": r["stage_ms"],
                            "dominant_kernel": {k2: top.get(k2) for k2 in ("kernel", "frac", "achieved", "share_of_step", "avg_launch_ms")},
                            "kernels": r["kernels"][:6],
                            "parallelism": (f"one set, suffix-array buckets sharded over {world} GPUs (csa_gpu_shard_*; path taken after the bucket "
                                            f"sort: {r['shard_path']}), rotations and blocks checked against the one-GPU run") if r["buckets"] else f"sets sharded over {world} GPU(s), no collective"}
            del r

    # ---- what ONE call of the drop-in costs: csa_gpu_find_rotations on one Mammals-shaped set, host buffers in, rotations out ----
    single = None
    if rank == 0:
        one = batch_sets(workload_batch("mammals", 1, seed=1000))[0]
        rf1 = RotationFinder(device=local)
        for _ in range(3):
            rf1.find_rotations(one, with_blocks=False)
        ts = []
        for _ in range(15):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r1 = rf1.find_rotations(one, with_blocks=False)
            ts.append(1e3 * (time.perf_counter() - t0))
        _, l1 = rf1.timings()
        rf1.close()
        single = {"ms": statistics.median(ts), "min_ms": min(ts), "calls": len(ts), "bases": sum(len(x) for x in one), "gpu_launches": l1,
                  "what": "wall clock of one csa_gpu_find_rotations (upload + all kernels + download) on ONE Mammals-shaped set of 12 mitogenomes, "
                          "host buffers in, rotations out: what `./CSA R` spends in the library per call (the reference needs ~0.3 s for the same set)"}

    # ---- the reference on this box's host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        k = a.cpu_sets or min(nsets, cores * 24)
        r = cpu_reference_rate(batch_sets(batch, 0, k), cores)
        if r is not None:
            cpu = {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            cpu["seconds"] = round(r["seconds"], 2)

    if rank == 0:
        buckets = head["buckets"]
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong" if buckets else "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": what, "sets_per_gpu_per_step": nsets, "bases_per_gpu_per_step": batch.nbases,
                           "sequences_per_set": int(batch.set_start[1]),
                           "parallelism": (f"one set, suffix-array buckets sharded over {world} GPUs (csa_gpu_shard_*; path taken after the bucket "
                                           f"sort: {head['shard_path']}), rotations and blocks checked against the one-GPU run") if buckets
                                          else f"sets sharded over {world} GPU(s), no collective",
                           "l2": "per-step working set (~55 B/base) far above the 126 MB L2; no flush needed",
                           "sets_ok": head["ok_sets"], "stage_ms_per_step": head["stage_ms"],
                           "stages": ["suffix array", "lcp", "common blocks", "block order", "chaining+rotations", "whole run"]},
                "e2e": {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                        "ms_per_step": head["e2e_ms_per_step"], "contexts": head["nctx"]},
                "gpu_launches": head["launches"], "clocks": head["clk"], "roofline": head["roofline"], "cpu_baseline": cpu,
                "kernels": head["kernels"], "workloads": others, "single_set": single}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
