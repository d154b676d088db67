#!/usr/bin/env python
"""Benchmark of the orgscorer hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the whole hot path (K1..K10) over one batch of synthetic contigs of
BASELINE.json's configs[1] shape (cfg2: 100k contigs, 2-8 genes, ~50 hits/gene, 5k species).
`value`  : contigs/s, device-timed (CUDA events on the engine's stream), inputs resident in HBM.
`e2e`    : same metric through the C-ABI plugin call with pinned HOST buffers (H2D + kernels + D2H).
N > 1    : one process per GPU (torchrun), every rank scores its own shard of the same shape
           (weak scaling, no data-path collective; compacted results are gathered over NCCL in
           the e2e leg), time = max over ranks.
--impl reference : the CPU restatement of the reference algorithm (oracle/, numpy, one process per
           host core) on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contigs/sec scored+classified"
UNIT = "contigs/s"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--contigs", type=int, default=None, help="contigs per GPU (default: config size)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="contigs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exact", action="store_true", help="exact pipeline for every contig (bit-exact crit / rank)")
    ap.add_argument("--wide", action="store_true", help="e2e leg with the wide 29 B/hit wire format")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (wfl_set_option)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle (a faithful restatement of the reference's Python/numpy code),
# one process per host core, contigs split into contiguous shards (contigs are independent).
# --------------------------------------------------------------------------------------------

def _oracle_worker(job):
    from oracle import orgscorer_oracle as oracle
    params, tax, arrays = job
    t = time.perf_counter()
    out = oracle.score_batch(params, tax, arrays)
    return time.perf_counter() - t, out["call_counts"].tolist()


def cpu_baseline_run(batch, params, tax, n_sample, cores):
    """Wall-clock contigs/s of the oracle over `n_sample` contigs using `cores` processes."""
    import multiprocessing as mp
    n_sample = min(n_sample, batch.n_contigs)
    cores = max(1, min(cores, n_sample))
    cuts = np.linspace(0, n_sample, cores + 1).astype(int)
    jobs = [(params.as_dict(), tax.tables(), batch.slice(int(a), int(b)).arrays())
            for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(jobs)) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    return n_sample / wall, wall, len(jobs), res


def make_workload(args, rank):
    from waafle_b200 import synth
    from waafle_b200.params import OrgscorerParams
    data = synth.generate_config(args.workload, n_contigs=args.contigs, seed=1000 + rank)
    tax = data.taxonomy()
    batch = data.to_batch(tax)
    params = OrgscorerParams(n_systems=1 if batch.hit_sysmask is not None else 0)
    return batch, params, tax


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def pinned_like(a):
    import torch
    t = torch.empty(a.shape, dtype=getattr(torch, str(a.dtype)), pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


def run_reference(args, rank, world):
    """CPU arm: rank 0 alone times the oracle port on a bounded sample of the workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = {"cfg2": 400, "cfg3": 150, "cfg5": 150, "cfg4": 2}.get(args.workload, 100)
    n_sample = args.cpu_sample or cores * per_core
    args.contigs = args.contigs or None
    a2 = argparse.Namespace(**vars(args))
    a2.contigs = max(n_sample, 64)
    batch, params, tax = make_workload(a2, 0)
    times = []
    for step in range(args.warmup + args.steps):
        rate, wall, used, _ = cpu_baseline_run(batch, params, tax, n_sample, cores)
        if step >= args.warmup:
            times.append(wall)
    T = sum(times)
    value = n_sample * args.steps / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_contigs_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": "{} contigs of {} per step, numpy oracle (restatement of the "
                                   "reference's Python/numpy orgscorer), one process per core".format(
                                       n_sample, args.workload)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args):
    from waafle_b200 import synth
    n = args.contigs or synth.CONFIGS[args.workload]["n_contigs"]
    desc = {"cfg2": "synthetic {} contigs, 2-8 genes, ~50 hits/gene, 5k-species taxonomy (BASELINE configs[1])",
            "cfg3": "synthetic {} contigs, 2-20 genes, 8-level taxonomy (BASELINE configs[2] shape)",
            "cfg4": "long-contig stress {} contigs, 100-130 genes, 550-species pools (BASELINE configs[3] shape)",
            "cfg5": "Prodigal-style {} contigs, short/empty loci, annotations (BASELINE configs[4] shape)"}
    return desc[args.workload].format(n)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    batch, params, tax = make_workload(args, rank)

    # ---- CPU baseline beside the GPU numbers (rank 0, N=1 only, before any CUDA init) ----
    cpu_baseline, c_ref = None, None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        per_core = {"cfg2": 1500, "cfg3": 600, "cfg5": 600, "cfg4": 4}.get(args.workload, 300)
        n_sample = args.cpu_sample or cores * per_core
        rate, wall, used, _ = cpu_baseline_run(batch, params, tax, n_sample, cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
                        "sample": "first {} contigs of the workload, numpy oracle, {} processes, {:.1f} s wall"
                                  .format(min(n_sample, batch.n_contigs), used, wall)}

        # second CPU arm: the C restatement (oracle/orgscorer_oracle.c), one thread per core, on the WHOLE
        # workload; its output doubles as the full-size parity reference for the GPU results below
        try:
            from oracle import c_oracle
            tc = time.perf_counter()
            c_ref = c_oracle.score_batch(params, tax, batch, threads=cores)
            wall_c = time.perf_counter() - tc
            cpu_baseline["c_port"] = {"value": batch.n_contigs / wall_c, "unit": UNIT, "cores": cores,
                                      "sample": "all {} contigs, C restatement of the reference, {} threads, "
                                                "{:.1f} s wall".format(batch.n_contigs, cores, wall_c)}
        except Exception as exc:   # the checker failing to build must not hide the GPU numbers
            cpu_baseline["c_port"] = {"unavailable": repr(exc)[:200]}

    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from waafle_b200.engine import Engine
    eng = Engine(local_rank, params, tax)
    if args.exact:
        eng.set_option("exact", 1)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ("value") ----
    eng.upload(batch)
    for _ in range(args.warmup):
        eng.run_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, score_ms, launches = 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run_resident()
        st = eng.stats()
        dev_ms += st["ms_kernels"]
        score_ms += st["ms_score_kernel"]
        launches += st["kernel_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    res = eng.download()
    st = eng.stats()
    parity = None
    if c_ref is not None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import compare_results
        rtol = 0.0 if args.exact else 1e-12
        diffs = compare_results(c_ref, res, score_rtol=rtol)
        parity = {"checked_contigs": batch.n_contigs, "against": "oracle/orgscorer_oracle.c",
                  "bit_exact": not diffs, "score_rtol": rtol,
                  "what": "calls, clades, synteny, loci flags, members, annotation winners bit-exact; crit / rank within score_rtol",
                  "diffs": [str(d)[:160] for d in diffs[:3]]}

    # ---- end-to-end leg: pinned host buffers through the plugin call ----
    from waafle_b200.engine import PinnedArena
    pin = PinnedArena()
    packed = (not args.wide) and batch.can_pack(len(tax.tables()["parent"]), params.n_systems)
    wire = batch.to_packed(params.min_scov) if packed else batch.arrays()
    harr = {k: pin.like(np.ascontiguousarray(v)) for k, v in wire.items()}
    eng.use_pinned_results(True)
    h2d_bytes = int(sum(v.nbytes for v in harr.values()))
    out = eng.score_batch(harr)
    d2h_bytes = int(sum(np.asarray(v).nbytes for v in out.values()))
    for _ in range(max(1, args.warmup - 1)):
        eng.score_batch(harr)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.score_batch(harr)
        if dist is not None:
            # only the compacted call counts cross NVLink in the timed loop
            cc = torch.from_numpy(out["call_counts"]).cuda()
            dist.all_reduce(cc)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t1)
    st_e2e = eng.stats()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- max over ranks ----
    tv = torch.tensor([dev_ms, score_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    dev_ms, score_ms, e2e_ms, wall_ms = tv.tolist()
    n_total = batch.n_contigs * world

    if rank == 0:
        # roofline of the scoring pipeline (every wfl_pipe_* launch of one step): algorithmic bytes per step
        # = 29 B/hit + 9 B/locus + 16 B/contig in, 40 B/contig + G(1+4S) B out,
        # plus 29 B/hit again for every extra taxonomy level a contig is evaluated at.
        H = np.diff(batch.hit_off)
        extra_levels = np.maximum(res["lifts"] - params.jump_taxonomy, 0)
        alg_bytes = batch.algorithmic_bytes(params.n_systems) + 29 * int((H * extra_levels).sum())
        peak, peak_src = FALLBACK_HBM_GBS, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured"
        except Exception:
            pass
        achieved = alg_bytes / (score_ms / args.steps * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                tj = json.load(fh)
                if tj.get("workload") == args.workload and tj.get("contigs") == batch.n_contigs:
                    traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": n_total * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "contigs_per_gpu": batch.n_contigs,
                       "hits_per_gpu": batch.n_hits, "loci_per_gpu": batch.n_loci,
                       "flags": "reference defaults (k1=0.5 k2=0.8 meld/meld range=0.05 weak-loci=ignore)",
                       "parallelism": "contig-sharded x{}".format(world),
                       "l2": "inputs ({:.0f} MB/GPU) exceed the 126 MB L2; no flush needed".format(
                           batch.algorithmic_bytes(params.n_systems) / 1e6),
                       "mean_levels_per_contig": float(st["levels"]) / max(1, batch.n_contigs),
                       "wall_ms_per_step": wall_ms / args.steps},
            "e2e": {"value": n_total * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "wire_format": "packed 14 B/hit" if packed else "wide 29 B/hit",
                    "last_call_ms": {"h2d_window": st_e2e["ms_h2d"], "kernels_window": st_e2e["ms_kernels"],
                                     "d2h": st_e2e["ms_d2h"]}},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "wfl_pipe_* (all launches of one step: prepare, then regroup / K2 sort / k2 / masks / one / two / lift per taxonomy level; the K2 kernel wfl_pipe_k2 is ~30% of it)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel_ms": score_ms / args.steps},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "clocks": sampler.summary(),
            "calls": {"lgt": int(res["call_counts"][0]), "no_lgt": int(res["call_counts"][1]),
                      "unclassified": int(res["call_counts"][2])},
            "engine_stats": {k: st[k] for k in ("matched_pairs", "groups", "levels", "pairs_tested",
                                                "pairs_scored", "workspace_retries", "smem_contigs", "fallback_contigs",
                                                "guard_trips", "refined_groups", "host_syncs")},
        }
        print(json.dumps(line))
    eng.close()
    pin.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
