#!/usr/bin/env python
"""bench.py — aligned query bp/s of the LexicMap query-side search path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): BASELINE.json configs[1] — 10,000 synthetic 1-kb queries vs a 1,000-genome synthetic index
(50 families x 20 members x 1 Mbp, SURVEY.md §8d generators, seeds 20260924/20260925). One "step" = one pass of the whole hot
path (sketch -> seed probe -> chain -> pseudo-align -> extend+WFA -> rows) over the 10k-query batch.
N>1 (torchrun, one rank per GPU): index image replicated, every rank searches its own 10k-query batch (weak scaling), no
data-path collective; one NCCL all-reduce of the per-rank hit/row counters at the end of the timed region.

`value`   = sum of query bases / CUDA-event time of lmg_search_staged (queries already in HBM), max over ranks.
`e2e`     = same metric through lmg_search_batch with HOST buffers (H2D of the queries and D2H of all rows inside the timed region).
`roofline`= seed-lookup kernel (k_probe_find): algorithmic bytes / its CUDA-event time / measured HBM peak.
`cpu_baseline` = the C++ oracle port of the reference path (the Go reference cannot run here: no Go toolchain) on the host cores.
--impl reference times that same CPU port with all host threads on the same workload.
"""
import argparse
import json
import os

# torchrun exports OMP_NUM_THREADS=1; the host side of the library (and the index builder) use OpenMP for list handling, so give every
# rank its share of the cores before any OpenMP runtime is loaded
_world = int(os.environ.get("WORLD_SIZE", 1))
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, _world)))
os.environ.setdefault("NCCL_DEBUG", "WARN")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORK = os.environ.get("LMG_BENCH_DIR", "/tmp/lmg_bench")
CFG = dict(families=int(os.environ.get("LMG_BENCH_FAMILIES", 50)), members=int(os.environ.get("LMG_BENCH_MEMBERS", 20)), genome_len=int(os.environ.get("LMG_BENCH_GLEN", 1000000)),
           n_queries=int(os.environ.get("LMG_BENCH_NQ", 10000)), query_len=int(os.environ.get("LMG_BENCH_QLEN", 1000)), genome_seed=20260924, query_seed=20260925)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def ensure_workload(rank):
    from lexicmap_b200 import build
    tools = build.build_tools()
    name = "c2_%dx%dx%d" % (CFG["families"], CFG["members"], CFG["genome_len"])
    idx = os.path.join(WORK, name + ".lmi")
    os.makedirs(WORK, exist_ok=True)
    if not os.path.exists(os.path.join(idx, "info.toml")):
        t = time.time()
        tmp = idx + ".tmp%d" % os.getpid()
        subprocess.check_call([tools, "index", "--synth", "%d,%d,%d,%d,20" % (CFG["families"], CFG["members"], CFG["genome_len"], CFG["genome_seed"]), "--out", tmp])
        if os.path.isdir(idx):   # an interrupted earlier build (no info.toml)
            import shutil
            shutil.rmtree(idx)
        os.rename(tmp, idx)
        log("index built in %.1fs -> %s" % (time.time() - t, idx))
    qf = os.path.join(WORK, "%s_q%d_%d_r%d.fa" % (name, CFG["n_queries"], CFG["query_len"], rank))
    if not os.path.exists(qf):
        subprocess.check_call([tools, "synth-queries", "--index", idx, "--n", str(CFG["n_queries"]), "--len", str(CFG["query_len"]), "--seed", str(CFG["query_seed"] + rank), "--out", qf + ".tmp"])
        os.rename(qf + ".tmp", qf)
    return idx, qf


class ClockSampler:
    """One long-lived `nvidia-smi -lms 200` (the profiling recipe's clocks line) running during the timed region; parsed afterwards.
    A single process instead of one spawn per sample keeps the driver's management lock out of the way of the timed CUDA calls."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu):
        self.gpu, self.proc, self.sm, self.max_sm, self.reasons = gpu, None, [], 0, set()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def finish(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
            out = ""
        for line in out.splitlines():
            o = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(o[0]))
                self.max_sm = float(o[1])
            except Exception:
                continue
            for n, v in zip(self.NAMES, o[2:]):
                if "Active" in v and "Not" not in v:
                    self.reasons.add(n)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm or None, "reasons": sorted(self.reasons)}


def probe_algorithmic_bytes(cnt):
    """SURVEY.md §8(d) / DESIGN.md §4 for k_probe_find2, sector-granular because the access is random. Per surviving probe: its 24-byte
    record + the 32-byte sector of its anchor-table entry; 32 B per search step actually taken (gallop + binary search over the bucket's
    keys); 16 B per entry scanned in the matching range (key + first value); 48 B per hit record written."""
    surv, steps, entries, hits = int(cnt[1]), int(cnt[2]), int(cnt[3]), int(cnt[4])
    return surv * (24 + 32) + steps * 32 + entries * 16 + hits * 48


def cpu_port_throughput(idx_dir, seqs, threads, target_s=12.0):
    from oracle_binding import Oracle
    o = Oracle(idx_dir)
    n0 = min(len(seqs), max(threads, 64))
    t = time.time()
    o.search(seqs[:n0], threads=threads)
    dt = max(time.time() - t, 1e-3)
    n = int(min(len(seqs), max(n0, n0 * target_s / dt)))
    t = time.time()
    rows, _, _ = o.search(seqs[:n], threads=threads)
    dt = time.time() - t
    bp = sum(len(s) for s in seqs[:n])
    o.close()
    return bp / dt, n, dt, len(rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    is_c2 = (CFG["families"], CFG["members"], CFG["genome_len"], CFG["n_queries"], CFG["query_len"]) == (50, 20, 1000000, 10000, 1000)
    workload = "%d synthetic %d-bp queries vs %d-genome synthetic index (%dx%dx%d bp), %s" % (
        CFG["n_queries"], CFG["query_len"], CFG["families"] * CFG["members"], CFG["families"], CFG["members"], CFG["genome_len"],
        "BASELINE.json configs[1]" if is_c2 else "LMG_BENCH_* override (not a BASELINE.json config as is)")
    config = {"workload": workload, "queries_per_gpu": CFG["n_queries"], "query_len": CFG["query_len"], "genomes": CFG["families"] * CFG["members"], "masks": 20000,
              "sharding": "by query, index replicated" if a.gpus > 1 else "single GPU", "l2": "index image (>1 GB) and per-batch buffers exceed the 126 MB L2; no explicit flush",
              "seeds": [CFG["genome_seed"], CFG["query_seed"]],
              "index": "built by lmi-tools with first-round LexicHash seeds only (the writer's --fill-deserts, the reference's default, is off; it would add ~34 % seed values on these genomes)"}
    from oracle_binding import read_fasta

    if a.impl == "reference":
        if rank != 0:
            return
        idx_dir, qf = ensure_workload(0)
        ids, seqs = read_fasta(qf)
        threads = os.cpu_count() or 1
        vals = []
        for i in range(a.warmup + a.steps):
            bps, n, dt, nrows = cpu_port_throughput(idx_dir, seqs, threads, target_s=8.0)
            if i >= a.warmup:
                vals.append((bps, n, dt))
        bps = float(np.mean([v[0] for v in vals]))
        n, dt = vals[-1][1], float(np.mean([v[2] for v in vals]))
        print(json.dumps({"impl": "reference", "metric": "aligned query bp/s", "value": bps, "unit": "bp/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": bps, "unit": "bp/s", "cores": threads, "kind": "port", "sample": "%d of the %d queries per step (C++ port of the reference path; Go toolchain absent)" % (n, len(seqs))},
                          "e2e": {"value": bps, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import lexicmap_b200
    from lexicmap_b200.api import pack_queries
    dist = None
    real_stdout = os.dup(1)
    os.dup2(2, 1)   # libraries (NCCL banner) must not pollute the one JSON line on stdout
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ensure_workload(0)
    if dist:
        dist.barrier()
    idx_dir, qf = ensure_workload(rank)
    ids, seqs = read_fasta(qf)
    t0 = time.time()
    idx = lexicmap_b200.Index(idx_dir, device=local)
    log("rank %d: image resident in %.1fs (%.2f GB, %d keys, %d values)" % (rank, time.time() - t0, idx.info.image_bytes / 1e9, idx.info.seed_keys, idx.info.seed_values))
    packed = pack_queries(seqs)
    total_bp = int(packed[1][-1])
    prm = idx.default_params()
    staged = idx.stage(packed=packed)
    launches0 = None

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    # ---- value leg: staged inputs
    for _ in range(a.warmup):
        idx.search_staged(staged, prm, collect=False)
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = int(idx.timing()[1][15])
    ms_steps, stage_ms, probe_ms, nrows, wall_ms, e2e_lib_ms, e2e_stage, kern_ms, kcnt = [], np.zeros(8), [], 0, [], [], np.zeros(8), np.zeros(16), None
    for _ in range(a.steps):
        nrows = idx.search_staged(staged, prm, collect=False)
        ms, cnt = idx.timing()
        ms_steps.append(ms[7])
        stage_ms += ms[:8]
        probe_ms.append(ms[8])
        wall_ms.append(ms[9])
        kern_ms += ms
        kcnt = cnt
    sync_all()
    launches = (int(idx.timing()[1][15]) - launches0) // max(a.steps, 1)
    # ---- e2e leg: host buffers in, rows out, every step
    e2e_ms = []
    for i in range(a.warmup + a.steps):
        t = time.perf_counter()
        nr = idx.search_count(packed, prm)
        torch.cuda.synchronize()
        if i >= a.warmup:
            e2e_ms.append((time.perf_counter() - t) * 1e3)
            e2e_lib_ms.append(idx.timing()[0][9])
            e2e_stage += idx.timing()[0][:8]
    sampler.finish()
    t_val = float(np.mean(ms_steps))
    t_e2e = float(np.mean(e2e_ms))
    hits = torch.tensor([float(nrows), float(total_bp), t_val, t_e2e], device="cuda", dtype=torch.float64)
    if dist:  # the one collective of the path: reduce the per-rank counters (NCCL over NVLink)
        tmax = hits[2:].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(hits[:2], op=dist.ReduceOp.SUM)
        hits[2:] = tmax
    rows_all, bp_all, t_val, t_e2e = (float(x) for x in hits.tolist())
    if rank != 0:
        idx.free_staged(staged)
        dist.destroy_process_group()
        return
    # ---- roofline of the seed-lookup kernel: statistics pass (untimed) on the same batch
    idx.anchors(seqs[:2000])
    _, cnt = idx.timing()
    scale = len(seqs) / 2000.0
    alg_bytes = probe_algorithmic_bytes(cnt) * scale
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_probe = float(np.mean(probe_ms)) * 1e-3
    achieved = alg_bytes / t_probe / 1e9 if t_probe > 0 else 0.0
    traffic = None   # DRAM bytes of the kernel per step from the committed `ncu --set full` capture of this workload (profiles/), if present
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if tr.get("workload_queries") == CFG["n_queries"] and tr.get("workload_genomes") == CFG["families"] * CFG["members"]:
            traffic = float(tr["dram_bytes_per_step"])   # per step, like algorithmic_bytes_per_step (a step launches the kernel once per lane)
    except Exception:
        pass
    # the same kernel alone on the GPU (one lane, so no other lane's kernels share the SMs / HBM during its launches)
    p1 = idx.default_params(lanes=1)
    iso = []
    for _ in range(3):
        idx.search_count(packed, p1)
        iso.append(idx.timing()[0][8])
    t_iso = float(np.mean(iso[1:])) * 1e-3
    # ---- CPU baseline on this box (bounded sample)
    threads = os.cpu_count() or 1
    cpu_s = float(os.environ.get("LMG_BENCH_CPU_S", 12.0))   # 0 skips the CPU leg (parameter sweeps only; the default run always reports it)
    cpu_bps, cpu_n, cpu_dt, _ = cpu_port_throughput(idx_dir, seqs, threads, target_s=cpu_s) if cpu_s > 0 else (0.0, 0, 0.0, None)
    config_lanes = int(kern_ms[12] / max(a.steps, 1) + 0.5)
    config["lanes"] = config_lanes   # concurrent sub-batches inside one call; stage_ms / kernel_ms are summed over the lanes (they overlap)
    out = {"metric": "aligned query bp/s", "value": bp_all / (t_val * 1e-3), "unit": "bp/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t_val,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config,
           "e2e": {"value": bp_all / (t_e2e * 1e-3), "unit": "bp/s", "h2d_bytes_per_step": int(packed[0].nbytes + packed[1].nbytes), "d2h_bytes_per_step": int(nrows * 136), "ms_per_step": t_e2e},
           "gpu_launches": launches, "rows_per_step": rows_all,
           "stage_ms": {k: float(v) / a.steps for k, v in zip(["h2d", "sketch", "seed_probe", "chain", "pseudo_align", "extend_wfa", "host_finish", "total"], stage_ms)},
           "roofline": {"bound": "hbm", "kernel": "k_probe_find2 (seed index lookup of the surviving probes)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": t_probe * 1e3, "launches_per_step": config_lanes,
                        "alone": {"kernel_ms": t_iso * 1e3, "achieved": alg_bytes / t_iso / 1e9 if t_iso > 0 else 0.0, "frac": (alg_bytes / t_iso / 1e9 / peak) if t_iso > 0 else 0.0, "note": "same batch through one lane: no concurrent kernels"}, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
           "cpu_baseline": {"value": cpu_bps, "unit": "bp/s", "cores": threads, "kind": "port", "sample": "%d of the %d queries, %.1fs (C++ port of the reference path; Go toolchain absent)" % (cpu_n, len(seqs), cpu_dt)},
           "debug": {"staged_call_wall_ms": float(np.mean(wall_ms)), "e2e_call_wall_ms_in_lib": float(np.mean(e2e_lib_ms)), "e2e_stage_ms": [float(x) / a.steps for x in e2e_stage],
                     "kernel_ms": {k: float(kern_ms[i]) / a.steps for k, i in [("wfa_prep+general", 10), ("wfa_fwd+bt", 11), ("extend", 13), ("pa_anchors", 14), ("pa_chain", 15)]},
                     "wfa_jobs": int(kcnt[9]), "wfa_fallback_first": int(kcnt[10]), "wfa_general_jobs": int(kcnt[11]), "probe_find_us": int(kcnt[13]), "probe_survivors": int(kcnt[1]), "probe_issued": int(kcnt[0]), "wfa_per_round": int(kcnt[14])},
           "clocks": sampler.summary()}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)
    idx.free_staged(staged)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
