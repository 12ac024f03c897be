#!/usr/bin/env python
"""bench.py -- Mbp/s classified by the `jaeger predict` hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W [--config 2|3|4]     # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W [--config C]  # the reference algorithm on the host cores
    torchrun ... bench.py --gpus N ... [--scaling weak|strong]          # N > 1: contig-sharded

Workloads (BASELINE.json `configs`):
  --config 2 (default, the one the metric is quoted on): the 1.4 M-parameter fragment architecture (declared
      stand-in, random init seed 0), synthetic 2-50 kbp contigs (i.i.d. bases, 0.1 % of contigs carry a run of N),
      --fsize 2000 --stride 1500.  One step = one pass of the hot path over `--batch-mbp` (64) Mbp of contigs per GPU;
      the 1 Gbp assembly is 16 such steps.
  --config 3: the 500 bp / 32-filter baseline model on `--fragments` (1 M) x 500 bp fragments per step (window-count-bound).
  --config 4: prophage mode, 8 synthetic 5 Mbp genomes per step: window scores + smoothing + region calling.
One *step* = pack -> window/encode -> conv stack -> heads -> per-contig aggregation (+ region calling for config 4).
Every step uses a different batch and one batch's activations are far larger than the 126 MB L2.

Prints ONE JSON line (rank 0).  `value` is timed on the device with CUDA events, the ASCII contigs already resident in
HBM.  `e2e` is the same metric through the reference-facing call -- `B200Engine.predict(WindowSource)` + the per-contig
table of `postprocess.contig_table` (+ `prophage.call_regions`) -- with HOST buffers in and the reference's result dict
(per-window logits, reliability, window metadata) back on the host: H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mbp/s classified (jaeger predict) at 1/2/4/8 B200 vs reference TF CPU"
WORKLOADS = {
    2: ("configs[1]: jaeger 1.4M-parameter fragment architecture (declared stand-in of jaeger_38341_1.4M_fragment: E128, "
        "conv k7 + 4x2 residual blocks k5 d3 C128, BN, GELU, 5 NMD taps, max pool, 6 classes + reliability head; random init "
        "seed 0), synthetic 2-50 kbp contigs, fsize 2000 stride 1500"),
    3: ("configs[2]: nn_config_500bp_baseline short-window model (E64, conv k7 C32 + 2x2 residual blocks k3 C32, BN, GELU, "
        "average pool, 3 classes; random init seed 2), synthetic 500 bp fragments, fsize 500 stride 500"),
    4: ("configs[3]: prophage mode (-p --lc 500000 -s 1.5) on synthetic 5 Mbp genomes (GC 0.5 with three 30-50 kbp islands of "
        "GC 0.35), stand-in 1.4M fragment model, fsize 2000 stride 1500: window scores + smoothing + region calling"),
}


# ---- synthetic inputs ---------------------------------------------------------------------------
def synth_lens(seed: int, target_bases: int) -> np.ndarray:
    """Contig lengths ~ U{2000..50000} until the sum reaches target_bases."""
    rng = np.random.default_rng(seed)
    lens = []
    tot = 0
    while tot < target_bases:
        n = int(rng.integers(2000, 50001))
        lens.append(n)
        tot += n
    return np.array(lens, dtype=np.int64)


def synth_bases(seed: int, lens: np.ndarray) -> np.ndarray:
    """i.i.d. uniform bases; 0.1 % of contigs get one run of 50-500 N."""
    rng = np.random.default_rng(seed)
    tot = int(lens.sum())
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=tot, dtype=np.uint8)]
    off = np.concatenate([[0], np.cumsum(lens)])
    for c in np.flatnonzero(rng.random(len(lens)) < 0.001):
        run = int(rng.integers(50, 501))
        a = int(off[c] + rng.integers(0, max(1, lens[c] - run)))
        seq[a:a + run] = ord("N")
    return seq


def synth_batch(seed: int, target_bases: int):
    lens = synth_lens(seed, target_bases)
    return synth_bases(seed + 1_000_003, lens), lens


def synth_genome(seed: int, n: int = 5_000_000) -> np.ndarray:
    """background GC 0.5 with three planted 30-50 kbp islands of GC 0.35."""
    rng = np.random.default_rng(seed)
    p = np.full(n, 0.5, dtype=np.float32)
    for _ in range(3):
        a = int(rng.integers(n // 25, n - n // 16))
        p[a:a + int(rng.integers(30_000, 50_000))] = 0.35
    gc = rng.random(n, dtype=np.float32) < p
    hi = rng.random(n, dtype=np.float32) < 0.5
    return np.where(gc, np.where(hi, ord("G"), ord("C")), np.where(hi, ord("A"), ord("T"))).astype(np.uint8)


class Workload:
    """What one step of a config processes, identically described for both arms."""

    def __init__(self, args):
        from jaeger_b200.modelspec import baseline_500bp_config, parse_project, standin_1p4m_config   # pure host code, no .so
        self.cfg = args.config
        self.fsize, self.stride = (500, 500) if self.cfg == 3 else (2000, 1500)
        self.spec = parse_project(baseline_500bp_config() if self.cfg == 3 else standin_1p4m_config())
        self.weight_seed = 2 if self.cfg == 3 else 0
        self.prophage = self.cfg == 4
        if self.cfg == 2:
            self.bases_per_gpu = int(args.batch_mbp * 1e6)
        elif self.cfg == 3:
            self.bases_per_gpu = int(args.fragments) * 500
        else:
            self.bases_per_gpu = int(args.genomes) * 5_000_000
        self.args = args

    def global_lens(self, step: int, world: int, strong: bool) -> np.ndarray:
        """The contig list of a step: identical on every rank, sharded afterwards."""
        total = self.bases_per_gpu * (1 if strong else world)
        if self.cfg == 2:
            return synth_lens(step + 1, total)
        if self.cfg == 3:
            return np.full(total // 500, 500, dtype=np.int64)
        return np.full(total // 5_000_000, 5_000_000, dtype=np.int64)

    def bases(self, step: int, rank: int, lens: np.ndarray, ids: np.ndarray) -> np.ndarray:
        if self.cfg == 2:
            return synth_bases(7919 * (step + 1) + rank, lens)
        if self.cfg == 3:
            rng = np.random.default_rng(2 + 31 * step + rank)
            return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(lens.sum()), dtype=np.uint8)]
        return np.concatenate([synth_genome(3 + 1000 * step + int(g)) for g in ids]) if len(ids) else np.zeros(0, np.uint8)

    def config_json(self, world: int, strong: bool) -> dict:
        unit = {2: f"{self.args.batch_mbp:g} Mbp of contigs", 3: f"{int(self.args.fragments)} fragments of 500 bp",
                4: f"{int(self.args.genomes)} genomes of 5 Mbp"}[self.cfg]
        return {"workload": WORKLOADS[self.cfg], "config_id": self.cfg,
                "step": f"one pass of the hot path over {unit} per {'job' if strong else 'GPU'}",
                "l2_policy": "inputs larger than L2: a different batch every step, activations per step >> 126 MB L2",
                "parallelism": "length-balanced contig sharding (LPT on window counts), per-contig records gathered on rank 0 once, after the last step of a leg"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz, self._halt = device, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis else device
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_min_mhz": float(min(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "hbm_gbs": float(d.get("hbm_gbs", 6650.0)),
                "source": "measured (MEASURED_PEAKS.json: bf16_tflops_sustained, hbm_gbs)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained, 6.65 TB/s)"}


# ---- the reference algorithm on the host (oracle port; bench's CPU legs only) ---------------------
def cpu_reference_step(wl: Workload, weights, seq: np.ndarray, lens: np.ndarray) -> tuple[int, int]:
    """window -> encode -> forward -> aggregate (+ smoothing / region calling for config 4) with the oracle."""
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import postprocess as opp
    from oracle import seqwin
    off = np.concatenate([[0], np.cumsum(lens)])
    recs = [(f"c{i}", seq[off[i]:off[i + 1]].tobytes().decode()) for i in range(len(lens))]
    wins = list(seqwin.fragment_windows(recs, wl.fsize, wl.stride))
    if not wins:
        return 0, 0
    tok = oenc.encode_windows([w.seq for w in wins], wl.fsize)
    outs = []
    for b in range(0, len(wins), 96):                       # reference default --batch 96
        outs.append(ofw.forward(wl.spec, weights, tok[b:b + 96]))
    y = {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
    y["meta_2"] = np.array([w.is_last for w in wins])
    opp.aggregate_numeric(y["prediction"], y.get("reliability"), y["meta_2"])
    if wl.prophage:
        from oracle import prophage as opro
        ends = np.flatnonzero(y["meta_2"]) + 1
        for a, b in zip(np.concatenate([[0], ends[:-1]]), ends):
            opro.segment(opro.smooth_scores(y["prediction"][a:b])[:, 1], 1.5)
    return int(lens.sum()), len(wins)


def cpu_sample(wl: Workload, seed: int, sample_bases: int):
    """A bounded prefix-sized sample of the workload's contig stream."""
    if wl.cfg == 2:
        return synth_batch(seed, sample_bases)
    if wl.cfg == 3:
        n = max(1, sample_bases // 500)
        rng = np.random.default_rng(seed)
        return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n * 500, dtype=np.uint8)], np.full(n, 500, np.int64)
    n = max(wl.fsize * 8, sample_bases)           # one genome fragment, long enough for the change-point search to have work
    return synth_genome(seed, n), np.array([n], dtype=np.int64)


def cpu_threads() -> int:
    import torch
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1: use every host core
    return torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm for the path, timed on the host cores.  TensorFlow (the reference's
    own engine) cannot be installed offline, so the arm times the oracle port (kind "port"); nothing of jaeger_b200's native
    library is imported or loaded here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from jaeger_b200.modelspec import init_random
    wl = Workload(args)
    weights = init_random(wl.spec, wl.weight_seed)
    cores = cpu_threads()
    sample = int(args.ref_sample_kbp * 1000)
    times, bases, wins = [], 0, 0
    for i in range(args.warmup + args.steps):
        seq, lens = cpu_sample(wl, 1000 + i, sample)
        t0 = time.perf_counter()
        b, w = cpu_reference_step(wl, weights, seq, lens)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt); bases += b; wins += w
    total = sum(times)
    v = bases / 1e6 / total
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": wl.config_json(world, args.scaling == "strong"),
            "windows_per_s": wins / total,
            "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {args.ref_sample_kbp:g} kbp of the workload's contig stream per step; oracle "
                                       "port of the reference path (NumPy windowing / encoding / aggregation, torch fp32 conv stack "
                                       "on all host threads); TensorFlow is not installable offline"},
            "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl: Workload, weights, budget_s: float):
    """Oracle port of the reference path timed on the host cores over a bounded sample (rank 0, N = 1 only)."""
    cores = cpu_threads()
    bases, total, n = 0, 0.0, 0
    while (total < budget_s and n < 50) or n == 0:
        seq, lens = cpu_sample(wl, 4242 + n, 60_000)
        t0 = time.perf_counter()
        b, _ = cpu_reference_step(wl, weights, seq, lens)
        total += time.perf_counter() - t0
        bases += b; n += 1
    return {"value": bases / 1e6 / total, "unit": "Mbp/s", "cores": cores, "kind": "port",
            "sample": f"{n} x 60 kbp of the same synthetic contig stream ({bases} bp, {total:.1f} s): oracle port "
                      "(NumPy windowing/encoding, torch fp32 conv stack on all host threads, NumPy aggregation)"}


# ---- this repo's CUDA path ------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from jaeger_b200 import B200Engine, WindowSource
    from jaeger_b200.modelspec import init_random
    from jaeger_b200.parallel import gather_contig_records, shard_contigs
    from jaeger_b200.postprocess import contig_table
    from jaeger_b200 import prophage as ppro

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = Workload(args)
    strong = args.scaling == "strong"
    weights = init_random(wl.spec, wl.weight_seed)
    eng = B200Engine(spec=wl.spec, weights=weights, device=local, workspace_gb=args.workspace_gb)
    dev = eng.tdev
    stream = eng._stream()
    n_batches = args.warmup + args.steps
    fsize, stride = wl.fsize, wl.stride
    # Length-balanced contig sharding: the step's contig list is identical on every rank and is
    # bin-packed on window counts; a rank materialises only its own shard.
    host_batches, shards_per_step, n_global = [], [], []
    for i in range(n_batches):
        glens = wl.global_lens(i, world, strong)
        shards = shard_contigs(glens, world, fsize, stride)
        mine = shards[rank]
        host_batches.append((wl.bases(i, rank, glens[mine], mine), glens[mine]))
        shards_per_step.append(shards)
        n_global.append(len(glens))
    pinned = [torch.from_numpy(s).pin_memory() for s, _ in host_batches]
    names = [[f"c{j}" for j in range(len(l))] for _, l in host_batches]
    offsets = [np.concatenate([[0], np.cumsum(l)]).astype(np.int64) for _, l in host_batches]
    phage_k = [c.lower() for c in eng.class_map["class"]].index("phage") if wl.prophage else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    pending = []       # per-contig records of the steps of the current leg, gathered on rank 0 at the END of the leg

    def gather_results(pred_sum, consensus, i):
        """NCCL is used only to deliver the per-contig results to rank 0.  The records of a step are kept on the
        device and the pre-sized gathers (parallel.gather_contig_records: no size exchange, no host synchronisation) run after the
        leg's last step, inside the timed region: a gather per step would make every rank wait for the slowest one every step, and
        the driver writes its tables at the end of a run anyway."""
        if world == 1:
            return
        rec = torch.cat([pred_sum.float(), consensus.float().unsqueeze(1)], dim=1)
        if rec.shape[0] != len(shards_per_step[i][rank]):         # contigs without windows report zero rows: pad by position
            full = torch.zeros((len(shards_per_step[i][rank]), rec.shape[1]), device=dev)
            full[:rec.shape[0]] = rec
            rec = full
        pending.append((rec, i))

    def flush_gathers():
        for rec, i in pending:
            gather_contig_records(rec, shards_per_step[i], n_global[i], dst=0)
        pending.clear()

    def device_step(i, x):
        agg, w, c = eng.classify_long(x, host_batches[i][1], fsize, stride)
        if wl.prophage and agg:
            ppro.segment_device(eng, agg["_logits"], agg["_window_offsets"], phage_k)
        if agg:
            gather_results(agg["pred_sum"], agg["consensus"], i)
        return w, c

    def e2e_step(i):
        """The reference-facing call: host buffers in, the reference's result dict + per-contig table out (on the host)."""
        src = WindowSource.from_host(names[i], pinned[i], offsets[i], fsize=fsize, stride=stride,
                                     outputs=("prediction", "reliability"), lazy_meta=True)
        y = eng.predict(src)
        data = contig_table(eng, y, fsize)
        if wl.prophage:
            ppro.call_regions(eng, data, eng.class_map, fsize, stride, lc=500_000, sensitivity=1.5)
        if world > 1:
            with torch.cuda.stream(stream):
                gather_results(torch.from_numpy(data["pred_sum"].astype(np.float32)).to(dev),
                               torch.from_numpy(data["consensus"].astype(np.float32)).to(dev), i)
        d2h = sum(int(np.asarray(v).nbytes) for k, v in y.items() if not k.startswith("meta_"))
        t = eng.windows
        d2h += int(t.counts.nbytes + t.skew100.nbytes)
        d2h += sum(int(np.asarray(data[k]).nbytes) for k in ("pred_sum", "pred_var", "consensus", "per_class_counts", "entropy", "energy", "frag_pred"))
        h2d = int(pinned[i].numel()) + int(y["prediction"].shape[0]) * 12 + int(len(offsets[i])) * 8
        return int(y["prediction"].shape[0]), h2d, d2h

    with torch.cuda.stream(stream):
        # ---- leg 1: device-resident inputs, CUDA-event timing -------------------------------------
        dev_batches = [p.to(dev) for p in pinned]
        for i in range(args.warmup):
            device_step(i, dev_batches[i])
        flush_gathers()
        barrier()
        eng.set_profiling(True)
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = eng.ctx.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        n_bases = n_windows = n_contigs = 0
        for i in range(args.warmup, n_batches):
            w, c = device_step(i, dev_batches[i])
            n_bases += int(host_batches[i][1].sum()); n_windows += w; n_contigs += c
        ev_own = torch.cuda.Event(enable_timing=True)
        ev_own.record(stream)                   # this rank's own work, before it meets the others in the gathers
        flush_gathers()
        ev1.record(stream)
        barrier()
        own_ms = ev0.elapsed_time(ev_own)
        ms = ev0.elapsed_time(ev1)
        launches = eng.ctx.launch_count - launches0
        clocks = sampler.stop()
        prof = eng.get_profile()
        eng.set_profiling(False)
        del dev_batches
    # ---- leg 2: end to end through the engine's public call with host buffers ---------------------
    e2e_step(0)                                 # one untimed end-to-end step: no first-use cost in the first timed one
    if world > 1:
        with torch.cuda.stream(stream):
            flush_gathers()
    barrier()
    t0 = time.perf_counter()
    h2d_bytes = d2h_bytes = 0
    for i in range(args.warmup, n_batches):
        _, h2d_bytes, d2h_bytes = e2e_step(i)
    if world > 1:
        with torch.cuda.stream(stream):
            flush_gathers()
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([ms, e2e_s * 1e3, float(n_bases), float(n_windows), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        tot_bases, tot_windows, launches = float(tsum[2]), float(tsum[3]), int(tsum[4])
        own = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(own, torch.tensor([own_ms], device=dev, dtype=torch.float64))
        rank_ms = [float(x[0]) / args.steps for x in own]
    else:
        e2e_ms, tot_bases, tot_windows = e2e_s * 1e3, float(n_bases), float(n_windows)

    line = None
    if rank == 0:
        value = tot_bases / 1e6 / (ms / 1e3)
        line = {"metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f16", "data": "synthetic", "config": wl.config_json(world, strong),
                "windows_per_s": tot_windows / (ms / 1e3),
                "roofline": roofline(eng, wl, prof, ms, tot_windows, world),
                "e2e": {"value": tot_bases / 1e6 / (e2e_ms / 1e3), "unit": "Mbp/s", "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(d2h_bytes),
                        "call": "B200Engine.predict(WindowSource.from_host(pinned host buffer)) -> reference result dict on the host, "
                                "then postprocess.contig_table" + (" + prophage.call_regions" if wl.prophage else "")},
                "gpu_launches": int(launches), "clocks": clocks}
        if world > 1:        # every rank's own device time per step before the final gathers: the spread is the GPUs' (power cap), not the path's
            line["rank_ms_per_step"] = [round(v, 2) for v in rank_ms]
        agree = ROOT / "profiles" / "label_agreement_r2.json"
        if agree.exists():
            try:
                line["label_agreement"] = json.loads(agree.read_text())
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, weights, args.cpu_baseline_seconds)
    # orderly shutdown: engine (CUDA library handles) first, then the process group, then a normal interpreter exit
    eng.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def roofline(eng, wl: Workload, prof, ms: float, tot_windows: float, world: int) -> dict:
    """Roofline of the dominant kernel family, from the per-launch CUDA events the library records on its own stream."""
    plan = eng.plan
    peaks = measured_peaks()
    lc = eng.codons_per_frame(wl.fsize, wl.fsize)
    conv_ms_total = sum(p[0] for p in prof)
    if "stack_resident_kernel" in eng.conv_kernel_names():
        # ONE launch runs the whole conv stack with every window resident in shared memory (csrc/conv_resident.cuh): its HBM traffic
        # is the tokens in and 32 floats out, so the kernel is bounded by the SM (tensor pipe at N = 32 + epilogue issue), not HBM.
        ms_k = sum(p[0] for p in prof)
        n_launch = max(1.0, sum(p[1] for p in prof))
        windows = prof[-1][2]
        flop = 0.0
        for c in plan.launches:
            if c.kind != 1:
                continue
            k = c.kernel.shape[0]
            l_out = ((lc - c.cum_shrink_in) >> c.halvings) - c.shrink
            flop += 2.0 * 6 * l_out * k * c.real_cin * c.real_cout
        tflops = flop * windows / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
        byts = (6.0 * ((lc + 3) // 4 * 4) + 4.0 * plan.real_feat_dim) * windows
        traffic = None
        summ = ROOT / "profiles" / f"ncu_summary_r2_config{wl.cfg}.json"
        if summ.exists():
            try:
                per_window = json.loads(summ.read_text()).get("dram_bytes_per_window")
                traffic = per_window * windows / n_launch if per_window else None
            except Exception:
                traffic = None
        return {"bound": "tensor", "achieved": tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tflops / peaks["tflops"],
                "kernel": eng.conv_kernel_names(), "traffic": traffic, "peak_source": peaks["source"],
                "avg_launch_ms": ms_k / n_launch, "kernel_share_of_step": conv_ms_total / ms,
                "algorithmic_flop_per_launch": flop * windows / n_launch, "algorithmic_bytes_per_launch": byts / n_launch,
                "hbm_gbs_algorithmic": byts / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0, "hbm_peak_gbs": peaks["hbm_gbs"],
                "windows_per_launch": windows / n_launch,
                "note": "window-resident kernel: activations never leave the SM, so neither roofline binds; the SS-mode N = 32 MMAs are "
                        "shared-memory-bandwidth limited and the epilogue warps are issue limited (DESIGN.md section 4)"}
    res_ms = res_launch = res_flop = res_bytes = res_windows = 0.0
    for li, (c, (pms, pl, pw)) in enumerate(zip(plan.launches, prof)):
        if li == 0 or c.kind != 1:
            continue                                    # the dominant family: the residual-block convolutions (not the stem)
        k, cin, cout = c.kernel.shape
        l_out = ((lc - c.cum_shrink_in) >> c.halvings) - c.shrink
        res_ms += pms; res_launch += pl; res_windows += pw
        res_flop += 2.0 * 6 * l_out * k * c.real_cin * c.real_cout * pw
        # per-layer streaming minimum at the layer's REAL channel count: read x, write y (+ read the shortcut), fp16
        res_bytes += 6.0 * l_out * 2.0 * (c.real_cin + (c.real_cout if c.out_buf >= 0 else 0) + (c.real_cout if c.sc_buf >= 0 else 0)) * pw
    avg_launch_s = res_ms / max(1.0, res_launch) * 1e-3
    traffic = None
    summ = ROOT / "profiles" / f"ncu_summary_r2_config{wl.cfg}.json"
    if not summ.exists():
        summ = ROOT / "profiles" / "ncu_summary_r1.json" if wl.cfg == 2 else summ
    if summ.exists():
        try:
            per_window = json.loads(summ.read_text()).get("conv_tc_dram_bytes_per_window_mean")
            traffic = per_window * (res_windows / max(1.0, res_launch)) if per_window else None
        except Exception:
            traffic = None
    common = {"kernel": eng.conv_kernel_names(), "traffic": traffic, "peak_source": peaks["source"],
              "avg_launch_ms": res_ms / max(1.0, res_launch), "kernel_share_of_step": conv_ms_total / ms,
              "algorithmic_flop_per_launch": res_flop / max(1.0, res_launch),
              "algorithmic_bytes_per_launch": res_bytes / max(1.0, res_launch)}
    tflops = res_flop / (res_ms * 1e-3) / 1e12 if res_ms > 0 else 0.0
    gbs = res_bytes / (res_ms * 1e-3) / 1e9 if res_ms > 0 else 0.0
    if wl.cfg == 3:      # 32-channel layers: bound by HBM, reported against the measured copy bandwidth
        return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                "tensor_tflops": tflops, **common}
    return {"bound": "tensor", "achieved": tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tflops / peaks["tflops"],
            "hbm_gbs_algorithmic": gbs, "hbm_peak_gbs": peaks["hbm_gbs"],
            "hbm_gbs_at_measured_traffic": (traffic / avg_launch_s / 1e9) if traffic and avg_launch_s > 0 else None,
            "whole_model_tflops": plan.flops_per_window(lc, algorithmic_stem_cin=wl.spec.embedding_size) * tot_windows / world / (ms * 1e-3) / 1e12,
            **common}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4], help="BASELINE.json config (1-based): 2 = the metric's config")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: a fixed batch per GPU per step; strong: one fixed global batch per step sharded over the GPUs")
    ap.add_argument("--batch-mbp", type=float, default=64.0, help="config 2: Mbp of contigs per GPU (weak) or per job (strong) per step")
    ap.add_argument("--fragments", type=float, default=1e6, help="config 3: 500 bp fragments per step")
    ap.add_argument("--genomes", type=int, default=8, help="config 4: 5 Mbp genomes per step")
    ap.add_argument("--workspace-gb", type=float, default=24.0)
    ap.add_argument("--ref-sample-kbp", type=float, default=60.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
