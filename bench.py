#!/usr/bin/env python
"""bench.py -- Mbp/s classified by the `jaeger predict` hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores
    torchrun ... bench.py --gpus N ...                        # N > 1: contig-sharded, weak scaling

Workload = BASELINE.json configs[1]: the 1.4 M-parameter fragment architecture (declared
stand-in, random init, seed 0), synthetic 2-50 kbp contigs (i.i.d. bases, 0.1 % of contigs carry a
run of N), --fsize 2000 --stride 1500.  One *step* = one pass of the hot path
(pack -> window/encode -> conv stack -> heads -> per-contig aggregation) over one batch of
`--batch-mbp` (default 64) Mbp of contigs per GPU; the 1 Gbp assembly is 16 such steps.  Every
step uses a different batch, and one batch's activations (~2.6 MB per window, ~40 k windows) are
far larger than the 126 MB L2, so no step finds its inputs cached.

Prints ONE JSON line (rank 0).  `value` is timed on the device with CUDA events with the ASCII
contigs already resident in HBM; `e2e` is the same metric through `B200Engine` with pinned HOST
buffers, H2D of the contigs and D2H of the per-contig results inside the timed region.
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
FSIZE, STRIDE = 2000, 1500
WORKLOAD = ("configs[1]: jaeger 1.4M-parameter fragment architecture (declared stand-in of "
            "jaeger_38341_1.4M_fragment: E128, conv k7 + 4x2 residual blocks k5 d3 C128, BN, GELU, 5 NMD taps, "
            "max pool, 6 classes + reliability head; random init seed 0), synthetic 2-50 kbp contigs, "
            "fsize 2000 stride 1500")


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


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


# ---------------------------------------------------------------------------------------------
def cpu_reference_step(spec, weights, seq: np.ndarray, lens: np.ndarray) -> tuple[int, int]:
    """The reference algorithm on the host: window -> encode -> forward -> aggregate (oracle)."""
    from oracle import encode as oenc
    from oracle import forward as ofw
    from oracle import postprocess as opp
    from oracle import seqwin
    off = np.concatenate([[0], np.cumsum(lens)])
    recs = [(f"c{i}", seq[off[i]:off[i + 1]].tobytes().decode()) for i in range(len(lens))]
    wins = list(seqwin.fragment_windows(recs, FSIZE, STRIDE))
    if not wins:
        return 0, 0
    tok = oenc.encode_windows([w.seq for w in wins], FSIZE)
    outs = []
    for b in range(0, len(wins), 96):                       # reference default --batch 96
        outs.append(ofw.forward(spec, weights, tok[b:b + 96]))
    y = {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
    y["meta_2"] = np.array([w.is_last for w in wins])
    opp.aggregate_numeric(y["prediction"], y.get("reliability"), y["meta_2"])
    return int(lens.sum()), len(wins)


def run_reference(args):
    import torch
    from jaeger_b200.modelspec import init_random, parse_project, standin_1p4m_config
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = parse_project(standin_1p4m_config())
    weights = init_random(spec, 0)
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1: use every host core
    cores = torch.get_num_threads()
    sample = int(args.ref_sample_kbp * 1000)
    times, bases, wins = [], 0, 0
    for i in range(args.warmup + args.steps):
        seq, lens = synth_batch(1000 + i, sample)
        t0 = time.perf_counter()
        b, w = cpu_reference_step(spec, weights, seq, lens)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt); bases += b; wins += w
    total = sum(times)
    v = bases / 1e6 / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{args.ref_sample_kbp} kbp of contigs per step"},
            "windows_per_s": wins / total,
            "cpu_baseline": {"value": v, "unit": "Mbp/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x {args.ref_sample_kbp} kbp of the same contig stream; "
                                       "oracle port (NumPy windowing/encoding + torch fp32 conv stack) of the reference path; "
                                       "TensorFlow is not installable offline"},
            "e2e": {"value": v, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from jaeger_b200 import B200Engine, parse_project, standin_1p4m_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "") and not os.environ.get("JG_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = parse_project(standin_1p4m_config())
    eng = B200Engine(spec=spec, device=local, seed=0, workspace_gb=args.workspace_gb)
    dev = eng.tdev
    stream = eng._stream()
    target = int(args.batch_mbp * 1e6)
    n_batches = args.warmup + args.steps
    # Length-balanced contig sharding: the step's contig list (world x batch Mbp,
    # identical on every rank) is bin-packed on window counts; a rank materialises only its shard.
    from jaeger_b200.parallel import gather_contig_records, shard_contigs
    host_batches, shard_ids, n_global = [], [], []
    for i in range(n_batches):
        glens = synth_lens(i + 1, world * target)
        mine = shard_contigs(glens, world, FSIZE, STRIDE)[rank]
        host_batches.append((synth_bases(7919 * (i + 1) + rank, glens[mine]), glens[mine]))
        shard_ids.append(torch.from_numpy(mine))
        n_global.append(len(glens))
    pinned = [torch.from_numpy(s).pin_memory() for s, _ in host_batches]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather_results(agg, i):
        """NCCL is used only to gather the per-contig results on rank 0."""
        if world == 1 or not agg:
            return
        rec = torch.cat([agg["pred_sum"].float(), agg["consensus"].float().unsqueeze(1)], dim=1)
        gather_contig_records(rec, shard_ids[i].to(dev), n_global[i], dst=0)

    results = {}
    with torch.cuda.stream(stream):
        # ---- leg 1: device-resident inputs, CUDA-event timing ---------------------------------
        dev_batches = [p.to(dev) for p in pinned]
        for i in range(args.warmup):
            agg, _, _ = eng.classify_long(dev_batches[i], host_batches[i][1], FSIZE, STRIDE)
            gather_results(agg, i)
        barrier()
        eng.set_profiling(True)
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = eng.ctx.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        n_bases = n_windows = n_contigs = 0
        for i in range(args.warmup, n_batches):
            agg, w, c = eng.classify_long(dev_batches[i], host_batches[i][1], FSIZE, STRIDE)
            gather_results(agg, i)
            n_bases += int(host_batches[i][1].sum()); n_windows += w; n_contigs += c
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng.ctx.launch_count - launches0
        clocks = sampler.stop()
        prof = eng.get_profile()
        eng.set_profiling(False)
        del dev_batches
        # ---- leg 2: end to end through the engine API with host buffers ------------------------
        d2h_bytes = 0
        # one untimed end-to-end step (pinned H2D -> engine -> D2H) so the first timed step pays no first-use cost
        xw = pinned[0].to(dev, non_blocking=True)
        aggw, _, _ = eng.classify_long(xw, host_batches[0][1], FSIZE, STRIDE)
        gather_results(aggw, 0)
        _ = {k: aggw[k].cpu() for k in ("pred_sum", "consensus")} if aggw else None
        del xw, aggw
        barrier()
        sampler2 = ClockSampler(local) if os.environ.get("JG_BENCH_E2E_CLOCKS") else None   # NVML polling perturbs the synchronous e2e steps
        if sampler2:
            sampler2.start()
        if os.environ.get("JG_BENCH_DEBUG"):
            eng.set_profiling(True)
        t0 = time.perf_counter()
        e2e_ev = []
        for i in range(args.warmup, n_batches):
            if os.environ.get("JG_BENCH_DEBUG"):
                e2e_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), time.perf_counter()))
                e2e_ev[-1][0].record(stream)
            tp0 = time.perf_counter()
            x = pinned[i].to(dev, non_blocking=True)
            agg, w, c = eng.classify_long(x, host_batches[i][1], FSIZE, STRIDE)
            tp1 = time.perf_counter()
            gather_results(agg, i)
            host = {k: agg[k].cpu() for k in ("pred_sum", "pred_var", "consensus", "per_class_counts", "entropy", "energy", "rel_pos")}
            if e2e_ev:
                conv_ms = sum(p_[0] for p_ in eng.get_profile())
                sys.stderr.write(f"[e2e phases] launch calls {1e3 * (tp1 - tp0):.1f} ms, d2h + wait {1e3 * (time.perf_counter() - tp1):.1f} ms, conv kernels {conv_ms:.1f} ms\n")
                eng.set_profiling(True)
            d2h_bytes = sum(v.numel() * v.element_size() for v in host.values())
            if e2e_ev:
                e2e_ev[-1][1].record(stream)
                e2e_ev[-1] = e2e_ev[-1] + (time.perf_counter(),)
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks_e2e = sampler2.stop() if sampler2 else {}
        for a, b, h0, h1 in e2e_ev:
            sys.stderr.write(f"[e2e step] device {a.elapsed_time(b):.1f} ms  host {1e3 * (h1 - h0):.1f} ms\n")
    t = torch.tensor([ms, e2e_s * 1e3, float(n_bases), float(n_windows), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        tot_bases, tot_windows, launches = float(tsum[2]), float(tsum[3]), int(tsum[4])
    else:
        e2e_ms, tot_bases, tot_windows = e2e_s * 1e3, float(n_bases), float(n_windows)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)
    value = tot_bases / 1e6 / (ms / 1e3)
    # ---- roofline of the dominant kernel: the k5 d3 residual conv launches ---------------------
    plan = eng.plan
    peak, peak_src = measured_peaks()
    lc = eng.codons_per_frame(FSIZE, FSIZE)
    res_ms = res_launch = 0.0
    res_flop = 0.0
    conv_ms_total = sum(p[0] for p in prof)
    for li, (c, (pms, pl, pw)) in enumerate(zip(plan.launches, prof)):
        if li == 0:
            continue
        k, cin, cout = c.kernel.shape
        res_ms += pms; res_launch += pl
        res_flop += 2.0 * 6 * (lc - c.cum_shrink_in - c.shrink) * k * cin * cout * pw
    achieved = res_flop / (res_ms * 1e-3) / 1e12 if res_ms > 0 else 0.0
    traffic = None
    summ = ROOT / "profiles" / "ncu_summary_r1.json"
    if summ.exists():
        try:
            per_window = json.loads(summ.read_text()).get("conv_tc_dram_bytes_per_window_mean")
            # the ncu capture is per window of one launch; scale to this run's average launch
            traffic = per_window * (sum(p[2] for p in prof[1:]) / max(1.0, res_launch)) if per_window else None
        except Exception:
            traffic = None
    hbm_peak = None
    try:
        hbm_peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs"))
    except Exception:
        pass
    avg_launch_s = res_ms / max(1.0, res_launch) * 1e-3
    roofline = {"bound": "tensor",
                "kernel": "the 16 k5 d3 C128 residual-conv launches of a forward pass: jg::tc2::conv_tc2_kernel<2> (CTA pair, 12 launches) + "
                          "jg::tc2::conv_tc2_kernel<3> (CTA pair with 3 epilogue groups, the 4 launches with NMD tap + second affine)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "avg_launch_ms": res_ms / max(1.0, res_launch),
                "hbm_gbs_at_measured_traffic": (traffic / avg_launch_s / 1e9) if traffic and avg_launch_s > 0 else None,
                "hbm_peak_gbs": hbm_peak,
                "algorithmic_flop_per_window_per_launch": 2.0 * 6 * 659 * 5 * 128 * 128,
                "kernel_share_of_step": conv_ms_total / ms,
                "whole_model_tflops": plan.flops_per_window(lc, algorithmic_stem_cin=spec.embedding_size) * tot_windows / world / (ms * 1e-3) / 1e12}
    line = {"metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_mbp_per_gpu_per_step": args.batch_mbp,
                       "l2_policy": "inputs larger than L2: a different 64 Mbp batch every step, ~100 GB of activations per step",
                       "parallelism": f"length-balanced contig sharding x{world} (LPT on window counts), NCCL gather of per-contig records only"},
            "windows_per_s": tot_windows / (ms / 1e3),
            "roofline": roofline,
            "e2e": {"value": tot_bases / 1e6 / (e2e_ms / 1e3), "unit": "Mbp/s",
                    "h2d_bytes_per_step": int(target + n_windows / args.steps * 12), "d2h_bytes_per_step": int(d2h_bytes),
                    **({"sm_mhz": clocks_e2e.get("sm_mhz"), "sm_min_mhz": clocks_e2e.get("sm_min_mhz")} if clocks_e2e else {})},
            "gpu_launches": int(launches), "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(eng.spec, eng.weights, args.cpu_baseline_seconds)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)      # skip interpreter-exit destructors that race the CUDA context teardown


def cpu_baseline(spec, weights, budget_s: float):
    """Oracle port of the reference path timed on the host cores over a bounded sample."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    seq, lens = synth_batch(4242, 60_000)
    t0 = time.perf_counter()
    b, w = cpu_reference_step(spec, weights, seq, lens)
    dt = time.perf_counter() - t0
    bases, total, n = b, dt, 1
    while total < budget_s and n < 50:
        seq, lens = synth_batch(4242 + n, 60_000)
        t0 = time.perf_counter()
        b, w = cpu_reference_step(spec, weights, seq, lens)
        total += time.perf_counter() - t0
        bases += b; n += 1
    return {"value": bases / 1e6 / total, "unit": "Mbp/s", "cores": cores, "kind": "port",
            "sample": f"{n} x 60 kbp of the same synthetic contig stream ({bases} bp, {total:.1f} s): oracle port "
                      "(NumPy windowing/encoding, torch fp32 conv stack on all host threads, NumPy aggregation)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-mbp", type=float, default=64.0)
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
