#!/usr/bin/env python
"""Round-2 profile summaries from the ncu artefacts `tools/profile_r2.sh` brings back in gpurun_out/:

    python tools/summarise_ncu_r2.py

writes profiles/launches_r2.csv (the launch list, trimmed to kernel / grid / block / ns), profiles/ncu_summary_r2_config2.json
(per-kernel-mode metrics of the `ncu --set full` capture; `bench.py` reads `conv_tc_dram_bytes_per_window_mean` from it for
`roofline.traffic`) and profiles/sass_histogram_r2.json (tcgen05 / TMA / matrix opcodes in the built library).
"""
from __future__ import annotations

import collections
import csv
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
ROWS_PER_WINDOW = 4096          # 6 frames x 672-row period rounded to 256 (csrc/jaeger_b200.cu model_geometry), fsize 2000
LC = 666
MODES = {0: "light (conv1 of a block)", 1: "light + shortcut (conv2 of a non-final block)",
         2: "final (conv2 + block-end taps)", 3: "final + max pool (last block)"}


def launch_list() -> dict:
    rows = list(csv.DictReader(l for l in open(OUT / "launches_r2.csv") if l.startswith('"')))
    with open(PROF / "launches_r2.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "duration_ns"])
        for r in rows:
            w.writerow([r["ID"], r["Kernel Name"].split("(")[0].replace("void ", ""), r["Grid Size"], r["Block Size"], r["Metric Value"]])
    agg: dict[str, list] = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    return {"launches": len(rows), "gpu_ms_total": tot,
            "by_kernel": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot, 4)}
                          for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}}


def full_capture() -> dict:
    raw = subprocess.run(["ncu", "-i", str(OUT / "ncu_ws_r2.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=1.0):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        u = units[i]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
        return v * mult * scale

    per_mode: dict[str, list] = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"conv_ws_kernel<(\d), (\d), (\d)>", name)
        if not m:
            continue
        mode, taps = int(m.group(1)), int(m.group(2))
        key = "stem_k7" if taps == 7 else f"mode{mode}"
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        per_mode[key].append({
            "kernel": name.split("(")[0].replace("void ", ""),
            "duration_ms": val(r, "gpu__time_duration.sum"),
            "dram_read_bytes": rd, "dram_write_bytes": wr,
            "windows": round(wr / (ROWS_PER_WINDOW * 256.0)),
            "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_realtime_pct": val(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers_per_thread": val(r, "launch__registers_per_thread"),
            "block": val(r, "launch__block_size"), "grid": val(r, "launch__grid_size"),
            "dynamic_smem_bytes": val(r, "launch__shared_mem_per_block_dynamic"),
        })
    out: dict = {}
    for key, caps in per_mode.items():
        n = len(caps)
        mean = {k: (sum(c[k] for c in caps) / n if isinstance(caps[0][k], float) else caps[0][k]) for k in caps[0]}
        mean["captures"] = n
        if key != "stem_k7":
            mean["what"] = MODES[int(key[-1])]
        out[key] = mean
    return out


def sass_histogram() -> dict:
    sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "jaeger_b200" / "libjaeger_b200.so")], capture_output=True, text=True).stdout
    cur, hist = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "STSM", "LDSM", "SYNCS", "MUFU", "HFMA2", "FFMA"]
    out = {}
    for f, c in hist.items():
        if "conv_ws" not in f and "conv_tc" not in f:
            continue
        name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        out[name] = {"instructions": sum(c.values()), **{k: c[k] for k in keys if c[k]}}
    return out


def main() -> None:
    ll = launch_list()
    cap = full_capture()
    per_window = {}
    for key, c in cap.items():
        per_window[key] = (c["dram_read_bytes"] + c["dram_write_bytes"]) / c["windows"]
    alg = {"mode0": 6 * LC * 2 * 256, "mode1": 6 * LC * 2 * 384, "mode2": 6 * LC * 2 * 384, "mode3": 6 * LC * 2 * 384}
    # the 16 residual launches of the 1.4M graph: 8 x mode0, 4 x mode1, 3 x mode2, 1 x mode3 (mode3 writes no activation; counted as mode2 here
    # when it was not among the 8 captured launches)
    m3 = per_window.get("mode3", per_window["mode2"])
    mean = (8 * per_window["mode0"] + 4 * per_window["mode1"] + 3 * per_window["mode2"] + m3) / 16
    summary = {
        "source": "tools/profile_r2.sh under gpurun (1 x B200): ncu --set full --clock-control none --import-source on -k regex:conv_ws -s 34 -c 8 "
                  "(python bench.py --steps 1 --warmup 1 --no-cpu-baseline); round 2, weights-stationary conv kernel (csrc/conv_ws.cuh)",
        "rows_per_window": ROWS_PER_WINDOW, "codons_per_frame": LC,
        "launch_list": ll,
        "conv_ws": cap,
        "conv_tc_dram_bytes_per_window": per_window,
        "algorithmic_bytes_per_window": alg,
        "conv_tc_dram_bytes_per_window_mean": mean,
        "note": "per-launch ncu durations are cold-cache and serialised; the bench's CUDA-event times are the ones the roofline uses. "
                "DRAM bytes per window exceed the algorithmic bytes by the frame padding (672-row period for 666 codons, 4096 rows for 4032) only.",
    }
    (PROF / "ncu_summary_r2_config2.json").write_text(json.dumps(summary, indent=1))
    (PROF / "sass_histogram_r2.json").write_text(json.dumps(sass_histogram(), indent=1))
    print(json.dumps({"per_window": per_window, "mean": mean, "share": {k: v["share"] for k, v in ll["by_kernel"].items()}}, indent=1))


if __name__ == "__main__":
    main()
