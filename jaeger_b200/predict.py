"""`jaeger predict` on the B200 engine: model lookup, engine call, post-processing, outputs.

Mirrors the reference driver `commands/predict.py:488-861` for the options that touch the hot
path (the option names and defaults are the reference CLI's, cli.py:122-371, including the experimental --crf family): same model
registry (`config.json["model_paths"]`, scanned like `AvailableModels`, utils/misc.py:334-392),
same output locations `<-o>/<model_id>/<base>.tsv` and `<base>_phages.tsv`, same `--overwrite`
rule, same prophage table source, same `--refine` family.  Options outside the path (plots, --onnx ...) are not accepted.

    python -m jaeger_b200.predict -i contigs.fasta -o out -m jaeger_xxx_1.4M_fragment --config config.json
    python -m jaeger_b200.predict -i contigs.fasta -o out -m standin      # random-init stand-in architecture
"""
from __future__ import annotations

import argparse
import json
import logging
import sys
import time
from collections import defaultdict
from pathlib import Path
from typing import Any

import numpy as np

logger = logging.getLogger("Jaeger")


def available_models(paths) -> dict[str, dict[str, Path]]:
    """utils/misc.py:346-392: every directory named `model` under each path."""
    models: dict[str, dict[str, Path]] = defaultdict(dict)
    for path in [Path(p) for p in ([paths] if isinstance(paths, (str, Path)) else paths)]:
        model_dirs = [p for p in path.rglob("model") if p.is_dir()]
        if path.name == "model" and path.is_dir():
            model_dirs.append(path)
        for md in model_dirs:
            for e in md.iterdir():
                if e.is_dir() and e.name.endswith("_graph"):
                    models[e.name.removesuffix("_graph")]["graph"] = e
                elif e.is_file():
                    for suffix, key in (("_classes.yaml", "classes"), ("_project.yaml", "project"), (".weights.h5", "weights")):
                        if e.name.endswith(suffix):
                            models[e.name[:-len(suffix)]][key] = e
    return dict(models)


def get_model_id(model: str) -> str:
    """utils/misc.py:395."""
    return model.split("_", 1)[1].rsplit("_", 1)[0]


def run_core_legacy(**kwargs: Any) -> dict[str, Any]:
    """`jaeger predict -m default` (commands/predict_legacy.py:34-357) on the B200 engine: the bundled
    legacy graph, the legacy encoder, pred_to_dict_legacy / write_output_legacy tables.
    legacy_weights: a SavedModel `variables/` directory holding the graph's tensors, or an .npz written
    by jaeger_b200.weights.save_npz_weights.  legacy_ood_dir: the reference's data/models/default
    directory (LR_ood_4_class_default.pkl, batch_means.npy, batch_std.npy); omitted -> no reliability."""
    from . import B200Engine, WindowSource, legacy
    from .postprocess import contig_table_legacy, write_output_legacy
    from .weights import load_npz_weights, read_tf_bundle

    t0 = time.time()
    input_path = Path(kwargs["input"])
    fsize, stride = int(kwargs.get("fsize", 2000)), int(kwargs.get("stride", 1500))
    min_len = kwargs.get("min_len") or fsize
    if min_len < fsize:                                                  # predict_legacy.py:57-64
        logger.warning(f"--min-len < --fsize is not supported in legacy prediction mode; using --min-len={fsize}.")
        min_len = fsize
    if not kwargs.get("legacy_weights"):
        raise ValueError("-m default needs --legacy-weights: the bundled model's SavedModel variables/ directory "
                         "(<jaeger>/data/models/test/jaeger_fragment_graph/variables) or an .npz written by jaeger_b200.weights.save_npz_weights")
    wpath = Path(kwargs["legacy_weights"])
    if not wpath.exists():
        raise ValueError(f"--legacy-weights {wpath} does not exist")
    if int(__import__("os").environ.get("WORLD_SIZE", "1")) > 1:
        raise ValueError("-m default runs in one process (the legacy path is not sharded): launch it without torchrun")
    weights = load_npz_weights(wpath) if wpath.suffix == ".npz" else legacy.weights_from_bundle(read_tf_bundle(wpath))
    ood = legacy.load_ood_params(kwargs["legacy_ood_dir"]) if kwargs.get("legacy_ood_dir") else None
    engine = B200Engine(legacy_weights=weights, all_labels=bool(kwargs.get("getalllabels")), device=int(kwargs.get("physicalid") or 0))
    out_dir = Path(kwargs["output"]) / "default"
    out_dir.mkdir(parents=True, exist_ok=True)
    base = input_path.stem
    table, phage_table = out_dir / f"{base}_jaeger.tsv", out_dir / f"{base}_phages_jaeger.tsv"     # predict_legacy.py:71-72
    if table.exists() and not kwargs.get("overwrite"):
        raise FileExistsError(f"{table} exists; use --overwrite")
    src = WindowSource(fasta=input_path, fsize=fsize, stride=stride, min_len=None,
                       dynamic_stride=bool(kwargs.get("dynamic_stride", False)),
                       dynamic_stride_threshold=float(kwargs.get("dynamic_stride_threshold", 10.0)),
                       batch=int(kwargs.get("batch", 96)), dustmask=bool(kwargs.get("dustmask", True)),
                       outputs=("prediction", "embedding"), lazy_meta=True)
    rec_off = src.load()[2]
    if not (np.diff(rec_off) >= min_len).any():
        raise ValueError(f"all records in {input_path} are < {min_len}bp")
    y_pred = engine.predict(src)
    t1 = time.time()
    term = None
    if kwargs.get("terminal_repeats", True):
        from .termini import scan_source
        term = scan_source(engine, src, fsize)
    data = contig_table_legacy(engine, y_pred, fsize, ood, term_repeats=term)
    labels = legacy.ALL_LABELS if kwargs.get("getalllabels") else legacy.DEFAULT_LABELS
    n_written = write_output_legacy(data, [labels[i] for i in range(4)], table, phage_table,
                                    reliability_cutoff=float(kwargs.get("rc", 0.5)), phage_score=float(kwargs.get("pc", 3)))
    logger.info(f"processed {n_written}/{len(rec_off) - 1} sequences")
    return {"table": table, "phage_table": phage_table, "num_written": n_written, "num": len(rec_off) - 1,
            "windows": int(y_pred["prediction"].shape[0]), "predict_seconds": t1 - t0, "data": data}


def _make_engine(kwargs: dict[str, Any]):
    """Model lookup (commands/predict.py:494-551) -> (engine, model_id, model_name, registry entry or None)."""
    from . import B200Engine, parse_project, standin_1p4m_config
    model_name = kwargs.get("model") or "default"                           # cli.py: the reference's default model
    if model_name == "standin" and not kwargs.get("model_path") and not kwargs.get("allow_random_weights"):
        raise ValueError("-m standin is a RANDOM-INITIALISED network (the benchmark architecture): its tables are meaningless. "
                         "Pass --allow-random-weights to run it anyway.")
    if kwargs.get("cpu"):
        raise RuntimeError("--cpu: this engine has no CPU path (use the reference's own engine for CPU runs)")
    for other in ("onnx", "quantized", "xla"):                           # the reference's alternate backends (predict.py:687-745)
        if kwargs.get(other):
            raise RuntimeError(f"--{other}: this is the B200 engine; the ONNX / TFLite / XLA backends belong to the reference's own driver")
    precision = kwargs.get("precision") or "fp16"                          # predict.py:604-613
    if precision not in ("fp32", "fp16", "bf16"):
        raise ValueError(f"--precision {precision!r} (use fp32, fp16 or bf16)")
    if precision != "fp16":
        # not silently ignored: this engine has ONE numeric mode (fp16 activations / weights, fp32 accumulation and heads)
        raise ValueError(f"--precision {precision} is not available on the B200 engine: it computes with fp16 activations and "
                         "weights, fp32 accumulation and fp32 heads (logits within the tolerance stated in DESIGN.md). Use --precision fp16.")
    # --mem (GB, predict.py:544, 615-623) caps the device workspace the window chunks are sized from
    workspace_gb = float(kwargs["mem"]) if kwargs.get("mem") else 16.0
    device = int(kwargs.get("physicalid") or 0)
    info = None
    if model_name == "standin" and not kwargs.get("model_path"):
        logger.warning("RANDOM-INITIALISED stand-in network (--allow-random-weights): the classifications below carry no meaning")
        engine = B200Engine(spec=parse_project(standin_1p4m_config()), device=device, workspace_gb=workspace_gb)
        model_id = "standin"
    elif kwargs.get("model_path"):                                         # predict.py:503-542
        info = available_models(kwargs["model_path"])
        if not info:
            raise ValueError(f"No model found in {kwargs['model_path']}")
        cls_models = {n: m for n, m in info.items() if m.get("graph") is not None and m.get("classes") is not None}
        if not cls_models:
            raise ValueError(f"No classification model found in {kwargs['model_path']}. Expected a *_graph directory, "
                             "*_classes.yaml, *_project.yaml, and *.weights.h5 files.")
        if len(cls_models) > 1:
            cls_models = {n: m for n, m in cls_models.items() if not n.endswith("_embedding")} or cls_models
        model_name = next(iter(cls_models))
        engine = B200Engine(cls_models[model_name], device=device, workspace_gb=workspace_gb)
        model_id = get_model_id(model_name)
    else:
        cfg = json.loads(Path(kwargs["config"]).read_text()) if kwargs.get("config") else {}
        info = available_models(cfg.get("model_paths", []))
        if model_name not in info:
            raise ValueError(f"model {model_name!r} not found; available: {sorted(info)}")
        engine = B200Engine(info[model_name], device=device, workspace_gb=workspace_gb)
        model_id = get_model_id(model_name)
    spc = getattr(engine, "string_processor_config", None) or {}        # predict.py:752-766
    fsize = int(kwargs.get("fsize", 2000))
    if spc.get("crop_size_nt") is not None:
        logger.info(f"model trained fragment length: {spc.get('crop_size_codons')} codons ({spc['crop_size_nt']} nt)")
    from .modelspec import crop_length_warning
    crop_msg = crop_length_warning(spc.get("crop_size_codons"), spc.get("crop_size_nt"), fsize)
    if crop_msg is not None:
        logger.warning(crop_msg)
    return engine, model_id, model_name, info


def _window_source(kwargs: dict[str, Any], fsize: int, stride: int, **where):
    from . import WindowSource
    make = WindowSource.from_host if "names" in where else WindowSource
    return make(**where, fsize=fsize, stride=stride, min_len=kwargs.get("min_len"),
                dynamic_stride=bool(kwargs.get("dynamic_stride", False)),
                dynamic_stride_threshold=float(kwargs.get("dynamic_stride_threshold", 10.0)),
                batch=int(kwargs.get("batch", 96)), dustmask=bool(kwargs.get("dustmask", True)),    # cli.py: --dustmask default on
                outputs=("prediction", "reliability") + (("embedding",) if kwargs.get("save_embedding") else ())
                + (("nmd",) if kwargs.get("save_nmd") else ()),          # the tables need neither embeddings nor NMD vectors
                lazy_meta=True)


def _classify_source(engine, src, kwargs: dict[str, Any], fsize: int, stride: int, model_name: str, info) -> dict[str, Any]:
    """One set of records through stages 1-4 (+ terminal repeats, --refine, -p): the per-contig summary frame and what hangs
    off it.  Shared by the whole-file run and by every chunk of a streamed run."""
    from .postprocess import contig_table, generate_summary
    from .prophage import call_regions
    t0 = time.time()
    y_pred = engine.predict(src)
    t1 = time.time()
    n_windows = int(y_pred["prediction"].shape[0]) if y_pred else 0
    logger.info(f"classified {n_windows} windows in {t1 - t0:.2f} s")
    crf_cost, crf_matrix = None, kwargs.get("crf_transition_matrix")          # predict.py:288-307
    if kwargs.get("crf"):
        crf_cost = float(kwargs.get("crf_switch_cost", 2.0))
        if isinstance(crf_matrix, (str, Path)):
            crf_matrix = json.loads(Path(crf_matrix).read_text())
    term = None
    if kwargs.get("terminal_repeats", True) and y_pred:                       # predict.py:679-685 (always on in the reference)
        from .termini import scan_source
        t_term = time.time()
        term = scan_source(engine, src, fsize)
        logger.info(f"terminal-repeat scan in {time.time() - t_term:.2f} s")
    data = contig_table(engine, y_pred, fsize, term_repeats=term, crf_switch_cost=crf_cost,
                        crf_prior=kwargs.get("crf_prior", "biological"), crf_transition_matrix=crf_matrix) if y_pred else None
    cm = engine.class_map
    refined = None
    if kwargs.get("refine") and data:                                          # predict.py:310-335
        from .refine import load_refinement, refined_contig_table
        refine_path = Path(kwargs["refine_file"]) if kwargs.get("refine_file") else (
            Path(info[model_name]["graph"]).parent / f"{model_name}_refine.yaml" if (info and model_name in info) else None)
        if refine_path is not None and refine_path.exists():
            try:
                refine_cfg = load_refinement(refine_path, expect_model=model_name)
                refined = refined_contig_table(engine, data["headers"], data["predictions"], data["offsets"], refine_cfg["taus"],
                                               mode=kwargs.get("refine_mode", "gated"), min_windows=int(kwargs.get("refine_min_windows", 3)),
                                               merge_split=kwargs.get("refine_merge_split", "half"),
                                               allow_merged_contig_call=bool(kwargs.get("refine_allow_merged_contig_call", False)),
                                               contig_hedge_margin=float(kwargs.get("refine_contig_hedge_margin", 1.0)))
                logger.info(f"Applied refinement calibration from {refine_path}")
            except Exception as e:
                logger.warning(f"Refinement failed: {e}; using default predictions")
        else:
            logger.warning(f"No refinement calibration found at {refine_path}; using default predictions")
    df = generate_summary(data, cm["class"], cm["index"], refined_contig=refined) if data else None
    regions = None
    if kwargs.get("prophage") and data:
        regions = call_regions(engine, data, cm, fsize, stride, lc=int(kwargs.get("lc", 500_000)),
                               sensitivity=float(kwargs.get("sensitivity", 1.5)))
    return {"y_pred": y_pred, "data": data, "df": df, "regions": regions, "windows": n_windows, "predict_seconds": t1 - t0}


def _write_prophage_outputs(engine, loaded, regions, out_dir: Path, base: str, fsize: int, stride: int, result: dict, append: bool = False,
                            genes: str | None = None):
    """<base>_prophage_regions.tsv and the att-site report <base>_prophages/prophages_jaeger.tsv (predict.py:372-376, 386-394,
    430-437; postprocess/prophages.py:706-873).  Region ends are snapped out of coding genes first (prophage_boundaries.py) when
    gene calls are available: the `--genes` table, or pyrodigal-gv when that package is installed; else the raw ends are used."""
    rows = [] if append else ["contig_id\tstart\tend\twindow_start\twindow_end\tscore"]
    for name, r in regions.items():
        for (ws, we), (s, e), sc in zip(r["ranges"], r["coords"], r["scores"]):
            rows.append(f"{name}\t{s}\t{e}\t{ws}\t{we}\t{sc:.3f}")
    with open(out_dir / f"{base}_prophage_regions.tsv", "a" if append else "w") as fh:
        fh.write("\n".join(rows) + ("\n" if rows else ""))
    result.setdefault("prophage_regions", {}).update(regions)
    from .termini import prophage_report_loaded, write_prophage_report
    t_att = time.time()
    try:                                              # the reference logs and carries on (predict.py:440-442)
        from . import prophage_boundaries as pb
        refined = None
        genes_of = pb.gene_source(pb.load_gene_table(genes) if genes else None, loaded)
        if genes_of is not None:
            refined = pb.refine_regions(regions, loaded[0], np.diff(loaded[2]), fsize, stride, genes_of)
            moved = sum(1 for rows in refined.values() for r in rows if (r[0], r[1]) != (r[2], r[3]))
            logger.info(f"gene-aware boundaries: {moved} of {sum(len(v) for v in refined.values())} regions moved")
            result.setdefault("refined_boundaries", {}).update(refined)
        report = prophage_report_loaded(engine, loaded, regions, fsize, stride, refined_boundaries=refined)
        if append and result.get("prophage_report") is not None:
            import pandas as pd
            report = pd.concat([result["prophage_report"], report], ignore_index=True)
        write_prophage_report(report, out_dir / f"{base}_prophages")
        result["prophage_report"] = report
        logger.info(f"prophage att-site report ({len(report)} regions) in {time.time() - t_att:.2f} s")
    except (ArithmeticError, ValueError, KeyError, IndexError) as e:
        result.setdefault("prophage_report", None)
        logger.error(f"an error {e!r} occurred during the prophage report step")


def run_core(**kwargs: Any) -> dict[str, Any]:
    if (kwargs.get("model") or "default") == "default" and not kwargs.get("model_path"):     # --model_path picks the model itself (predict.py:503-542)
        return run_core_legacy(**kwargs)
    from . import WindowSource
    from .parallel import dist_env, merge_rank_frames, shard_contigs, shard_loaded
    from .postprocess import write_tables

    t0 = time.time()
    input_path = Path(kwargs["input"])
    fsize, stride = int(kwargs.get("fsize", 2000)), int(kwargs.get("stride", 1500))
    # one process per GPU under torchrun: contigs are sharded over the ranks, rank 0 writes
    world, rank, local_rank = dist_env()
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        kwargs["physicalid"] = local_rank
    # inputs that should not be held whole are streamed in chunks of whole records (ingest.py)
    stream_mbp = kwargs.get("stream_mbp")
    if stream_mbp is None and input_path.exists() and input_path.stat().st_size > int(kwargs.get("stream_above_gb", 2.0) * 1e9):
        stream_mbp = 512.0
    if stream_mbp:
        return _run_core_streaming(kwargs, float(stream_mbp), world, rank, t0)
    engine, model_id, model_name, info = _make_engine(kwargs)
    out_dir = Path(kwargs["output"]) / model_id                          # predict.py:551
    out_dir.mkdir(parents=True, exist_ok=True)
    base = input_path.stem
    table, phage_table = out_dir / f"{base}.tsv", out_dir / f"{base}_phages.tsv"      # predict.py:571-572
    if table.exists() and not kwargs.get("overwrite"):
        raise FileExistsError(f"{table} exists; use --overwrite")                     # predict.py:574-578
    min_len = kwargs.get("min_len")
    src = _window_source(kwargs, fsize, stride, fasta=input_path)
    t_load = time.time()
    rec_off = src.load()[2]
    n_records = len(rec_off) - 1
    logger.info(f"read {n_records} records, {int(rec_off[-1])} bases in {time.time() - t_load:.2f} s")
    if not (np.diff(rec_off) >= (min_len or fsize)).any():
        raise ValueError(f"all records in {input_path} are < {min_len or fsize}bp")   # utils/fs.py:99-115
    mine = np.arange(n_records)
    if world > 1:
        # balance on long-pass window counts; a short contig of the two-pass mode is one window
        lens = np.diff(rec_off)
        eff = np.where((lens < fsize) & (lens >= (min_len or fsize)), fsize, lens)
        mine = shard_contigs(eff, world, fsize, stride)[rank]
        src._loaded = shard_loaded(src.load(), mine)
    if kwargs.get("crf"):
        logger.warning("CRF window decoding is experimental; results may change between releases")
    t_post = time.time()
    from .parallel import exchange_errors
    try:
        res, failure = _classify_source(engine, src, kwargs, fsize, stride, model_name, info), None
    except Exception as e:              # told to the other ranks before anybody enters the gather
        res, failure = None, e
    exchange_errors(failure, world, rank)
    y_pred, data, df, regions, n_windows = res["y_pred"], res["data"], res["df"], res["regions"], res["windows"]
    t1 = t_post + res["predict_seconds"]
    cm = engine.class_map
    if world > 1:
        if df is not None:       # row -> (pass, index of the contig in the FASTA), the single-process row order
            first = data["offsets"][:-1]
            local = engine.windows.contig[first]
            df["_pass"] = (engine.windows.seqlen[first] < fsize).astype(np.int64)
            df["_gid"] = mine[local]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({"df": df, "regions": regions, "windows": n_windows}, gathered, dst=0)
        if kwargs.get("window_scores") and data:       # per-window arrays stay with the rank that owns the contigs
            off = data["offsets"]
            np.savez(out_dir / f"{base}_window_scores.rank{rank}.npz", headers=data["headers"], lengths=data["length"],
                     predictions=np.array([data["predictions"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object),
                     gc_skews=np.array([data["gc_skews"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object),
                     gcs=np.array([data["gcs"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object), allow_pickle=True)
            kwargs["window_scores"] = False
        if rank != 0:
            return {"table": table, "phage_table": phage_table, "rank": rank, "windows": n_windows, "predict_seconds": t1 - t0}
        df = merge_rank_frames([g["df"] for g in gathered])
        n_windows = sum(g["windows"] for g in gathered)
        if kwargs.get("prophage"):
            regions = {k: v for g in gathered if g["regions"] for k, v in g["regions"].items()}
    n_written = write_tables(df, cm["class"], bool(data.get("has_reliability", True)) if data else True, table, phage_table,
                             reliability_cutoff=float(kwargs.get("rc", 0.1)), phage_score=float(kwargs.get("pc", 3))) if df is not None else 0
    result = {"table": table, "phage_table": phage_table, "num_written": n_written, "num": n_records,
              "windows": n_windows, "predict_seconds": t1 - t0}
    logger.info(f"aggregation + tables in {time.time() - t_post - res['predict_seconds']:.2f} s")
    logger.info(f"processed {n_written}/{n_records} sequences")
    if kwargs.get("prophage"):
        full = WindowSource(fasta=input_path).load() if world > 1 else src.load()
        _write_prophage_outputs(engine, full, regions or {}, out_dir, base, fsize, stride, result, genes=kwargs.get("genes"))
    if kwargs.get("getsequences"):                                        # predict.py:444-455
        from .postprocess import write_fasta_from_results
        full = WindowSource(fasta=input_path).load() if world > 1 else src.load()
        n_seq = write_fasta_from_results(full, phage_table, out_dir / f"{base}_phages_jaeger.fasta")
        logger.info(f"{base}_phages_jaeger.fasta created ({n_seq} records)")
    if y_pred and world == 1:                                             # _save_auxiliary_outputs, predict.py:66-112
        headers = y_pred["meta_0"] if (kwargs.get("save_embedding") or kwargs.get("save_nmd")) else None
        if kwargs.get("save_embedding") and "embedding" in y_pred:
            np.savez(out_dir / f"{base}_embedding.npz", embedding=y_pred["embedding"], headers=headers)
            logger.info(f"{base}_embedding.npz created")
        if kwargs.get("save_nmd") and "nmd" in y_pred:
            np.savez(out_dir / f"{base}_nmd.npz", embedding=y_pred["nmd"], headers=headers)   # the reference keeps the key name "embedding"
            logger.info(f"{base}_nmd.npz created")
    if kwargs.get("window_scores") and data:                              # predict.py:458-470
        off = data["offsets"]
        np.savez(out_dir / f"{base}_window_scores.npz", headers=data["headers"], lengths=data["length"],
                 predictions=np.array([data["predictions"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object),
                 gc_skews=np.array([data["gc_skews"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object),
                 gcs=np.array([data["gcs"][off[i]:off[i + 1]] for i in range(len(off) - 1)], dtype=object), allow_pickle=True)
    return result


def _run_core_streaming(kwargs: dict[str, Any], stream_mbp: float, world: int, rank: int, t0: float) -> dict[str, Any]:
    """`jaeger predict` over an input that is never held whole (BASELINE config 5): chunks of whole records stream through two
    pinned host buffers (ingest.FastaChunks: the next chunk is parsed while the current one is on the device), every chunk goes
    through stages 1-4 and leaves only its per-contig rows behind, so host and device memory are bounded by the chunk size,
    not by the file.  Under torchrun every rank streams its own byte slice of a plain-text file (a gzip stream is read by every
    rank, which keeps every world-th chunk); rank 0 gathers the per-contig tables and writes them in the single-process row
    order (long-pass contigs in file order, then short-pass contigs)."""
    import pandas as pd
    from .ingest import FastaChunks, is_gzip, rank_byte_range
    from .parallel import merge_rank_frames, normalise_joined_columns
    from .postprocess import write_tables
    for opt in ("window_scores", "save_embedding", "save_nmd"):
        if kwargs.get(opt):
            raise NotImplementedError(f"--{opt.replace('_', '-')} keeps per-window arrays of the whole run in memory and is not available "
                                      "when the input is streamed (pass --stream-mbp 0 to load the file whole)")
    input_path = Path(kwargs["input"])
    fsize, stride = int(kwargs.get("fsize", 2000)), int(kwargs.get("stride", 1500))
    min_len = kwargs.get("min_len") or fsize
    engine, model_id, model_name, info = _make_engine(kwargs)
    out_dir = Path(kwargs["output"]) / model_id
    out_dir.mkdir(parents=True, exist_ok=True)
    base = input_path.stem
    table, phage_table = out_dir / f"{base}.tsv", out_dir / f"{base}_phages.tsv"
    if table.exists() and not kwargs.get("overwrite"):
        raise FileExistsError(f"{table} exists; use --overwrite")
    gz = is_gzip(input_path)
    byte_range = (0, -1) if (world == 1 or gz) else rank_byte_range(input_path, rank, world)
    chunks = FastaChunks(input_path, chunk_bases=int(stream_mbp * 1e6), byte_range=byte_range)
    frames, n_windows, n_records, n_eligible, predict_s, gid0, n_chunks = [], 0, 0, 0, 0.0, 0, 0
    rank_term = 0 if (gz or world == 1) else (rank << 40)          # byte slices are in file order: rank-major row order
    result: dict[str, Any] = {"table": table, "phage_table": phage_table}
    cm = engine.class_map
    has_rel = True
    first_prophage = True
    failure = None
    for k, (names, bases, offsets) in enumerate(chunks):
        if failure is not None:
            break
        lens = np.diff(offsets)
        n_here = len(names)
        mine_chunk = world == 1 or not gz or (k % world == rank)
        if mine_chunk:
            n_records += n_here
            n_eligible += int((lens >= min_len).sum())
        if mine_chunk and (lens >= min_len).any():
            src = _window_source(kwargs, fsize, stride, names=names, bases=bases, offsets=offsets)
            try:
                res = _classify_source(engine, src, kwargs, fsize, stride, model_name, info)
            except Exception as e:      # kept until every rank reaches the exchange below (no collective inside the loop)
                failure = e
                continue
            predict_s += res["predict_seconds"]
            n_windows += res["windows"]
            if res["df"] is not None:
                data, df = res["data"], res["df"]
                first = data["offsets"][:-1]
                df["_pass"] = (engine.windows.seqlen[first] < fsize).astype(np.int64)
                df["_gid"] = rank_term + gid0 + engine.windows.contig[first].astype(np.int64)
                frames.append(df)
                has_rel = bool(data.get("has_reliability", True))
            if kwargs.get("prophage") and res["regions"]:
                _write_prophage_outputs(engine, src.load(), res["regions"], out_dir, base, fsize, stride, result, append=not first_prophage,
                                        genes=kwargs.get("genes"))
                first_prophage = False
            logger.info(f"chunk {k}: {n_here} records, {int(offsets[-1])} bases, {res['windows']} windows")
        gid0 += n_here
        n_chunks += 1
    from .parallel import exchange_errors
    exchange_errors(failure, world, rank)
    df = normalise_joined_columns(pd.concat(frames, ignore_index=True)) if frames else None
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object({"df": df, "windows": n_windows, "records": n_records, "eligible": n_eligible}, gathered, dst=0)
        if rank != 0:
            return {"table": table, "phage_table": phage_table, "rank": rank, "windows": n_windows, "predict_seconds": predict_s}
        df = merge_rank_frames([g["df"] for g in gathered], keep_order_columns=True)
        n_windows, n_records, n_eligible = (sum(g[key] for g in gathered) for key in ("windows", "records", "eligible"))
    if n_eligible == 0:
        raise ValueError(f"all records in {input_path} are < {min_len}bp")            # utils/fs.py:99-115
    if df is not None and len(df):
        df = df.sort_values(["_pass", "_gid"], kind="stable").reset_index(drop=True).drop(columns=["_pass", "_gid"])
    n_written = write_tables(df, cm["class"], has_rel, table, phage_table, reliability_cutoff=float(kwargs.get("rc", 0.1)),
                             phage_score=float(kwargs.get("pc", 3))) if df is not None else 0
    result.update({"num_written": n_written, "num": n_records, "windows": n_windows, "predict_seconds": predict_s,
                   "streamed_chunks": n_chunks, "wall_seconds": time.time() - t0})
    logger.info(f"processed {n_written}/{n_records} sequences in {result['streamed_chunks']} chunks")
    if kwargs.get("getsequences"):           # a second streaming pass over the input picks the records of the phage table
        from .postprocess import write_fasta_from_results
        n_seq = 0
        out_fa = out_dir / f"{base}_phages_jaeger.fasta"
        out_fa.write_bytes(b"")
        for names, bases, offsets in FastaChunks(input_path, chunk_bases=int(stream_mbp * 1e6)):
            n_seq += write_fasta_from_results((names, bases, offsets), phage_table, out_fa, append=True)
        logger.info(f"{base}_phages_jaeger.fasta created ({n_seq} records)")
    return result


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="jaeger_b200 predict", description=__doc__.split("\n")[0])
    ap.add_argument("-i", "--input", required=True)
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("-m", "--model", default="default", help="model name from the config's model_paths; `default` = the bundled legacy model")
    ap.add_argument("--allow-random-weights", dest="allow_random_weights", action="store_true",
                    help="required for -m standin: the random-initialised benchmark architecture (meaningless tables)")
    ap.add_argument("--config", default=None, help="config.json with model_paths (utils/misc.py:309-331)")
    ap.add_argument("--legacy-weights", dest="legacy_weights", default=None, help="-m default: SavedModel variables/ dir or .npz")
    ap.add_argument("--legacy-ood-dir", dest="legacy_ood_dir", default=None, help="-m default: dir with the reliability model files")
    ap.add_argument("--getalllabels", action="store_true")
    ap.add_argument("--fsize", type=int, default=2000)
    ap.add_argument("--stride", type=int, default=1500)
    ap.add_argument("--batch", type=int, default=96)
    ap.add_argument("--min-len", dest="min_len", type=int, default=None)
    ap.add_argument("--dynamic-stride", dest="dynamic_stride", action="store_true")
    ap.add_argument("--dynamic-stride-threshold", dest="dynamic_stride_threshold", type=float, default=10.0)
    ap.add_argument("--dustmask", dest="dustmask", action="store_true", default=True)
    ap.add_argument("--no-dustmask", dest="dustmask", action="store_false")
    ap.add_argument("--refine", action="store_true", help="apply post-hoc refinement using the model's <model>_refine.yaml")
    ap.add_argument("--refine-file", dest="refine_file", default=None, help="calibration file to use instead of the model's own")
    ap.add_argument("--refine-mode", dest="refine_mode", choices=["gated", "weighted", "unweighted"], default="gated")
    ap.add_argument("--refine-min-windows", dest="refine_min_windows", type=int, default=3)
    ap.add_argument("--refine-merge-split", dest="refine_merge_split", choices=["half", "full"], default="half")
    ap.add_argument("--refine-allow-merged-contig-call", dest="refine_allow_merged_contig_call", action="store_true")
    ap.add_argument("--refine-contig-hedge-margin", dest="refine_contig_hedge_margin", type=float, default=1.0)
    ap.add_argument("--crf", action="store_true", help="(experimental) joint Viterbi decoding of window labels")
    ap.add_argument("--crf-switch-cost", dest="crf_switch_cost", type=float, default=2.0)
    ap.add_argument("--crf-prior", dest="crf_prior", choices=["biological", "uniform"], default="biological")
    ap.add_argument("--crf-transition-matrix", dest="crf_transition_matrix", default=None)
    ap.add_argument("--no-terminal-repeats", dest="terminal_repeats", action="store_false", default=True,
                    help="skip the terminal-repeat scan (the reference always runs it)")
    ap.add_argument("--rc", type=float, default=0.1)
    ap.add_argument("--pc", type=float, default=3)
    ap.add_argument("-p", "--prophage", action="store_true")
    ap.add_argument("--lc", type=int, default=500_000)
    ap.add_argument("--genes", default=None, help="gene calls (GFF3 / BED / TSV contig,begin,end) used to snap prophage ends out of "
                                                  "coding genes (the reference calls pyrodigal-gv itself; used here too when installed)")
    ap.add_argument("-s", "--sensitivity", type=float, default=1.5)
    ap.add_argument("--physicalid", type=int, default=0)
    ap.add_argument("--model_path", dest="model_path", default=None, help="directory holding the model instead of the config's model_paths")
    ap.add_argument("--mem", type=float, default=None, help="device workspace cap in GB (default 16)")
    ap.add_argument("--precision", choices=["fp32", "fp16", "bf16"], default="fp16")
    ap.add_argument("--cpu", action="store_true", help="rejected: there is no CPU path")
    for other in ("--onnx", "--quantized", "--xla"):
        ap.add_argument(other, action="store_true", help="rejected: an alternate backend of the reference's own driver")
    ap.add_argument("--workers", type=int, default=4, help="accepted for CLI compatibility: windowing / encoding run on the device")
    ap.add_argument("--plot-type", dest="plot_type", default="none", choices=["circular", "linear", "both", "none"],
                    help="accepted for CLI compatibility: plots are outside the hot path and are not drawn")
    ap.add_argument("-v", "--verbose", action="count", default=1, help="-vv debug, -v info")
    ap.add_argument("--save-embedding", dest="save_embedding", action="store_true", help="write <base>_embedding.npz")
    ap.add_argument("--save-nmd", dest="save_nmd", action="store_true", help="write <base>_nmd.npz")
    ap.add_argument("--window-scores", dest="window_scores", action="store_true")
    ap.add_argument("--getsequences", action="store_true", help="write the records of the phage table to <base>_phages_jaeger.fasta")
    ap.add_argument("--overwrite", action="store_true")
    ap.add_argument("--stream-mbp", dest="stream_mbp", type=float, default=None,
                    help="stream the input in chunks of this many Mbp of whole records (bounded memory); default: 512 for inputs above 2 GB, 0 = never")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose >= 2 else logging.INFO, format="%(asctime)s %(levelname)s [jaeger_b200] %(message)s")
    if args.plot_type != "none" and args.prophage:
        logger.warning(f"--plot-type {args.plot_type}: prophage plots are not drawn by this engine (tables only)")
    try:
        res = run_core(**vars(args))
    except Exception as e:                                               # predict.py:811-816
        logger.error(f"an error {e} occured during inference")
        return 1
    finally:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    if res.get("rank"):
        return 0
    logger.info(f"wrote {res['table']} ({res['num_written']} contigs, {res['windows']} windows, {res['predict_seconds']:.2f} s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
