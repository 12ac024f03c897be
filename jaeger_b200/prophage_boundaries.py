"""Gene-aware refinement of prophage boundaries (reference: postprocess/prophage_boundaries.py:52-193, driver glue
commands/predict.py:386-394, consumer postprocess/prophages.py:759-772).

The segmentation step gives region ends on the window grid; the reference snaps every end that falls inside a coding gene to that
gene's outer end (left ends to the gene start, right ends to the gene end, by at most `2 * fsize` bases) so that a called prophage
never stops inside a gene, and runs the att-site search on the refined ends.  The gene calls come from pyrodigal-gv, a third-party
gene finder that is outside this path's scope: here the caller supplies them -- a gene table given with `--genes` (GFF3 / BED / TSV
of any gene caller, e.g. prodigal-gv run once per assembly) or, when the `pyrodigal_gv` package is importable, the same
`ViralGeneFinder(meta=True)` the reference uses.  Without gene calls the raw boundaries are used, which is the reference's
`refined_boundaries=None` path.  Pure host code: a few intervals per called region.
"""
from __future__ import annotations

import bisect
import gzip
import logging
from pathlib import Path
from typing import Any, Callable, Iterable

import numpy as np

logger = logging.getLogger("jaeger_b200")

Genes = list[tuple[int, int]]          # sorted 0-based half-open intervals


def is_intergenic(position: int, genes: Genes) -> bool:
    """prophage_boundaries.py:52-59."""
    for start, end in genes:
        if start <= position < end:
            return False
        if start > position:
            break
    return True


def refine_boundary(position: int, genes: Genes, side: str, max_extension: int | None = None) -> int:
    """prophage_boundaries.py:62-113: a boundary inside a gene moves to that gene's start (left) / end (right), by at most
    `max_extension` bases; the FIRST gene (in sorted order) containing the position decides, as in the reference."""
    if side not in {"left", "right"}:
        raise ValueError(f"side must be 'left' or 'right', got {side!r}")
    containing = next(((s, e) for s, e in genes if s <= position < e), None)
    if containing is None:
        return position
    refined = containing[0] if side == "left" else containing[1]
    if max_extension is not None and abs(refined - position) > max_extension:
        logger.warning("Boundary refinement exceeded max_extension (%d bp); capping %s boundary.", max_extension, side)
        refined = position + max_extension if side == "right" else position - max_extension
    return refined


def refine_region(raw_start: int, raw_end: int, genes: Genes, max_extension: int | None = None) -> tuple[int, int]:
    """prophage_boundaries.py:116-137."""
    return (refine_boundary(raw_start, genes, "left", max_extension=max_extension),
            refine_boundary(raw_end, genes, "right", max_extension=max_extension))


def refine_regions(regions: dict[str, Any], names: Iterable[str], lengths: Iterable[int], fsize: int, stride: int | None,
                   genes_of: Callable[[str, int], Genes | None], max_extension: int | None = None) -> dict[str, list[tuple[int, int, int, int]]]:
    """refine_prophage_boundaries (prophage_boundaries.py:140-193) over the called regions: header ->
    [(raw_start, raw_end, refined_start, refined_end)], clipped to the contig.  `regions`: header -> {"ranges", "scores"}
    (prophage.call_regions) or the reference's [ranges, scores] pair; `genes_of(header, index)` returns the contig's genes or
    None when there are none on record (then the contig keeps its raw boundaries, like a contig without predicted genes)."""
    if max_extension is None:
        max_extension = 2 * fsize
    step = stride or fsize
    out: dict[str, list[tuple[int, int, int, int]]] = {}
    for ci, (name, length) in enumerate(zip(names, lengths)):
        header = name.strip().replace(",", "___")
        if header not in regions:
            continue
        reg = regions[header]
        cords = reg["ranges"] if isinstance(reg, dict) else reg[0]
        if len(cords) == 0:
            out[header] = []
            continue
        genes = genes_of(header, ci) or []
        rows = []
        for start_idx, end_idx in cords:
            raw_start, raw_end = int(start_idx * step), int((end_idx - 1) * step + fsize)
            r_start, r_end = refine_region(raw_start, raw_end, genes, max_extension=max_extension)
            rows.append((raw_start, raw_end, max(r_start, 0), min(r_end, int(length))))
        out[header] = rows
    return out


_TABLES: dict[tuple[str, float], dict[str, Genes]] = {}      # parsed gene tables (a streamed run asks once per chunk)


def load_gene_table(path: str | Path) -> dict[str, Genes]:
    """Gene intervals per contig from a gene caller's output: GFF3 / GTF (columns 1, 4, 5: 1-based closed; `CDS` / `gene`
    features), BED (columns 1-3: 0-based half-open) or a headerless TSV `contig, begin, end` with 1-based closed coordinates
    (what pyrodigal's `Gene.begin` / `.end` are, prophage_boundaries.py:45-48).  Contig ids are matched after the same clean-up
    as the FASTA headers (first word; `,` -> `___`).  Returns sorted 0-based half-open intervals."""
    path = Path(path)
    key = (str(path.resolve()), path.stat().st_mtime)
    if key in _TABLES:
        return _TABLES[key]
    opener = gzip.open if path.suffix == ".gz" else open
    kind = path.name[:-3] if path.suffix == ".gz" else path.name
    kind = kind.rsplit(".", 1)[-1].lower()
    genes: dict[str, Genes] = {}
    with opener(path, "rt") as fh:
        for line in fh:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            f = line.rstrip("\n").split("\t")
            if kind in ("gff", "gff3", "gtf"):
                if len(f) < 5 or f[2] not in ("CDS", "gene"):
                    continue
                contig, begin, end = f[0], int(f[3]) - 1, int(f[4])
            elif kind == "bed":
                contig, begin, end = f[0], int(f[1]), int(f[2])
            else:
                if len(f) < 3 or not f[1].lstrip("-").isdigit():
                    continue                    # a header line
                contig, begin, end = f[0], int(f[1]) - 1, int(f[2])
            if end < begin:
                begin, end = end, begin
            genes.setdefault(contig.split()[0].replace(",", "___"), []).append((begin, end))
    for v in genes.values():
        v.sort()
    _TABLES[key] = genes
    return genes


def pyrodigal_gene_finder():
    """`find_genes` of the reference (prophage_boundaries.py:33-49) when pyrodigal-gv is installed, else None."""
    try:
        import pyrodigal_gv                     # noqa: PLC0415  (optional third-party gene caller)
    except ImportError:
        return None
    finder = pyrodigal_gv.ViralGeneFinder(meta=True)

    def find(sequence: str) -> Genes:
        return sorted((int(g.begin) - 1, int(g.end)) for g in finder.find_genes(sequence))
    return find


def gene_source(gene_table: dict[str, Genes] | None, loaded) -> Callable[[str, int], Genes | None] | None:
    """`genes_of` for `refine_regions`: the `--genes` table when given, else pyrodigal-gv over the loaded FASTA when importable,
    else None (no refinement)."""
    if gene_table is not None:
        return lambda header, ci: gene_table.get(header.split()[0])
    find = pyrodigal_gene_finder()
    if find is None:
        return None
    names, host, offsets = loaded
    buf = host.numpy() if hasattr(host, "numpy") else np.asarray(host)
    return lambda header, ci: find(bytes(buf[int(offsets[ci]):int(offsets[ci + 1])]).decode())
