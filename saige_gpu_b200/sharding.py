"""Marker -> rank map of the multi-GPU path (host-side mirror of `owns()` in csrc/store.cu).

RAW (.bim-order) markers are dealt to ranks block-cyclically in blocks of SHARD_BLOCK consecutive markers; a rank stores the
QC-passing markers of its blocks.  Every rank therefore holds ~M_chr/world markers of EVERY chromosome and the
leave-one-chromosome-out products stay balanced (the reference's contiguous column slabs leave most ranks idle there:
gpuSymMatMult.cu:202-204), and both ingest passes of a rank touch the same 1/world of the .bed.  Local rows keep the global
order, hence a chromosome [start, end] is one contiguous local row range on every rank.  When every marker passes QC (the
synthetic workloads) raw index and QC index coincide and `local_markers(M, ...)` is the map itself."""
import numpy as np

SHARD_BLOCK = 1024          # must equal SGB_SHARD_BLOCK in csrc/sgb_internal.h


def owner_of_marker(gidx, world):
    return (np.asarray(gidx) // SHARD_BLOCK) % world


def local_markers(M, rank, world):
    """Global indices of the QC'd markers stored on `rank`, ascending."""
    g = np.arange(M)
    return g[owner_of_marker(g, world) == rank]


def local_markers_qc(qc_mask, rank, world):
    """Global QC indices stored on `rank` for a real marker set: qc_mask[m] = raw marker m passed QC."""
    raw = np.nonzero(np.asarray(qc_mask))[0]
    return np.nonzero(owner_of_marker(raw, world) == rank)[0]


def local_range(loc2glob, start, end):
    """Local row range [lo, hi) covering global markers start..end inclusive."""
    return int(np.searchsorted(loc2glob, start, "left")), int(np.searchsorted(loc2glob, end, "right"))
