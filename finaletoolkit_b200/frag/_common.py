"""Shared helpers of the feature modules: alias resolution, device access, text writers."""
from __future__ import annotations

import warnings

import numpy as np


def resolve_length_aliases(min_length, max_length, fraction_low, fraction_high, stacklevel=3):
    """Deprecated ``fraction_low/high`` aliases (frag/_wps.py:110-140, frag/_motif_common.py:97-138):
    an alias alone warns; alias + modern spelling raises ValueError."""
    if fraction_low is not None and min_length is None:
        min_length = fraction_low
        warnings.warn("fraction_low is deprecated. Use min_length instead.", category=DeprecationWarning,
                      stacklevel=stacklevel)
    elif fraction_low is not None and min_length is not None:
        warnings.warn("fraction_low is deprecated. Use min_length instead.", category=DeprecationWarning,
                      stacklevel=stacklevel)
        raise ValueError("fraction_low and min_length cannot both be specified")
    if fraction_high is not None and max_length is None:
        max_length = fraction_high
        warnings.warn("fraction_high is deprecated. Use max_length instead.", category=DeprecationWarning,
                      stacklevel=stacklevel)
    elif fraction_high is not None and max_length is not None:
        warnings.warn("fraction_high is deprecated. Use max_length instead.", category=DeprecationWarning,
                      stacklevel=stacklevel)
        raise ValueError("fraction_high and max_length cannot both be specified")
    return min_length, max_length


def dist_context():
    """``DistContext`` when ``torch.distributed`` is initialised with more than one rank (one process
    per GPU; contigs are then LPT-sharded over the ranks and genome-wide totals all-reduced), else None."""
    import sys
    if "torch.distributed" not in sys.modules and "torch" not in sys.modules:
        return None
    from ..sharding import DistContext
    ctx = DistContext()
    return ctx if ctx.on and ctx.world > 1 else None


def is_writer(ctx) -> bool:
    """Under a multi-rank launch only rank 0 writes output files / stdout."""
    return ctx is None or ctx.rank == 0


def group_by_contig(contigs):
    """{contig: [indices]} preserving first-appearance order."""
    groups: dict = {}
    for i, c in enumerate(contigs):
        groups.setdefault(c, []).append(i)
    return groups


def bedgraph_lines(contig: str, start: int, scores: np.ndarray) -> str:
    """``contig\\tpos\\tpos+1\\tscore\\n`` per position (frag/_multi_wps.py:328-341), vectorised."""
    n = scores.shape[0]
    if n == 0:
        return ""
    pos = np.arange(start, start + n, dtype=np.int64)
    a = np.char.add(contig + "\t", pos.astype(str))
    a = np.char.add(np.char.add(a, "\t"), (pos + 1).astype(str))
    a = np.char.add(np.char.add(a, "\t"), scores.astype(np.int64).astype(str))
    return "\n".join(a.tolist()) + "\n"
