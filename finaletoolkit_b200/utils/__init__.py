"""Host utilities mirroring finaletoolkit.utils for the hot path (reference utils/__init__.py:8-57)."""
from .utils import (chrom_sizes_to_dict, chrom_sizes_to_list, frag_array, frag_generator, frags_in_region,
                    gen_kmers, get_intervals, overlaps, reverse_complement, _none_eq, _none_geq, _none_leq)

from ._agg_bw import agg_bw

__all__ = ["agg_bw", "chrom_sizes_to_dict", "chrom_sizes_to_list", "frag_array", "frag_generator", "frags_in_region",
           "gen_kmers", "get_intervals", "overlaps", "reverse_complement", "_none_eq", "_none_geq", "_none_leq"]
