"""Gap-track helpers used by the DELFI bin filter (mirror of finaletoolkit.genome)."""
from .gaps import ContigGaps, GenomeGaps

__all__ = ["GenomeGaps", "ContigGaps"]
