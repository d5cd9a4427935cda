"""Host-side I/O: fragment decode to columns, reference genomes, bigWig."""
from .alignment import AlignmentWrapper, Fragment
from .fragments import FragmentTable, as_table, load_fragments
from .reference import ReferenceWrapper

__all__ = ["AlignmentWrapper", "Fragment", "FragmentTable", "as_table", "load_fragments", "ReferenceWrapper"]
