"""Centromere / telomere / short-arm intervals - API mirror of reference genome/gaps.py.

Only what the DELFI window path consumes is here: ``GenomeGaps`` built from a BED4 gap file
(contig, start, stop, type in {centromere, telomere, short_arm}), its per-contig view
``ContigGaps`` and the arm / in_tcmere predicates (genome/gaps.py:92-268).  The reference also
ships the UCSC hg19 / hg38 gap tracks as package data; those files are not redistributed here,
so ``GenomeGaps.ucsc_hg19()`` / ``b37()`` / ``hg38()`` ask for a BED4 made with the reference's
``finaletoolkit gap-bed`` instead.
"""
from __future__ import annotations

import gzip
from sys import stdout

import numpy as np

from ..exceptions import UnsupportedFormatError

__all__ = ["GenomeGaps", "ContigGaps"]

_GAP_DTYPE = [("contig", "<U32"), ("start", "<i8"), ("stop", "<i8"), ("type", "<U32")]


def _overlaps_any(start, stop, rows) -> bool:
    return bool(np.any((stop > rows["start"]) & (start < rows["stop"])))


class GenomeGaps:
    """All gaps of a genome, from a BED4 file (genome/gaps.py:40-62)."""

    def __init__(self, gaps_bed=None) -> None:
        if gaps_bed is None:
            return
        self._set_gaps(np.atleast_1d(np.genfromtxt(gaps_bed, dtype=_GAP_DTYPE)))

    def _set_gaps(self, gaps) -> None:
        self.gaps = gaps
        self.centromeres = gaps[gaps["type"] == "centromere"]
        self.telomeres = gaps[gaps["type"] == "telomere"]
        self.short_arms = gaps[gaps["type"] == "short_arm"]

    @classmethod
    def _bundled(cls, name):
        raise UnsupportedFormatError(
            f"The UCSC {name} gap track is package data of the reference and is not bundled here; "
            "pass a BED4 gap file (e.g. written by `finaletoolkit gap-bed`).")

    @classmethod
    def ucsc_hg19(cls):
        return cls._bundled("hg19")

    @classmethod
    def b37(cls):
        return cls._bundled("b37")

    @classmethod
    def hg38(cls):
        return cls._bundled("hg38")

    def _of(self, rows, contig):
        return rows[rows["contig"] == contig]

    def in_tcmere(self, contig, start, stop):
        """True when the interval overlaps a centromere or any telomere; None without a centromere
        (genome/gaps.py:92-133)."""
        cen = self._of(self.centromeres, contig)
        if not cen.shape[0]:
            return None
        tel = self._of(self.telomeres, contig)
        return _overlaps_any(start, stop, cen) or (bool(tel.shape[0]) and _overlaps_any(start, stop, tel))

    def overlaps_gap(self, contig, start, stop):
        """genome/gaps.py:135-143."""
        rows = self._of(self.gaps, contig)
        if not rows.shape[0]:
            return None
        return _overlaps_any(start, stop, rows)

    def get_arm(self, contig, start, stop) -> str:
        """genome/gaps.py:145-168."""
        if stop < start:
            raise ValueError("start must be less than stop")
        cen = self._of(self.centromeres, contig)
        has_short_arm = self._of(self.short_arms, contig).shape[0] > 0
        name = contig.replace("chr", "")
        if stop < cen["start"][0]:
            return "NOARM" if has_short_arm else f"{name}p"
        if start > cen["stop"][0]:
            return f"{name}q"
        return "NOARM"

    def get_contig_gaps(self, contig):
        """Per-contig view, or None when the contig has no centromere (genome/gaps.py:170-182)."""
        cen = self._of(self.centromeres, contig)
        if not cen.shape[0]:
            return None
        tel = self._of(self.telomeres, contig)
        return ContigGaps(contig, (cen[0]["start"], cen[0]["stop"]), [(t["start"], t["stop"]) for t in tel],
                          self._of(self.short_arms, contig).shape[0] > 0)

    def to_bed(self, output_file) -> None:
        """Sorted BED4, name = gap type (genome/gaps.py:184-207)."""
        text = "".join(f"{g['contig']}\t{g['start']}\t{g['stop']}\t{g['type']}\n" for g in np.sort(self.gaps))
        if str(output_file) == "-":
            stdout.write(text)
        elif str(output_file).endswith(".gz"):
            with gzip.open(output_file, "wt") as fh:
                fh.write(text)
        else:
            with open(output_file, "w") as fh:
                fh.write(text)


class ContigGaps:
    """Centromere / telomeres of one contig (genome/gaps.py:210-268)."""

    def __init__(self, contig, centromere, telomeres, has_short_arm=False) -> None:
        self.contig = contig
        self.centromere = centromere
        self.telomeres = list(telomeres)
        self.has_short_arm = has_short_arm

    def in_tcmere(self, start, stop) -> bool:
        """Overlaps the centromere, or overlaps EVERY telomere - the reference's ``all`` is kept on
        purpose (genome/gaps.py:226-248); the CUDA kernel applies the same rule per fragment."""
        in_centromere = stop > self.centromere[0] and start < self.centromere[1]
        in_telomeres = bool(self.telomeres) and all(stop > a and start < b for a, b in self.telomeres)
        return bool(in_centromere or in_telomeres)

    def in_gap(self, start, stop) -> bool:
        """genome/gaps.py:250-259: like in_tcmere, but an empty telomere list counts as inside."""
        in_centromere = stop > self.centromere[0] and start < self.centromere[1]
        return bool(in_centromere or all(stop > a and start < b for a, b in self.telomeres))

    def get_arm(self, start, stop) -> str:
        if stop < start:
            raise ValueError("start must be less than stop")
        name = self.contig.replace("chr", "")
        if stop < self.centromere[0]:
            return "NOARM" if self.has_short_arm else f"{name}p"
        if start > self.centromere[1]:
            return f"{name}q"
        return "NOARM"
