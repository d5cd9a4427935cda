"""Exception hierarchy mirroring the reference's (exceptions.py:23-64): every
error subclasses the builtin the original raised so existing handlers keep working."""
from __future__ import annotations


class FinaleToolkitError(Exception):
    """Base class for all toolkit-specific errors."""


class InvalidInputError(FinaleToolkitError, ValueError):
    """Malformed or inconsistent user input."""


class UnsupportedFormatError(InvalidInputError):
    """Input file format the toolkit cannot read."""


class MissingReferenceError(InvalidInputError):
    """A reference genome is required but was not given."""


class MissingIndexError(FinaleToolkitError, FileNotFoundError):
    """A required index (.tbi/.bai/.crai/.fai) is missing."""


class ContigNotFoundError(InvalidInputError):
    """Requested contig absent from a reference or alignment."""


class ContigMismatchError(InvalidInputError):
    """Contigs of two files are incompatible."""


class OutOfBoundsError(InvalidInputError, IndexError):
    """Queried interval falls outside chromosome bounds."""


class BigWigWriteError(FinaleToolkitError, OSError):
    """A background compression / write batch of the bigWig writer failed.  Deliberately NOT a
    RuntimeError: the API mirrors catch RuntimeError to skip out-of-order intervals like the
    reference does (frag/_multi_wps.py:319-325) and must not swallow real I/O failures."""


__all__ = [n for n in dir() if n.endswith("Error")]
