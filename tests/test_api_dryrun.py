"""CPU dry run of the public-API tests: the bodies of ``tests/test_gpu_api.py`` (reference fixtures and golden
texts through wps / multi_wps / coverage / frag_length* / end + breakpoint motifs / cleavage / delfi / BAM input) with the
kernel wrappers swapped for the oracle (``tests/host_shim.py``).  Covers the host side of those calls - argument
handling, warnings and errors, interval bookkeeping, statistics, text / bigWig writers - where no GPU exists; the
``-m gpu`` run of the same bodies covers the kernels."""
import numpy as np
import pytest

import host_shim
import test_gpu_api as A
from test_gpu_api import fx, syn  # noqa: F401  (fixtures)

_NEEDS_DEVICE = {"test_multi_wps_streams_large_contigs"}   # the streamed pipeline has no host stand-in


@pytest.fixture(autouse=True)
def _oracle_device_layer(monkeypatch):
    from finaletoolkit_b200.io import fragments
    from finaletoolkit_b200.io.reference import ReferenceWrapper

    def host_sequence(self, contig, device=None):
        codes, n_mask = self.contig_arrays(contig)
        s = np.frombuffer(b"ACGT", np.uint8)[np.asarray(codes) & 3].copy()
        s[np.asarray(n_mask).astype(bool)] = ord("N")
        return s.tobytes()

    fragments._CACHE.clear()
    host_shim.install(monkeypatch, {})
    monkeypatch.setattr(ReferenceWrapper, "device_contig", host_sequence)
    yield
    fragments._CACHE.clear()


for _name in dir(A):
    if _name.startswith("test_") and _name not in _NEEDS_DEVICE:
        globals()[_name] = getattr(A, _name)
del _name

# the API-level bodies of three more GPU test modules: the delfi() table, the cleavage drivers, agg_bw
from test_gpu_cleavage import test_golden_cases_through_the_api  # noqa: E402,F401
from test_gpu_delfi import test_delfi_api  # noqa: E402,F401
from test_gpu_agg import test_agg_bw_api_golden  # noqa: E402,F401
