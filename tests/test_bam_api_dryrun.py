"""CPU dry run of the BAM replay: the public API end to end (decoder, read-level fetch grouping, per-feature
drivers, writers) with the kernel wrappers swapped for the oracle (tests/host_shim.py).  What the GPU run of the
same bodies (tests/test_gpu_bam.py) adds is the kernels themselves."""
import numpy as np
import pytest

import bam_replay as R
import host_shim


@pytest.fixture()
def bam(tmp_path, golden, monkeypatch):
    from finaletoolkit_b200.io import fragments
    from helpers import golden_codes
    fragments._CACHE.clear()
    fx = R.make_fixture(tmp_path, golden)
    seqs = {}
    for c, n in fx["m"]["refs"]:
        codes, nm = golden_codes(fx["g"], c, n)
        s = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
        s[nm] = ord("N")
        seqs[c] = s.tobytes()
    host_shim.install(monkeypatch, seqs)
    yield fx
    fragments._CACHE.clear()


def test_table_carries_read1(bam):
    R.check_table_carries_read1(bam)

def test_wps_and_multi_wps(bam):
    R.check_wps_and_multi_wps(bam)

def test_coverage(bam):
    R.check_coverage(bam)

def test_fragment_lengths(bam):
    R.check_fragment_lengths(bam)

def test_motifs(bam):
    R.check_motifs(bam)

def test_cleavage(bam):
    R.check_cleavage(bam)

def test_delfi_bins(bam):
    R.check_delfi_bins(bam)
