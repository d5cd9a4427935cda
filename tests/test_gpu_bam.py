"""BAM input through the CUDA path: the READ-level region selection of the reference
(io/alignment.py:242-247) reproduced for every per-region feature (bodies in tests/bam_replay.py)."""
import pytest

import bam_replay as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bam(tmp_path_factory, golden):
    from finaletoolkit_b200.io import fragments
    fragments._CACHE.clear()
    yield R.make_fixture(tmp_path_factory.mktemp("gpu_bam"), golden)
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
