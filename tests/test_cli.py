"""CLI surface: subcommand names, flag <-> kwarg bijection (as reference tests/test_cli.py:41-78),
and one end-to-end stdout check (reference tests/test_cli.py:155-180) on the GPU."""
import importlib
import inspect
import os

import pytest
from click.testing import CliRunner

from helpers import write_text_gz


def test_help_and_kwarg_bijection():
    from finaletoolkit_b200.cli.main_cli import COMMANDS, main_cli
    r = CliRunner().invoke(main_cli, ["--help"])
    assert r.exit_code == 0
    expected = {"wps", "cleavage-profile", "adjust-wps", "coverage", "frag-length-bins", "frag-length-intervals", "end-motifs",
                "interval-end-motifs", "breakpoint-motifs", "interval-breakpoint-motifs", "delfi", "agg-bw", "mds", "regional-mds"}
    assert set(COMMANDS) == expected
    for name, (module, func, _, spec) in COMMANDS.items():
        assert CliRunner().invoke(main_cli, [name, "--help"]).exit_code == 0
        accepted = set(inspect.signature(getattr(importlib.import_module(module), func)).parameters)
        params = {names[-1] if kind == "opt" else names[0] for kind, names, _ in spec}
        params = (params - {"strand"}) | ({"both_strands", "negative_strand"} if "strand" in params else set())
        assert params <= accepted, (name, params - accepted)


def test_mds_commands(tmp_path, manifest):
    from finaletoolkit_b200.cli.main_cli import main_cli
    f = manifest["fixture17"]
    p = tmp_path / "dif.tsv"; p.write_text(f["end_motifs_dif_tsv"])
    import subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "finaletoolkit_b200.cli", "mds", str(p)], capture_output=True, text=True, cwd=repo)
    assert r.returncode == 0 and float(r.stdout) == f["mds_from_dif_tsv"]
    p2 = tmp_path / "ivl.tsv"; p2.write_text(f["end_motifs_intervals_dif_tsv"])
    out = tmp_path / "rmds.bed"
    assert CliRunner().invoke(main_cli, ["regional-mds", str(p2), str(out)]).exit_code == 0
    assert float(out.read_text().split("\t")[4]) == f["regional_mds"][0][1]


@pytest.mark.gpu
def test_coverage_stdout(tmp_path, manifest):
    from finaletoolkit_b200.cli.main_cli import main_cli
    m = manifest["fixture17"]
    frag = write_text_gz(tmp_path / "a.frag.gz", m["frag_gz_text"])
    ivl = tmp_path / "intervals.bed"; ivl.write_text(m["intervals_bed"])
    import subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "finaletoolkit_b200.cli", "coverage", frag, str(ivl), "-n"],
                       capture_output=True, text=True, cwd=repo)
    assert r.returncode == 0, r.stderr
    assert "12\t34443118\t34443538\t.\t0.25" in r.stdout and "12\t34444968\t34446115\t.\t0.4375" in r.stdout
    # wps with the default '-o -' is a ValueError in the reference too
    cs = tmp_path / "cs"; cs.write_text(m["chrom_sizes"])
    r = CliRunner().invoke(main_cli, ["wps", frag, str(ivl), "--chrom-sizes", str(cs)])
    assert isinstance(r.exception, ValueError)
