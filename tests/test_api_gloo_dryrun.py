"""World-size-2 CPU run (gloo) of the public API's multi-rank paths with the device layer swapped for the oracle
(tests/host_shim.py): contigs LPT-sharded over the ranks, one all-reduce per genome-wide result, rank 0 writes.
Input is the read-level BAM fixture, so the per-rank work also goes through ``per_fetch``.  Every rank must return
the full answer; it must equal the reference's golden text / the single-process result."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
repo = os.environ["FTK_REPO"]
sys.path[:0] = [repo, os.path.join(repo, "tests")]
from _pytest.monkeypatch import MonkeyPatch
import host_shim
import finaletoolkit_b200 as F
import finaletoolkit_b200.device as D
mp = MonkeyPatch()
seqs = {k: v.encode() for k, v in json.load(open(os.environ["FTK_SEQS"])).items()}
host_shim.install(mp, seqs)
mp.setattr(D, "require_cuda", lambda device=None: torch.device("cpu"))
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["FTK_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank = dist.get_rank()
d = os.environ["FTK_DIR"]
bam, tb, tiles = os.path.join(d, "read1.bam"), os.path.join(d, "read1.2bit"), os.path.join(d, "tiles.bed")
out = os.path.join(d, "cov_dist.bed")
res = F.coverage(bam, tiles, out, normalize=True, scale_factor=1e6, intersect_policy="any")
em = F.end_motifs(bam, tb, k=3, output_file=os.path.join(d, "em_dist.tsv"))
bp = F.breakpoint_motifs(bam, tb, k=4)
from finaletoolkit_b200.distributed import owned_contigs
from finaletoolkit_b200.io.fragments import load_fragments
print(json.dumps({"rank": rank, "mine": owned_contigs(load_fragments(bam)), "cov": [list(r) for r in res],
                  "em": [float(v) for _, v in em], "bp": [float(v) for _, v in bp]}))
dist.barrier(); dist.destroy_process_group()
'''


def test_gloo_world2_public_api(tmp_path, golden):
    import bam_replay as R
    from helpers import golden_codes
    fx = R.make_fixture(tmp_path, golden)
    seqs = {}
    for c, n in fx["m"]["refs"]:
        codes, nm = golden_codes(fx["g"], c, n)
        s = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
        s[nm] = ord("N")
        seqs[c] = s.tobytes().decode()
    (tmp_path / "seqs.json").write_text(json.dumps(seqs))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", FTK_REPO=REPO, FTK_PORT=port, FTK_DIR=str(tmp_path),
                   FTK_SEQS=str(tmp_path / "seqs.json"))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    got = [json.loads(o[0].strip().splitlines()[-1]) for o in outs]
    assert sorted(got[0]["mine"] + got[1]["mine"]) == ["chrA", "chrB"] and got[0]["mine"] and got[1]["mine"]   # really sharded
    # every rank holds the full answer, and it is the reference's (golden text of the same call on this BAM)
    case = next(c for c in fx["m"]["coverage"] if c["kwargs"].get("normalize"))
    text = "".join(f"{c}\t{s}\t{e}\t{n}\t{v}\n" for c, s, e, n, v in got[0]["cov"])
    assert text == case["text"] and got[0]["cov"] == got[1]["cov"]
    assert open(tmp_path / "cov_dist.bed").read() == case["text"]          # written once, by rank 0
    assert got[0]["em"] == got[1]["em"] and got[0]["bp"] == got[1]["bp"]
    # single-process result of the same calls (same shim)
    import host_shim
    from _pytest.monkeypatch import MonkeyPatch
    import finaletoolkit_b200 as F
    from finaletoolkit_b200.io import fragments
    mp = MonkeyPatch()
    try:
        fragments._CACHE.clear()
        host_shim.install(mp, {k: v.encode() for k, v in seqs.items()})
        em = [float(v) for _, v in F.end_motifs(fx["path"], fx["tb"], k=3)]
        bp = [float(v) for _, v in F.breakpoint_motifs(fx["path"], fx["tb"], k=4)]
    finally:
        mp.undo()
        fragments._CACHE.clear()
    assert got[0]["em"] == em and got[0]["bp"] == bp and abs(sum(em) - 1.0) < 1e-12
    assert open(tmp_path / "em_dist.tsv").read().count("\n") >= 64
