"""N-GPU check of the distributed genome-wide reductions (NCCL).  Not collected by pytest:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py

Every rank computes the sharded result; rank 0 compares with the oracle's single-process stream.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from finaletoolkit_b200 import distributed as FD  # noqa: E402
from finaletoolkit_b200.io.fragments import FragmentTable  # noqa: E402
from finaletoolkit_b200.sharding import DistContext  # noqa: E402
from finaletoolkit_b200.synth import synth_fragments  # noqa: E402
from oracle import oracle as O  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = DistContext()
contigs = {"1": 900_000, "2": 700_000, "3": 650_000, "4": 300_000, "5": 120_000, "6": 90_000, "7": 30_000}
cols = {c: synth_fragments(n, n // 4, i, seed_base=6000) for i, (c, n) in enumerate(contigs.items())}
table = FragmentTable(cols)
mine = FD.owned_contigs(table, ctx)
total = FD.genome_total_coverage(table, quality_threshold=30, ctx=ctx, device=dev)
d = FD.genome_length_distribution(table, min_length=0, max_length=None, quality_threshold=30, ctx=ctx, device=dev)
if ctx.rank == 0:
    frs = {c: O.Frags(*cols[c]) for c in contigs}
    exp_total = sum(O.single_coverage(fr, 0, None) for fr in frs.values())
    exp = O.merge_dists(O.length_dist(fr, None, None, 0, None, "midpoint", 30) for fr in frs.values())
    assert total == exp_total, (total, exp_total)
    assert d == exp and list(d) == list(exp), "genome-wide dict / first-seen order mismatch"
    print(f"dist check ok: world={ctx.world} rank0 owns {mine}; total={total}, {len(d)} distinct lengths, "
          f"first keys {list(d)[:5]}")
dist.barrier()
dist.destroy_process_group()
