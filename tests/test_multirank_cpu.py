"""World-size-2 gloo test (CPU) of the row-sharding host logic: shard ranges, global row numbering, MIN reduction.
Each rank records its shard with the C++ front-end (host sink), the ORACLE evaluates it (no GPU here), and the
reduced first-unsatisfied row must equal the single-rank answer."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from bellpepper_b200 import fixtures, sharding
    from oracle import c_api
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    fid, blocks = 1, 5
    msg = fixtures.chain_message(blocks)
    b0, b1 = sharding.split_range(blocks, rank, world)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        digest, row_base = t.sha256(msg, b0, b1)
        lens, cols, coeffs, inputs, aux = t.host_csr()
    out = {}
    for name, victim in (("clean", None), ("early", 700), ("late", aux.shape[0] - 3000)):
        a = aux.copy()
        if victim is not None:
            a[victim] = [1 - int(a[victim][0]), 0, 0, 0]
        inst = c_api.Instance(fid, lens, cols, coeffs, inputs, a)
        res = torch.tensor([sharding.to_global(inst.check(2, False), row_base)], dtype=torch.int64)
        sharding.reduce_first_unsatisfied(res, world)
        out[name] = sharding.from_reduced(int(res.item()))
    rows = torch.tensor([lens.size // 3], dtype=torch.int64)
    dist.all_reduce(rows)
    out["rows_total"] = int(rows.item())
    out["row_base"] = row_base
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()
    """
) % ROOT


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import numpy as np

    from bellpepper_b200 import fixtures, sharding
    from oracle import c_api

    assert sharding.split_range(10, 0, 3) == (0, 3) and sharding.split_range(10, 2, 3) == (6, 10)
    assert sharding.to_global(-1, 5) == sharding.SATISFIED and sharding.to_global(7, 5) == 12
    # single-rank reference answers
    fid, blocks = 1, 5
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.sha256(fixtures.chain_message(blocks))
        lens, cols, coeffs, inputs, aux = t.host_csr()
    want = {"rows_total": lens.size // 3}
    for name, victim in (("clean", None), ("early", 700), ("late", aux.shape[0] - 3000)):
        a = aux.copy()
        if victim is not None:
            a[victim] = [1 - int(a[victim][0]), 0, 0, 0]
        r = c_api.Instance(fid, lens, cols, coeffs, inputs, a).check(2, False)
        want[name] = None if r < 0 else r
    assert want["clean"] is None and want["early"] is not None and want["late"] > want["early"]

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    import json

    got = json.loads(line[len("RESULT "):])
    for k in ("clean", "early", "late", "rows_total"):
        assert got[k] == want[k], (k, got, want)
    assert got["row_base"] == 0


def test_split_units_by_rows_balances_rows():
    """The sha256 chain's shards: contiguous, exhaustive, and balanced in ROWS although rank 0 also owns the input-bit rows."""
    from bellpepper_b200.sharding import split_units_by_rows

    for blocks, world in ((4096, 8), (4096, 2), (170, 3), (5, 8), (64, 1)):
        rpb, lead = 26192, 8 * (64 * blocks - 9)
        cuts = [split_units_by_rows(blocks, rpb, lead, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == blocks
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))  # contiguous, nothing lost
        rows = [(b1 - b0) * rpb + (lead if r == 0 else 0) for r, (b0, b1) in enumerate(cuts)]
        assert sum(rows) == lead + blocks * rpb
        if blocks >= 8 * world:
            assert max(rows) - min(rows) <= rpb  # within one block of each other
