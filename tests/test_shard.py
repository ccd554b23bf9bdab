"""N > 1 path on CPU: two gloo ranks shard four stations, each decodes its own (with the oracle standing in for the
GPU chain, which is what the ranks would run on a B200), rank 0 gathers — result equals the single-process run."""
from __future__ import annotations

import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_station_range_partitions():
    from conftest import load_module
    sh = load_module("sdrb_shard", "real-time-sdr_b200/shard.py")
    for total in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                rg = sh.station_range(r, world, total)
                seen.extend(rg)
                for s in rg:
                    assert sh.owner_of(s, world, total) == r
            assert seen == list(range(total))
    assert [len(sh.station_range(r, 8, 1024)) for r in range(8)] == [128] * 8


WORKER = textwrap.dedent("""
    import os, sys, zlib
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests")); sys.path.insert(0, os.path.join({root!r}, "oracle"))
    from conftest import load_module
    import oracle_py
    sh = load_module("sdrb_shard", "real-time-sdr_b200/shard.py")
    gen = load_module("sdrgen", "real-time-sdr_b200/sdrgen.py")
    dist.init_process_group("gloo")
    rank, world, total, nblocks = dist.get_rank(), dist.get_world_size(), 4, 3
    orc = oracle_py.Oracle()
    rows = []
    for k in sh.station_range(rank, world, total):
        iq = gen.generate_iq(gen.Station.for_stream(k), gen.block_pairs(0) * nblocks)
        rows.append(orc.chain(0, "s", iq)["pcm"])
    local = np.stack(rows) if rows else np.zeros((0, 2940 * nblocks), np.int16)
    dist.barrier()
    allrows = sh.gather_rows(local, rank, world, total, dist)
    if rank == 0:
        np.save({out!r}, allrows)
    dist.destroy_process_group()
""")


def test_two_rank_gloo_gather(tmp_path, oracle, sdrgen):
    out = str(tmp_path / "gathered.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out)
    assert got.shape == (4, 2940 * 3)
    for k in range(4):
        iq = sdrgen.generate_iq(sdrgen.Station.for_stream(k), sdrgen.block_pairs(0) * 3)
        assert np.array_equal(got[k], oracle.chain(0, "s", iq)["pcm"]), k
