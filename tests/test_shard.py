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


GPU_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from conftest import load_module
    sh = load_module("sdrb_shard", "real-time-sdr_b200/shard.py")
    gen = load_module("sdrgen", "real-time-sdr_b200/sdrgen.py")
    capi = load_module("sdrb_capi", "real-time-sdr_b200/capi.py")
    ndev = torch.cuda.device_count()
    backend = "nccl" if ndev >= 2 else "gloo"   # two ranks cannot share one GPU under NCCL
    dist.init_process_group(backend)
    rank, world, total, nblocks = dist.get_rank(), dist.get_world_size(), 5, 4
    device = rank % ndev
    torch.cuda.set_device(device)
    mine = sh.station_range(rank, world, total)
    bb = 2 * gen.block_pairs(0)
    iq = np.stack([gen.generate_iq(gen.Station.for_stream(k), gen.block_pairs(0) * nblocks) for k in mine])
    pcm, groups = [], []
    with capi.Chain(0, "r", n_streams=len(mine), device=device) as ch:
        ch.set_overlap(True)
        for b in range(nblocks):
            ch.process_host(np.ascontiguousarray(iq[:, b * bb:(b + 1) * bb]))
            pcm.append(ch.read_pcm().copy())
            groups.append(ch.read_rds()["n_bits"].astype(np.int32).copy())
    local = np.concatenate(pcm, axis=1)
    bits = np.stack(groups, axis=1)
    dist.barrier()
    allpcm = sh.gather_rows(local, rank, world, total, dist)
    allbits = sh.gather_rows(bits, rank, world, total, dist)
    if rank == 0:
        np.savez({out!r}, pcm=allpcm, n_bits=allbits, backend=backend)
    dist.destroy_process_group()
""")


import pytest  # noqa: E402


@pytest.mark.gpu
def test_two_rank_cuda_chains_match_single_process(tmp_path, oracle, sdrgen):
    """The N > 1 product path with a result check: two ranks, each its own CUDA chain over its own stations (3 + 2 of 5),
    rows gathered on rank 0 (NCCL when the box has two GPUs, gloo with both ranks on the one GPU otherwise): equal to the
    oracle's single-process run."""
    out = str(tmp_path / "gathered.npz")
    script = tmp_path / "worker_gpu.py"
    script.write_text(GPU_WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = np.load(out)
    assert got["pcm"].shape == (5, 2940 * 4)
    for k in range(5):
        iq = sdrgen.generate_iq(sdrgen.Station.for_stream(k), sdrgen.block_pairs(0) * 4)
        want = oracle.chain(0, "r", iq)
        assert np.array_equal(got["pcm"][k], want["pcm"]), k
        assert int(got["n_bits"][k].sum()) == want["rds_bits"].size, k
