"""Concurrent pinned host->device copy rate of the box: what bounds the e2e leg of bench.py at N ranks.

Every rank copies one step's input (150.7 MB, the size bench.py moves per step for 1024 stations) from pinned host memory to
its GPU `reps` times, all ranks starting together, with a device->host copy of one step's results (6 MB) running the other
way on a second stream.  Variants: default pinned memory / write-combined pinned memory; one contiguous copy / a pitched
1024-row 2-D copy (what sdrb_chain_process_host issues for a caller pitch that differs from the device pitch).

    python tools/h2d_concurrent.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/h2d_concurrent.py
Rank 0 prints one JSON line (append it to profiles/h2d_concurrent_r2.json).
"""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import sys

import torch
import torch.distributed as dist


def cudart():
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
        glob.glob("/usr/local/cuda/lib64/libcudart.so*")
    for p in cands:
        try:
            return C.CDLL(p)
        except OSError:
            pass
    raise SystemExit("libcudart not found")


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    S, pitch, bb, reps = 1024, 147200, 147000, 16
    n = S * pitch
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    back_d = torch.empty(6 << 20, dtype=torch.uint8, device=dev)
    back_h = torch.empty(6 << 20, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    res = {"ranks": world, "bytes_per_copy": n, "reps": reps}
    for mem, flags in (("default", 0), ("write_combined", 4)):
        hp = C.c_void_p()
        if rt.cudaHostAlloc(C.byref(hp), n, flags) != 0:
            res[mem] = "cudaHostAlloc failed"
            continue
        C.memset(hp, 128, n)
        for shape in ("contiguous", "pitched_2d"):
            def copy():
                if shape == "contiguous":
                    rt.cudaMemcpyAsync(d.data_ptr(), hp, n, 1, s1.cuda_stream)
                else:
                    rt.cudaMemcpy2DAsync(d.data_ptr(), pitch, hp, pitch, bb, S, 1, s1.cuda_stream)
            for _ in range(2):
                copy()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s1)
            for _ in range(reps):
                copy()
                with torch.cuda.stream(s2):
                    back_h.copy_(back_d, non_blocking=True)
            e1.record(s1)
            torch.cuda.synchronize()
            moved = reps * (n if shape == "contiguous" else S * bb)
            gbs = moved / (e0.elapsed_time(e1) * 1e-3) / 1e9
            t = torch.tensor([gbs, gbs], dtype=torch.float64, device=dev)
            if world > 1:
                lo = t.clone()
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                res[f"{mem}_{shape}_GBps"] = {"sum": round(float(t[0]), 2), "min_rank": round(float(lo[0]), 2), "per_rank_mean": round(float(t[0]) / world, 2)}
            else:
                res[f"{mem}_{shape}_GBps"] = {"sum": round(gbs, 2), "min_rank": round(gbs, 2), "per_rank_mean": round(gbs, 2)}
        rt.cudaFreeHost(hp)
    try:
        res["cpus"] = os.cpu_count()
        res["numa_nodes"] = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    except Exception:
        pass
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
