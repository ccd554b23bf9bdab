#!/usr/bin/env python
"""Headline benchmark: input MS/s decoded (stereo + RDS), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S] [--config r0|s0|m0|m2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): configs[4] of BASELINE.json — 1024 independent synthetic stereo+RDS stations at
2.4 MS/s each PER GPU, mode 0 type 'r'.  One step = one block (73 500 IQ pairs = 30.625 ms of signal) of every
station.  Stations shard over ranks with no data-path collective (weak scaling: 1024 stations per GPU); the
only torch.distributed traffic is the barrier and the max-over-ranks of the step time.

  value  whole-job MS/s with the uint8 IQ already resident in HBM (sdrb_chain_process_device), CUDA events on the
         launching stream, max over ranks.  16 distinct step inputs of 150 MB each are cycled, so every step reads
         input that is not in L2 (126 MB).
  e2e    the same metric through the host-facing call: sdrb_chain_process_host from pinned host memory (H2D inside
         the timed region) + sdrb_chain_read_pcm + sdrb_chain_read_rds (D2H inside), wall clock around a synchronise.
  roofline      dominant kernel of the step (by CUDA-event time measured here), algorithmic bytes / time vs measured HBM peak
  cpu_baseline  the reference's own CPU binary (oracle/_ref/project, built from the unmodified sources) timed on this
                box's host cores on a bounded sample of the same station; falls back to the oracle port if absent.

  roofline also carries, as scalars the driver's record keeps: bound ("latency": the step is one dependent chain per
         PLL sample), hbm_frac, pll_cycles_per_sample, fir_frac_no_fma / fir_frac_fma_peak (FIR MAC rate against the measured
         unfused multiply+add issue peak / the FMA peak), strong_* (BASELINE configs[4] as written: the 1024 stations SPLIT
         over the N ranks), capacity_* (4096 stations per GPU: the FIR-bound saturation rate, and capacity_fir_frac_no_fma: the
         FIR kernels' MAC rate at that batch, kernels serialised), sustained_* (a >= 5 s run).
  e2e    additionally link_gbs (a bare pinned host->device copy of the same size, all ranks at once, device->host running
         the other way) and frac_of_link.

--impl reference runs only that CPU arm (no GPU work) and prints the same JSON shape.
--config selects BASELINE.json configs[0..3] (m0 mono, m2 mono 147/800, s0 stereo, r0 stereo+RDS; --streams 1 for the
single-stream form, which also reports the block latency); the default is configs[4] = r0 with 1024 stations per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

UNIT = "MS/s"
N_STATIONS = 8       # distinct synthetic stations generated on the host
N_INPUTS = 16        # distinct step inputs resident on the device (each n_streams x 147 000 B)

# BASELINE.json configs: [4] is the headline (default), [0]..[3] are selected with --config m0|m2|s0|r0 (--streams 1 for
# the single-stream form the reference binary runs).  macs = algorithmic MACs per stream-block.
CONFIGS = {
    "r0": {"mode": 0, "kind": "r", "name": "stereo+RDS", "fs": "2.4", "baseline_config": 4,
           "macs": {"rf_frontend": 1_484_700, "if_bands": 2_227_050, "rds_carrier_bpf": 742_350, "audio": 296_940, "rds_backend": 572_872}},
    "s0": {"mode": 0, "kind": "s", "name": "stereo", "fs": "2.4", "baseline_config": 2,
           "macs": {"rf_frontend": 1_484_700, "if_bands": 1_484_700, "audio": 296_940}},
    "m0": {"mode": 0, "kind": "m", "name": "mono", "fs": "2.4", "baseline_config": 0,
           "macs": {"rf_frontend": 1_484_700, "audio": 148_470}},
    "m2": {"mode": 2, "kind": "m", "name": "mono, 147/800 resampler", "fs": "2.4", "baseline_config": 1,
           "macs": {"rf_frontend": 1_616_000, "audio": 148_470}},
}
CFG = CONFIGS["r0"]  # replaced by --config in main()


def metric_name():
    return f"input MS/s decoded ({CFG['name']})"


def workload(streams, bp):
    return (f"mode {CFG['mode']} {CFG['name']}, {streams} independent synthetic station(s) per GPU at {CFG['fs']} MS/s, "
            f"block = {bp} IQ pairs")


def config_dict(streams, bp, pitch):
    """`config` of the JSON line: the same for this arm and for --impl reference (which times the reference's CPU binary on a
    bounded sample of this workload, described in its cpu_baseline.sample)."""
    return {"workload": workload(streams, bp), "mode": CFG["mode"], "type": CFG["kind"], "streams_per_gpu": streams,
            "baseline_config": f"BASELINE.json configs[{CFG['baseline_config']}]",
            "l2": f"{N_INPUTS} distinct step inputs of {streams * pitch / 1e6:.1f} MB cycled"
                  + (" (each larger than the 126 MB L2)" if streams * pitch > 126e6 else " (their sum exceeds the 126 MB L2)" if N_INPUTS * streams * pitch > 126e6
                     else " and a 256 MB L2 flush buffer written between timed steps")}


def load_mod(name, rel):
    return entry._load(name, os.path.join(ROOT, rel))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation
# ------------------------------------------------------------------------------------------------
def _station_files(gen, nblocks, count, pad_blocks=3):
    """`count` distinct synthetic stations as raw IQ files in /dev/shm (pad blocks keep the reference's EOF race away from
    the measured part: the binary exits on EOF while its consumers may still hold 1-2 blocks)."""
    mode = CFG["mode"]
    bp = gen.block_pairs(mode)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    paths = []
    for k in range(count):
        iq = gen.generate_iq(gen.Station.for_stream(k, fs=gen.mode_fs(mode)), bp * nblocks)
        iq = np.concatenate([iq, np.tile(iq[: 2 * bp], pad_blocks)])
        path = os.path.join(d, f"sdrb_bench_{os.getpid()}_{k}.raw")
        iq.tofile(path)
        paths.append(path)
    return paths, bp * (nblocks + pad_blocks)


def cpu_reference_run(nblocks=98, steps=1, warmup=0, distinct=4):
    """The reference's CPU implementation of this config on the box's host cores.
    Returns dict(value MS/s aggregate, cores, kind, sample, ms_per_step)."""
    gen = load_mod("sdrgen", "real-time-sdr_b200/sdrgen.py")
    mode, kind = CFG["mode"], CFG["kind"]
    cores = os.cpu_count() or 1
    binary = os.path.join(ROOT, "oracle", "_ref", "project")
    if os.path.exists(binary) and os.access(binary, os.X_OK):
        inst = max(1, cores // 3)  # one instance = 3 threads (RF, audio, rds), src/project.cpp:134-136
        paths, pairs = _station_files(gen, nblocks, min(distinct, inst))

        def one_round():
            t0 = time.perf_counter()
            procs = [subprocess.Popen([binary, str(mode), kind], stdin=open(paths[i % len(paths)], "rb"), stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL) for i in range(inst)]
            for p in procs:
                p.wait()
            return time.perf_counter() - t0

        single = None
        try:
            for _ in range(warmup):
                one_round()
            times = [one_round() for _ in range(max(1, steps))]
            # the same reference functions on ONE thread (the function-level harness of oracle/): MS/s per core
            harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
            if os.path.exists(harness):
                out = paths[0] + ".rec"
                t0 = time.perf_counter()
                subprocess.run([harness, "chain", str(mode), kind, paths[0], out, "-1", "out"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                single = pairs / (time.perf_counter() - t0) / 1e6
                if os.path.exists(out):
                    os.unlink(out)
        finally:
            for p in paths:
                os.unlink(p)
        dt = statistics.median(times)
        return {"value": inst * pairs / dt / 1e6, "unit": UNIT, "cores": min(cores, 3 * inst), "kind": "reference",
                "single_thread": None if single is None else round(single, 2),
                "sample": f"{inst} concurrent instance(s) of oracle/_ref/project {mode} {kind} (the reference's own threaded binary, 3 threads "
                          f"each) on {len(paths)} distinct synthetic stations, {nblocks + 3} blocks ({pairs / gen.mode_fs(mode):.1f} s of signal) "
                          f"per instance, median of {len(times)} round(s); a bounded sample of the workload in `config`",
                "ms_per_step": dt * 1e3}
    # the reference binary was not built (no /root/reference at build time): the oracle port
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    orc = oracle_py.Oracle()
    bp = gen.block_pairs(mode)
    one = gen.generate_iq(gen.Station.for_stream(0, fs=gen.mode_fs(mode)), bp * nblocks)
    nstreams = cores
    iq = np.tile(one, nstreams)
    for _ in range(warmup):
        orc.run_batch(mode, kind, iq, nstreams, nblocks, cores)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        orc.run_batch(mode, kind, iq, nstreams, nblocks, cores)
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    return {"value": nstreams * bp * nblocks / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port, {nstreams} streams x {nblocks} blocks on {cores} threads, median of {len(times)}",
            "ms_per_step": dt * 1e3}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every few milliseconds by a thread while the timed region runs
    (NVML in-process; `nvidia-smi -lms` as the fallback: it needs ~100 ms to start, so it is started before warm-up)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, pci_bus_id=None, period_s=0.004):
        self.rows = []  # (t, sm_mhz, [reasons])
        self.max_mhz = None
        self.source = None
        self._stop = threading.Event()
        self.proc = None
        self.th = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if pci_bus_id:
                try:
                    h = nv.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                    (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        mask = int(get_reasons(h))
                        self.rows.append((time.perf_counter(), mhz, [n for b, n in bits if mask & b]))
                    except Exception:
                        pass
                    self._stop.wait(period_s)

            self.source = "nvml"
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            return
        except Exception:
            pass
        exe = shutil.which("nvidia-smi")
        if exe:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                mhz = float(c[0])
                self.max_mhz = max(self.max_mhz or 0.0, float(c[1]))
            except (ValueError, IndexError):
                continue
            self.rows.append((time.perf_counter(), mhz, [self.NAMES[i] for i in range(4) if len(c) >= 7 and c[3 + i].lower().startswith("active")]))

    def stop(self, t0, t1):
        """Summary of the samples taken inside [t0, t1] (perf_counter times of the timed region)."""
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi and NVML unavailable"], "samples": 0}
        self._stop.set()
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
        elif self.th:
            self.th.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        inside = len(rows)
        if not rows:  # a region shorter than one sampling period: the nearest samples either side
            rows = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:2]
        sm = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[2]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": inside, "source": self.source}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_inputs(torch, gen, n_streams, bb, pitch, dev, first_station=0, n_inputs=N_INPUTS):
    """n_inputs step inputs [n_streams][pitch] on the device.  Stream s plays station s % N_STATIONS, delayed by
    s // N_STATIONS blocks, so that no two streams of a step read the same bytes at the same time."""
    bp = bb // 2
    fs = gen.mode_fs(CFG["mode"])
    nst = min(N_STATIONS, max(1, n_streams))
    src = np.stack([gen.generate_iq(gen.Station.for_stream(k, fs=fs), bp * N_INPUTS).reshape(N_INPUTS, bb) for k in range(nst)])
    src_d = torch.from_numpy(src).to(dev)  # [station][block][bb]
    s = torch.arange(n_streams, device=dev) + first_station  # global station index (rank r owns a contiguous range)
    inputs = []
    for g in range(n_inputs):
        buf = torch.empty((n_streams, pitch), dtype=torch.uint8, device=dev)
        buf[:, :bb] = src_d[s % nst, (g + s // nst) % N_INPUTS]
        if pitch > bb:
            buf[:, bb:] = 128
        inputs.append(buf)
    return inputs


def alg_bytes_per_launch(S, info, kind):
    """Algorithmic bytes per launch of each kernel family = what it must read and write once."""
    bb, n_if, n_aud = info.block_bytes, info.if_block, info.audio_block
    nf = {"m": 0, "s": 2, "r": 3}[kind]
    t = {"rf_frontend": S * (bb + 4 * n_if)}
    if nf:
        t["if_bands"] = S * 4 * n_if * (1 + nf)
        t["pll"] = S * 4 * n_if * 2 * (2 if kind == "r" else 1)
        t["audio"] = S * (4 * n_if * 4 + 2 * 2 * n_aud)   # fm_demod, stereo band, NCO phase, (stereo_dc never leaves the SM) + PCM
    else:
        t["audio"] = S * (4 * n_if + 2 * n_aud)
    if kind == "r":
        t["rds_carrier_bpf"] = S * 4 * n_if * 2
        t["rds_backend"] = S * (4 * n_if * 2 + 128)          # rds band + NCO phase in, record out
    return t


def timed_steps(torch, dist, ch, inputs, pitch, stream, steps, first, world, flush=None):
    """steps blocks from device-resident inputs, bracketed by CUDA events on the launching stream; max over ranks (ms)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_in = len(inputs)
    e0.record(stream)
    for i in range(steps):
        if flush is not None:
            flush.add_(1)  # 256 MB written between timed steps: nothing of the previous step's input stays in L2
        ch.process_device(inputs[(first + i) % n_in].data_ptr(), pitch)
    ch.join()  # the launching stream waits for the internal streams: e1 closes the whole region
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=inputs[0].device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def side_run(torch, dist, capi, gen, S, world, local_rank, dev, stream, steps, first_station, inputs=None, n_inputs=4, kernel_blocks=0):
    """A second chain of S stations per rank (strong-scaling split / capacity point): whole-job MS/s and ms per step;
    with kernel_blocks > 0 also the per-kernel times of that batch, kernels serialised (median over kernel_blocks blocks)."""
    ch = capi.Chain(CFG["mode"], CFG["kind"], n_streams=S, device=local_rank)
    try:
        bb, bp = ch.info.block_bytes, ch.info.block_pairs
        pitch = (bb + 255) // 256 * 256
        ch.set_stream(stream.cuda_stream)
        ch.set_overlap(True)
        own = inputs is None
        if own:
            inputs = build_inputs(torch, gen, S, bb, pitch, dev, first_station=first_station, n_inputs=n_inputs)
        flush = None
        if S * pitch < 126e6:
            flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
        for i in range(6):
            ch.process_device(inputs[i % len(inputs)].data_ptr(), pitch)
        ch.join()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = timed_steps(torch, dist, ch, inputs, pitch, stream, steps, 6, world, flush)
        out = {"value": world * S * bp * steps / (ms * 1e-3) / 1e6, "ms_per_step": ms / steps, "streams_per_gpu": S}
        if kernel_blocks > 0:
            try:
                ch.join()
                torch.cuda.synchronize()
                ch.set_overlap(False)
                acc = {}
                for i in range(kernel_blocks):
                    ch.set_profiling(True)  # new window per block
                    ch.process_device(inputs[i % len(inputs)].data_ptr(), pitch)
                    for k, v in ch.kernel_times().items():
                        acc.setdefault(k, []).append(v)
                ch.set_profiling(False)
                out["kernel_ms_serialised"] = {k: statistics.median(v) for k, v in acc.items()}
            except Exception as e:  # the side figures never cost the main ones
                out["kernel_error"] = str(e)[:200]
        return out
    finally:
        ch.close()


def link_rate(torch, dist, dev, nbytes, world, reps=8):
    """Bare pinned host->device copy of one step's input size, every rank at once, with a device->host copy of the step's
    result size running the other way on a second stream: the ceiling of the e2e leg on this box (GB/s, this rank / sum)."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    back_d = torch.empty(6 << 20, dtype=torch.uint8, device=dev)
    back_h = torch.empty(6 << 20, dtype=torch.uint8).pin_memory()
    s2 = torch.cuda.Stream(device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            back_h.copy_(back_d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    total = gbs
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total = float(t.item())
    return gbs, total


def main():
    global CFG
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="stations per GPU")
    ap.add_argument("--config", default="r0", choices=sorted(CONFIGS), help="BASELINE.json config (default r0 = configs[4] with --streams 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling, capacity and sustained runs")
    ap.add_argument("--sustain-seconds", type=float, default=5.0)
    args = ap.parse_args()
    CFG = CONFIGS[args.config]
    mode, kind = CFG["mode"], CFG["kind"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line (rank 0): everything else that libraries print there (e.g. NCCL's version
    # banner) is sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    gen = load_mod("sdrgen", "real-time-sdr_b200/sdrgen.py")
    bp_cfg = gen.block_pairs(mode)
    pitch_cfg = (2 * bp_cfg + 255) // 256 * 256

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(nblocks=98, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": metric_name(), "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args.streams, bp_cfg, pitch_cfg),
                "cpu_baseline": {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                 "single_thread_ms_per_s": r.get("single_thread")},
                "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the receive chain has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = {}
    try:  # NUMA-local pinned buffers: bind to the CPUs next to this GPU before anything is allocated
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        numa = load_mod("sdrb_shard", "real-time-sdr_b200/shard.py").bind_process_to_gpu_node(bus)
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    entry.build()
    capi = load_mod("sdrb_capi", "real-time-sdr_b200/capi.py")

    S = args.streams
    ch = capi.Chain(mode, kind, n_streams=S, device=local_rank)
    bb, bp = ch.info.block_bytes, ch.info.block_pairs
    pitch = (bb + 255) // 256 * 256
    stream = torch.cuda.current_stream()
    ch.set_stream(stream.cuda_stream)
    ch.set_overlap(True)
    part = ch.sm_partition()  # (PLL SMs, FIR SMs) of the green-context split, (0, 0) if the driver refused it
    shard = load_mod("sdrb_shard", "real-time-sdr_b200/shard.py")
    mine = shard.station_range(rank, world, world * S)  # weak scaling: S stations per rank
    inputs = build_inputs(torch, gen, S, bb, pitch, dev, first_station=mine.start)
    flush = None
    if N_INPUTS * S * pitch < 126e6:  # a small batch would sit in L2: flush between timed steps (declared in config.l2)
        flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = None
    if rank == 0:  # started before the warm-up so that it is already sampling when the timed region begins
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            bus_id = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        except Exception:
            bus_id = None
        sampler = ClockSampler(local_rank, bus_id)
    for i in range(args.warmup):
        ch.process_device(inputs[i % N_INPUTS].data_ptr(), pitch)
    ch.join()
    barrier()
    launches0 = ch.launch_count()
    ch.set_profiling(True)  # one CUDA event pair per kernel launch, on the stream it runs on: averaged over the timed region
    t0 = time.perf_counter()
    ms = timed_steps(torch, dist, ch, inputs, pitch, stream, args.steps, args.warmup, world, flush)
    t1 = time.perf_counter()
    launches = ch.launch_count() - launches0
    kernel_ms_timed = ch.kernel_times()  # mean per launch over the timed region (overlap mode: kernels of neighbouring blocks co-run)
    ch.set_profiling(False)
    clocks = sampler.stop(t0, t1) if sampler else None
    value = world * S * bp * args.steps / (ms * 1e-3) / 1e6

    # ---- a sustained run (seconds, not milliseconds) with its own clock record
    sustained = None
    if not args.no_extras and args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds / max(ms / args.steps * 1e-3, 1e-6)))
        n_sus = min(n_sus, 200000)
        s2 = None
        if rank == 0:
            s2 = ClockSampler(local_rank, bus_id, period_s=0.02)
        barrier()
        ts0 = time.perf_counter()
        ms_sus = timed_steps(torch, dist, ch, inputs, pitch, stream, n_sus, 0, world, flush)
        ts1 = time.perf_counter()
        csus = s2.stop(ts0, ts1) if s2 else None
        sustained = {"value": round(world * S * bp * n_sus / (ms_sus * 1e-3) / 1e6, 1), "seconds": round(ms_sus * 1e-3, 2), "steps": n_sus,
                     "ms_per_step": round(ms_sus / n_sus, 4), "clocks": csus}

    # ---- per-kernel times (CUDA events around every kernel, serialised) -> dominant kernel for the roofline
    ch.set_overlap(False)
    ch.set_profiling(True)
    acc = {}
    nprof = 6
    for i in range(nprof):
        ch.set_profiling(True)  # new window per block: the median over blocks is reported
        ch.process_device(inputs[i % N_INPUTS].data_ptr(), pitch)
        for k, v in ch.kernel_times().items():
            acc.setdefault(k, []).append(v)
    ch.set_profiling(False)
    ch.set_overlap(True)
    kernel_ms = {k: statistics.median(v) for k, v in acc.items()}   # serialised (one kernel at a time): clean per-kernel numbers
    dom = max(kernel_ms_timed, key=kernel_ms_timed.get)              # dominant kernel of the timed region
    n_if = ch.info.if_block
    alg_bytes = alg_bytes_per_launch(S, ch.info, kind)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes.get(dom, 0) / (kernel_ms_timed[dom] * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/), scaled to this batch
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if args.config == "r0":
            traffic = int(t["dram_bytes_per_launch"][dom] * S / t["streams"])
    except Exception:
        pass
    sm_clock = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    # no-FMA MAC issue peak: tools/ubench.cu measures 60.5 MAC lanes/clk/SM for FMUL+FADD (and the same for the packed
    # FFMA2+FADD2 pair: packed instructions issue at half rate), i.e. half of the 128-lane FP32 pipe, x 148 SMs x max SM clock
    peak_no_fma = 60.5 * 148 * sm_clock / 1e12
    peak_fma = 128 * 148 * sm_clock / 1e12
    macs = CFG["macs"]
    fir_ms = sum(v for k, v in kernel_ms.items() if k in macs)
    mac_rate = S * sum(macs.values()) / (fir_ms * 1e-3) / 1e12 if fir_ms > 0 else 0.0
    pll_ms = kernel_ms.get("pll", 0.0)
    pll_cycles = pll_ms * 1e-3 / n_if * sm_clock
    latency_bound = dom == "pll"
    roofline = {"bound": "latency" if latency_bound else "fp32-issue", "kernel": dom, "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 5), "hbm_frac": round(achieved / hbm_peak, 5), "traffic": traffic,
                "algorithmic_bytes": alg_bytes.get(dom, 0),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                "bound_note": "achieved/peak/frac are the dominant kernel's algorithmic bytes per launch over its mean launch time against the "
                              "measured HBM peak, as the contract asks; the kernel is not HBM-bound: k_pll is one dependent chain per sample "
                              "(pll_cycles_per_sample x samples per block / SM clock = its duration whatever the batch), the FIR kernels are "
                              "bound by FP32 issue of the unfused multiply+add (fir_frac_no_fma)",
                "pll_cycles_per_sample": round(pll_cycles, 1), "pll_ms": round(pll_ms, 4),
                "fir_frac_no_fma": round(mac_rate / peak_no_fma, 4), "fir_frac_fma_peak": round(mac_rate / peak_fma, 4),
                "fir_tmacs_per_s": round(mac_rate, 3), "fir_ms_serialised": round(fir_ms, 4),
                "peak_tmacs_no_fma": round(peak_no_fma, 2), "peak_tmacs_fma": round(peak_fma, 2),
                "realtime_factor": round(value / (world * S * gen.mode_fs(mode) / 1e6), 2),
                "kernel_ms": {k: round(v, 4) for k, v in kernel_ms_timed.items()},
                "kernel_ms_serialised": {k: round(v, 4) for k, v in kernel_ms.items()},
                "fir_frac_no_fma_per_kernel": {k: round(S * m / (kernel_ms[k] * 1e-3) / 1e12 / peak_no_fma, 3) for k, m in macs.items() if kernel_ms.get(k, 0) > 0},
                "hbm_gbs_per_kernel": {k: round(alg_bytes[k] / (v * 1e-3) / 1e9, 1) for k, v in kernel_ms.items() if k in alg_bytes and v > 0},
                "timing": "kernel_ms: mean per launch over the timed region, CUDA events on each kernel's own stream (overlap mode, "
                          "kernels of neighbouring blocks run concurrently); kernel_ms_serialised: the same kernels one at a time"}
    if sustained:
        roofline.update({"sustained_value": sustained["value"], "sustained_seconds": sustained["seconds"],
                         "sustained_ms_per_step": sustained["ms_per_step"],
                         "sustained_sm_mhz": (sustained["clocks"] or {}).get("sm_mhz"), "sustained_reasons": (sustained["clocks"] or {}).get("reasons")})

    # ---- BASELINE configs[4] as written (strong scaling: the 1024 stations split over the ranks) and the capacity point
    if not args.no_extras and args.config == "r0" and S == 1024:
        if world > 1:
            Ss = max(1, S // world)
            r = side_run(torch, dist, capi, gen, Ss, world, local_rank, dev, stream, max(32, min(args.steps, 256)), rank * Ss,
                         inputs=[x[:Ss] for x in inputs])
            roofline.update({"strong_value": round(r["value"], 1), "strong_ms_per_step": round(r["ms_per_step"], 4),
                             "strong_streams_per_gpu": Ss, "strong_total_streams": Ss * world,
                             "strong_speedup_vs_one_gpu_expected": "~1.0: a step lasts as long as k_pll's chain (1 ms) whatever the batch"})
        else:
            roofline.update({"strong_value": round(value, 1), "strong_ms_per_step": round(ms / args.steps, 4),
                             "strong_streams_per_gpu": S, "strong_total_streams": S})
        try:
            cap = side_run(torch, dist, capi, gen, 4096, world, local_rank, dev, stream, 48, rank * 4096, n_inputs=4, kernel_blocks=4)
            roofline.update({"capacity_streams_per_gpu": 4096, "capacity_value": round(cap["value"], 1),
                             "capacity_ms_per_step": round(cap["ms_per_step"], 4),
                             "capacity_realtime_stations_per_gpu": int(cap["value"] / world / 2.4)})
            ck = cap.get("kernel_ms_serialised") or {}
            cap_fir_ms = sum(v for k, v in ck.items() if k in macs)
            if cap_fir_ms > 0:
                # the FIR kernels at the batch where they bound the step: the 1024-station figure above includes the partly filled
                # last wave of the short kernels (2.3 / 3.1 / 1.7 waves of CTAs for audio / carrier filter / RDS back end)
                roofline.update({"capacity_fir_frac_no_fma": round(4096 * sum(macs.values()) / (cap_fir_ms * 1e-3) / 1e12 / peak_no_fma, 4),
                                 "capacity_fir_ms_serialised": round(cap_fir_ms, 4),
                                 "capacity_kernel_ms_serialised": {k: round(v, 4) for k, v in ck.items()}})
        except Exception as e:  # never lose the headline line to the side run
            roofline["capacity_error"] = str(e)[:200]

    # ---- end to end through the host-facing call
    e2e = None
    if not args.no_e2e:
        n_host = 4
        pinned = capi.PinnedBuffer(n_host * S * pitch)
        hv = pinned.array.reshape(n_host, S, pitch)
        for g in range(n_host):
            hv[g] = inputs[g].cpu().numpy()
        pcm = capi.PinnedBuffer(S * ch.info.pcm_per_block * 2)
        pcm_v = pcm.array.view(np.int16).reshape(S, ch.info.pcm_per_block)
        k_e2e = max(4, min(args.steps, 96))  # 0.27 s at the PCIe rate: long enough for the link to reach its steady rate on every rank
        rec = np.zeros(S, capi.RDS_RECORD_DTYPE) if kind == "r" else None

        def e2e_steps(n):
            # software pipeline of depth 1 over the public calls: block i is issued, then the results of block i-1 are
            # read (lag 1) while block i is in flight; the last block's results are read with lag 0.
            ch.process_host_ptr(hv[0].ctypes.data, pitch)
            for i in range(1, n):
                ch.process_host_ptr(hv[i % n_host].ctypes.data, pitch)
                ch.read_results(1, pcm_v, rec)
            ch.read_results(0, pcm_v, rec)

        e2e_steps(3)
        barrier()
        t0e = time.perf_counter()
        e2e_steps(k_e2e)
        ch.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        d2h = S * (ch.info.pcm_per_block * 2 + (rec.dtype.itemsize if rec is not None else 0))
        e2e = {"value": round(world * S * bp * k_e2e / dt / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": S * bb,
               "d2h_bytes_per_step": d2h, "steps": k_e2e, "h2d_gbs_per_gpu": round(S * bb * k_e2e / dt / 1e9, 2),
               "call": "sdrb_chain_process_host + sdrb_chain_read_results (PCM + RDS records of every block, pinned host buffers)"}
        if S == 1:  # single-stream form (the reference binary's use): latency of one block through the public calls
            lat = []
            for i in range(12):
                tl = time.perf_counter()
                ch.process_host_ptr(hv[i % n_host].ctypes.data, pitch)
                ch.read_results(0, pcm_v, rec)
                lat.append((time.perf_counter() - tl) * 1e3)
            e2e["block_latency_ms"] = round(statistics.median(lat), 3)
            e2e["block_signal_ms"] = round(bp / gen.mode_fs(mode) * 1e3, 3)
        pinned.free()
        pcm.free()
        try:
            mine_gbs, total_gbs = link_rate(torch, dist, dev, max(S * pitch, 1 << 20), world)
            e2e.update({"link_gbs": round(mine_gbs, 2), "link_gbs_all_ranks": round(total_gbs, 2),
                        "frac_of_link": round((world * S * bb * k_e2e / dt / 1e9) / total_gbs, 3),
                        "link_note": "bare pinned host->device copy of one step's input size on every rank at once with a device->host "
                                     "copy running the other way (measured here, after the e2e leg)"})
        except Exception as e:
            e2e["link_error"] = str(e)[:200]

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        r = cpu_reference_run(nblocks=98, steps=3 if world == 1 else 1, warmup=0)
        cpu = {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "single_thread_ms_per_s": r.get("single_thread")}
    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {"metric": metric_name(), "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": dict(config_dict(S, bp, pitch), sm_partition={"pll_sms": part[0], "fir_sms": part[1]}),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "sustained": sustained, "host": {"numa_binding_rank0": numa, "cpus": os.cpu_count()}}
        emit(line)
    ch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
