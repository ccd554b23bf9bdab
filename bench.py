#!/usr/bin/env python
"""Headline benchmark: input MS/s decoded (stereo + RDS), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streams S]
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

--impl reference runs only that CPU arm (no GPU work) and prints the same JSON shape.
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

METRIC = "input MS/s decoded (stereo+RDS)"
UNIT = "MS/s"
MODE, KIND = 0, "r"
N_STATIONS = 8       # distinct synthetic stations generated on the host
N_INPUTS = 16        # distinct step inputs resident on the device (each n_streams x 147 000 B)
WORKLOAD = "mode 0 stereo+RDS, {s} independent synthetic stations per GPU at 2.4 MS/s, block = 73500 IQ pairs"

# algorithmic work per stream-block
MACS_PER_STREAM_BLOCK = 5_323_912


def load_mod(name, rel):
    return entry._load(name, os.path.join(ROOT, rel))


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation
# ------------------------------------------------------------------------------------------------
def _station_file(gen, nblocks, pad_blocks=3):
    """Synthetic station 0 as a raw IQ file in /dev/shm (pad blocks keep the reference's EOF race away from the
    measured part: the binary exits on EOF while its consumers may still hold 1-2 blocks)."""
    bp = gen.block_pairs(MODE)
    iq = gen.generate_iq(gen.Station(), bp * nblocks)
    iq = np.concatenate([iq, np.tile(iq[: 2 * bp], pad_blocks)])
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, f"sdrb_bench_{os.getpid()}.raw")
    iq.tofile(path)
    return path, bp * (nblocks + pad_blocks)


def cpu_reference_run(nblocks=98, steps=1, warmup=0):
    """Returns dict(value MS/s aggregate, cores, kind, sample, ms_per_step)."""
    gen = load_mod("sdrgen", "real-time-sdr_b200/sdrgen.py")
    cores = os.cpu_count() or 1
    binary = os.path.join(ROOT, "oracle", "_ref", "project")
    if os.path.exists(binary) and os.access(binary, os.X_OK):
        path, pairs = _station_file(gen, nblocks)
        inst = max(1, cores // 3)  # one instance = 3 threads (RF, audio, rds), src/project.cpp:134-136

        def one_round():
            t0 = time.perf_counter()
            procs = [subprocess.Popen([binary, "0", "r"], stdin=open(path, "rb"), stdout=subprocess.DEVNULL,
                                      stderr=subprocess.DEVNULL) for _ in range(inst)]
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
                out = path + ".rec"
                t0 = time.perf_counter()
                subprocess.run([harness, "chain", "0", "r", path, out, "-1", "out"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                single = pairs / (time.perf_counter() - t0) / 1e6
                if os.path.exists(out):
                    os.unlink(out)
        finally:
            os.unlink(path)
        dt = statistics.median(times)
        return {"value": inst * pairs / dt / 1e6, "unit": UNIT, "cores": min(cores, 3 * inst), "kind": "reference",
                "single_thread": None if single is None else round(single, 2),
                "sample": f"{inst} concurrent instance(s) of oracle/_ref/project 0 r (3 threads each), {nblocks + 3} blocks "
                          f"({pairs / 2.4e6:.1f} s of signal) of synthetic station 0 per instance, median of {len(times)}",
                "ms_per_step": dt * 1e3}
    # the reference binary was not built (no /root/reference at build time): the oracle port
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    orc = oracle_py.Oracle()
    bp = gen.block_pairs(MODE)
    one = gen.generate_iq(gen.Station(), bp * nblocks)
    nstreams = cores
    iq = np.tile(one, nstreams)
    for _ in range(warmup):
        orc.run_batch(MODE, KIND, iq, nstreams, nblocks, cores)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        orc.run_batch(MODE, KIND, iq, nstreams, nblocks, cores)
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
def build_inputs(torch, gen, n_streams, bb, pitch, dev, first_station=0):
    """N_INPUTS step inputs [n_streams][pitch] on the device.  Stream s plays station s % N_STATIONS, delayed by
    s // N_STATIONS blocks, so that no two streams of a step read the same bytes at the same time."""
    bp = bb // 2
    src = np.stack([gen.generate_iq(gen.Station.for_stream(k), bp * N_INPUTS).reshape(N_INPUTS, bb) for k in range(N_STATIONS)])
    src_d = torch.from_numpy(src).to(dev)  # [station][block][bb]
    s = torch.arange(n_streams, device=dev) + first_station  # global station index (rank r owns a contiguous range)
    inputs = []
    for g in range(N_INPUTS):
        buf = torch.empty((n_streams, pitch), dtype=torch.uint8, device=dev)
        buf[:, :bb] = src_d[s % N_STATIONS, (g + s // N_STATIONS) % N_INPUTS]
        if pitch > bb:
            buf[:, bb:] = 128
        inputs.append(buf)
    return inputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="stations per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
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

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(nblocks=98, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD.format(s=args.streams), "mode": MODE, "type": KIND},
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
    gen = load_mod("sdrgen", "real-time-sdr_b200/sdrgen.py")

    S = args.streams
    ch = capi.Chain(MODE, KIND, n_streams=S, device=local_rank)
    bb, bp = ch.info.block_bytes, ch.info.block_pairs
    pitch = (bb + 255) // 256 * 256
    stream = torch.cuda.current_stream()
    ch.set_stream(stream.cuda_stream)
    ch.set_overlap(True)
    shard = load_mod("sdrb_shard", "real-time-sdr_b200/shard.py")
    mine = shard.station_range(rank, world, world * S)  # weak scaling: S stations per rank
    inputs = build_inputs(torch, gen, S, bb, pitch, dev, first_station=mine.start)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, first):
        for i in range(n):
            ch.process_device(inputs[(first + i) % N_INPUTS].data_ptr(), pitch)

    # ---- device-resident throughput
    sampler = None
    if rank == 0:  # started before the warm-up so that it is already sampling when the timed region begins
        pr = torch.cuda.get_device_properties(local_rank)
        try:
            bus_id = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        except Exception:
            bus_id = None
        sampler = ClockSampler(local_rank, bus_id)
    run_steps(args.warmup, 0)
    ch.join()
    barrier()
    launches0 = ch.launch_count()
    ch.set_profiling(True)  # one CUDA event pair per kernel launch, on the stream it runs on: averaged over the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    run_steps(args.steps, args.warmup)
    ch.join()  # the launching stream waits for the internal streams: e1 closes the whole region
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = ch.launch_count() - launches0
    kernel_ms_timed = ch.kernel_times()  # mean per launch over the timed region (overlap mode: kernels of neighbouring blocks co-run)
    ch.set_profiling(False)
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * S * bp * args.steps / (ms * 1e-3) / 1e6

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
    n_if, n_rds, n_aud = ch.info.if_block, ch.info.rds_block, ch.info.audio_block
    # algorithmic bytes per launch of each kernel = what it must read and write once
    alg_bytes = {
        "rf_frontend": S * (bb + 4 * n_if),
        "if_bands": S * 4 * n_if * 4,
        "rds_carrier_bpf": S * 4 * n_if * 2,
        "pll": S * 4 * n_if * 4,
        "mix": S * 4 * n_if * 6,
        "audio": S * (4 * n_if * 2 + 2 * 2 * n_aud),
        "rds_backend": S * (4 * n_if + 128),
    }
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
        traffic = int(t["dram_bytes_per_launch"][dom] * S / t["streams"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "algorithmic_bytes": alg_bytes.get(dom, 0),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                "kernel_ms": {k: round(v, 4) for k, v in kernel_ms_timed.items()},
                "hbm_gbs_per_kernel": {k: round(alg_bytes[k] / (v * 1e-3) / 1e9, 1) for k, v in kernel_ms.items() if k in alg_bytes and v > 0},
                "kernel_ms_serialised": {k: round(v, 4) for k, v in kernel_ms.items()},
                "hbm_note": "hbm_gbs_per_kernel: algorithmic bytes / serialised kernel time; every kernel is far below the HBM peak, the FIR "
                            "kernels are graded against the no-FMA FP32 issue peak (fp32.per_kernel_frac), the slower of the two rooflines",
                "timing": "kernel_ms: mean per launch over the timed region, CUDA events on each kernel's own stream (overlap mode, "
                          "kernels of neighbouring blocks run concurrently); kernel_ms_serialised: the same kernels one at a time",
                "note": "the chain is FP32-issue / latency bound, not HBM bound (DESIGN.md section 5): see fp32"}
    fir_ms = sum(v for k, v in kernel_ms.items() if k not in ("pll", "mix"))
    mac_rate = S * MACS_PER_STREAM_BLOCK / (fir_ms * 1e-3) / 1e12
    # no-FMA MAC issue peak: tools/ubench.cu measures 60.5 MAC lanes/clk/SM for FMUL+FADD (and the same for the packed
    # FFMA2+FADD2 pair: packed instructions issue at half rate), i.e. half of the 128-lane FP32 pipe, x 148 SMs x max SM clock
    sm_clock = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    peak_no_fma = 60.5 * 148 * sm_clock / 1e12
    fp32 = {"fir_tmacs_per_s": round(mac_rate, 3), "peak_tmacs_per_s_no_fma": round(peak_no_fma, 2),
            "peak_note": "60.5 MAC lanes/clk/SM measured for the unfused multiply+add (profiles/ubench_r1.txt) x 148 SMs x max SM clock; "
                         "a bit-exact MAC is one multiply and one add, never an FMA, so this is half the FP32 FMA peak",
            "frac": round(mac_rate / peak_no_fma, 4), "fir_kernels_ms": round(fir_ms, 4),
            "per_kernel_frac": {k: round(S * m / (kernel_ms[k] * 1e-3) / 1e12 / peak_no_fma, 3)
                                for k, m in (("rf_frontend", 1_484_700), ("if_bands", 2_227_050), ("rds_carrier_bpf", 742_350),
                                             ("audio", 296_940), ("rds_backend", 572_872)) if k in kernel_ms},
            "pll_ns_per_step": round(kernel_ms.get("pll", 0) * 1e6 / n_if, 1),
            "pll_cycles_per_step": round(kernel_ms.get("pll", 0) * 1e-3 / n_if * sm_clock, 0),
            "pll_chain_floor_cycles": 300,
            "pll_note": "k_pll is bound by one dependent chain per sample (7350 sequential samples per block); the floor is that "
                        "chain's length with the measured instruction latencies (DESIGN.md section 5)"}

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
        k_e2e = max(4, min(args.steps, 24))
        rec = np.zeros(S, capi.RDS_RECORD_DTYPE)

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
        e2e = {"value": round(world * S * bp * k_e2e / dt / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": S * bb,
               "d2h_bytes_per_step": S * (ch.info.pcm_per_block * 2 + rec.dtype.itemsize), "steps": k_e2e,
               "call": "sdrb_chain_process_host + sdrb_chain_read_results (PCM + RDS records of every block, pinned host buffers)"}
        pinned.free()
        pcm.free()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(nblocks=98, steps=3, warmup=0)
        cpu = {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
               "single_thread_ms_per_s": r.get("single_thread")}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD.format(s=S), "mode": MODE, "type": KIND, "streams_per_gpu": S,
                           "l2": f"{N_INPUTS} distinct step inputs of {S * pitch / 1e6:.0f} MB cycled (each larger than the 126 MB L2)",
                           "realtime_factor": round(value / (world * S * 2.4), 2)},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "fp32": fp32, "cpu_baseline": cpu,
                "host": {"numa_binding_rank0": numa, "cpus": os.cpu_count()}}
        emit(line)
    ch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
