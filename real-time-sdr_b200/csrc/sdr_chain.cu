// libsdr_b200: host orchestration and C ABI (include/sdr_b200.h) of the B200 receive chain.
//
// One sdrb_chain owns the carried state of n_streams independent stations on one GPU and turns one
// block of 8-bit IQ per station into PCM audio and an RDS record per call, doing what the reference's
// three thread bodies do per block (/root/reference/src/rffrontend.cpp:45-76, src/mono.cpp:29-49,
// src/stereo.cpp:69-114, src/rds.cpp:95-192).  There is no CPU fallback: without a usable GPU every
// entry point that needs one fails with SDRB_ERR_NO_DEVICE / SDRB_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sdr_b200.h"
#include "sdr_kernels.cuh"

using namespace sdrb;

static_assert(sizeof(RdsRecord) == sizeof(sdrb_rds_record), "RdsRecord must mirror sdrb_rds_record");
static_assert(sizeof(PllStateAbi) == sizeof(sdrb_pll_state), "PllStateAbi must mirror sdrb_pll_state");

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SDRB_ERR_NO_DEVICE : SDRB_ERR_CUDA;
}
#define CU(x)                                             \
    do {                                                  \
        cudaError_t e_ = (x);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #x);  \
    } while (0)

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

struct Ring {
    float* base = nullptr;
    size_t pitch = 0, slot = 0;
    int halo = 0, n = 0;
    RingView view(long long b) const {
        RingView v;
        v.cur = base + (size_t)(b % kNRing) * slot + halo;
        v.nxt = base + (size_t)((b + 1) % kNRing) * slot + halo;
        v.pitch = pitch;
        v.halo = halo;
        v.n = n;
        return v;
    }
    const float* cur(long long b) const { return base + (size_t)(b % kNRing) * slot + halo; }
};

// Per-kernel-family device timing: a ring of CUDA event pairs, one pair per launch, recorded on the stream the kernel
// runs on.  sdrb_chain_kernel_times reports the mean over the launches since profiling was switched on.
constexpr int kTimedSlots = 128;
struct Timed {
    const char* name;
    std::vector<cudaEvent_t> e0, e1;
    long long n;  // launches recorded since profiling was enabled
};

}  // namespace

struct sdrb_chain {
    sdrb_config cfg;
    sdrb_chain_info info;
    int S = 0;
    int up = 1, down = 5;
    bool stereo = false, rds = false;
    bool rds_gated = false;  // type 'r' in a mode without an RDS back end: records are reported as gated
    int poisoned = 0;        // a launch failed in the middle of a block: the carried state is inconsistent (SDRB_ERR_STATE from then on)
    long long blocks_since_load = 0;  // blocks processed since creation / the last state load: results exist only for those
    int pll_max_ctas = kPllMaxCtas;  // SMs given to k_pll (tuning knob: environment variable SDRB_PLL_MAX_CTAS)
    long long block = 0;  // index of the next block to process
    long long launches = 0;
    cudaStream_t stream = nullptr;  // the caller-visible stream: inputs are ordered on it, joins land on it
    bool own_stream = false;
    // overlap mode: front end of block b+1 concurrent with PLL / back end of block b (see process_block)
    cudaStream_t s_front = nullptr, s_pll = nullptr, s_back = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in = nullptr, ev_front[kNRing] = {}, ev_pll[kNRing] = {}, ev_back[kNRing] = {}, ev_h2d[2] = {}, ev_consumed[2] = {}, ev_join = nullptr;
    // SM partition (green contexts): the PLL stream owns pll_sms SMs, the FIR streams the rest; 0/0 when not in use
    void* gctx_pll = nullptr;
    void* gctx_fir = nullptr;
    int part_pll_sms = 0, part_fir_sms = 0;
    bool input_is_host = false;  // the block being issued comes from process_host (its input is consumed by the H2D copy)
    bool pending = false;  // work issued on the internal streams that the main stream has not been joined with
    // taps
    Taps101 rf_h, pilot_h, stereo_h, rds_h, rds114_h, rrc_h, audio_h;
    float* d_audio_pm = nullptr;   // phase-major audio taps (up > 1)
    float* d_audio_tm = nullptr;   // thread-major audio taps of the phase-class resampler (k_audio_updown_pc), nullptr: not applicable
    int* d_audio_res = nullptr;    // output residue per thread of that kernel
    float2* d_rds_res_taps = nullptr;  // RDS low-pass taps as the resampler lanes read them, [kResIter][kResLanes]
    int4* d_rds_res_lane = nullptr;    // [kResLanes] the lanes (ResLane, sdr_kernels.cuh)
    // input
    uint8_t* d_iq[2] = {nullptr, nullptr};  // staging for process_host
    size_t iq_pitch = 0;
    uint8_t* d_iq_halo[2] = {nullptr, nullptr};
    // rings
    Ring fm, pilot, sband, rband, gpilot, trig19, trig114, sdc, rdc;
    // states
    PllStateDev* d_pll[2] = {nullptr, nullptr};
    float* d_filt_state[2] = {nullptr, nullptr};
    RdsStreamState* d_rds_state = nullptr;
    RdsRecord* d_rec[2] = {nullptr, nullptr};
    unsigned int* d_rds_overflow = nullptr;  // [3], see RdsArgs::overflow
    unsigned long long* d_pll_redo = nullptr;  // [2], see PllLoop::redo
    // outputs, double buffered by block parity (a lagged read of block b-1 may overlap block b)
    int16_t* d_pcm[2] = {nullptr, nullptr};
    size_t pcm_pitch = 0;
    // optional stage dumps
    float *d_ids = nullptr, *d_qds = nullptr, *d_carrier = nullptr, *d_ipll = nullptr, *d_rdelay = nullptr,
          *d_mono = nullptr, *d_sfilt = nullptr, *d_rfilt = nullptr, *d_rclean = nullptr;
    // profiling
    bool profiling = false;
    bool overlap = false;
    std::vector<Timed> timed;
    std::vector<void*> allocs;
    bool guard = false;                                // SDRB_GUARD=1: canary zones around every allocation
    struct GuardedAlloc { char* base; size_t bytes; };
    std::vector<GuardedAlloc> guards;
};

namespace {

// ---- SM partition -------------------------------------------------------------------------------------------------
// k_pll is one dependent chain per sample: its warps must never wait for an SM or share one.  Stream priority plus a
// whole-SM shared-memory reservation (round 1) keeps FIR CTAs off the SMs a running PLL kernel holds, but not off the
// SMs it frees between two launches: the next launch then waits until those CTAs have drained (measured 0.03-0.11 ms per
// step, profiles/README.md).  Green contexts (driver API, CUDA 12.4+) split the device for good: the PLL stream is
// created in a context that owns `pll_sms` SMs, the front- and back-end streams in one that owns the rest.  The driver
// entry points are looked up at run time (no link-time dependency on libcuda: the library must load on a CPU-only box);
// if anything is missing or refused the chain keeps its ordinary priority streams.
struct GreenApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};
const GreenApi& green_api() {
    static GreenApi api = [] {
        GreenApi a;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        a.ok = get("cuDeviceGet", (void**)&a.DeviceGet) && get("cuDeviceGetDevResource", (void**)&a.DeviceGetDevResource) &&
               get("cuDevSmResourceSplitByCount", (void**)&a.DevSmResourceSplitByCount) &&
               get("cuDevResourceGenerateDesc", (void**)&a.DevResourceGenerateDesc) && get("cuGreenCtxCreate", (void**)&a.GreenCtxCreate) &&
               get("cuGreenCtxDestroy", (void**)&a.GreenCtxDestroy) && get("cuGreenCtxStreamCreate", (void**)&a.GreenCtxStreamCreate);
        return a;
    }();
    return api;
}
// Creates the two contexts and the three compute streams in them; false (nothing created, nothing leaked) if unavailable.
bool make_sm_partition(sdrb_chain* c, int pll_sms, int prio_lo, int prio_hi) {
    const GreenApi& g = green_api();
    if (!g.ok) return false;
    CUdevice dev;
    CUdevResource all, grp, rest;
    unsigned int ngrp = 1;
    if (g.DeviceGet(&dev, c->cfg.device) != CUDA_SUCCESS) return false;
    if (g.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    if ((int)all.sm.smCount < pll_sms + 8) return false;
    if (g.DevSmResourceSplitByCount(&grp, &ngrp, &all, &rest, 0, (unsigned)pll_sms) != CUDA_SUCCESS || ngrp != 1) return false;
    if (grp.sm.smCount < (unsigned)pll_sms || rest.sm.smCount == 0) return false;
    CUdevResourceDesc d_pll, d_fir;
    if (g.DevResourceGenerateDesc(&d_pll, &grp, 1) != CUDA_SUCCESS || g.DevResourceGenerateDesc(&d_fir, &rest, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx gp = nullptr, gf = nullptr;
    if (g.GreenCtxCreate(&gp, d_pll, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (g.GreenCtxCreate(&gf, d_fir, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        g.GreenCtxDestroy(gp);
        return false;
    }
    CUstream sp = nullptr, sf = nullptr, sb = nullptr;
    const bool ok = g.GreenCtxStreamCreate(&sp, gp, CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS &&
                    g.GreenCtxStreamCreate(&sf, gf, CU_STREAM_NON_BLOCKING, prio_lo) == CUDA_SUCCESS &&
                    g.GreenCtxStreamCreate(&sb, gf, CU_STREAM_NON_BLOCKING, prio_lo) == CUDA_SUCCESS;
    if (!ok) {
        for (CUstream st : {sp, sf, sb})
            if (st) cudaStreamDestroy((cudaStream_t)st);
        g.GreenCtxDestroy(gp);
        g.GreenCtxDestroy(gf);
        return false;
    }
    c->s_pll = (cudaStream_t)sp;
    c->s_front = (cudaStream_t)sf;
    c->s_back = (cudaStream_t)sb;
    c->gctx_pll = gp;
    c->gctx_fir = gf;
    c->part_pll_sms = (int)grp.sm.smCount;
    c->part_fir_sms = (int)rest.sm.smCount;
    return true;
}

// Guard mode (environment variable SDRB_GUARD=1 at chain creation; a debugging aid, compute-sanitizer being unavailable
// on the GPU pool): every allocation gets a canary zone on both sides, sdrb_chain_check_guards() verifies them.
constexpr size_t kGuardBytes = 512;
constexpr int kGuardByte = 0xA5;

int dalloc(sdrb_chain* c, void** p, size_t bytes, int fill = 0) {
    if (!c->guard) {
        CU(cudaMalloc(p, bytes));
        c->allocs.push_back(*p);
        CU(cudaMemsetAsync(*p, fill, bytes, c->stream));
        return SDRB_OK;
    }
    const size_t body = round_up(bytes, 256);  // keeps the alignment cudaMalloc gives and the kernels rely on
    char* base = nullptr;
    CU(cudaMalloc((void**)&base, body + 2 * kGuardBytes));
    c->allocs.push_back(base);
    CU(cudaMemsetAsync(base, kGuardByte, body + 2 * kGuardBytes, c->stream));
    CU(cudaMemsetAsync(base + kGuardBytes, fill, bytes, c->stream));
    c->guards.push_back({base, bytes});
    *p = base + kGuardBytes;
    return SDRB_OK;
}

int ring_alloc(sdrb_chain* c, Ring& r, int n, int halo) {
    r.n = n;
    r.halo = halo;
    r.pitch = round_up((size_t)halo + n + 4, 4);
    r.slot = r.pitch * c->S;
    return dalloc(c, (void**)&r.base, sizeof(float) * r.slot * kNRing);
}

void copy_taps(Taps101& t, const std::vector<float>& h) { memcpy(t.h, h.data(), sizeof(float) * kTaps); }

Timed* timer_for(sdrb_chain* c, const char* name) {
    for (auto& t : c->timed)
        if (strcmp(t.name, name) == 0) return &t;
    Timed t{name, {}, {}, 0};
    t.e0.resize(kTimedSlots);
    t.e1.resize(kTimedSlots);
    for (int i = 0; i < kTimedSlots; i++) {
        cudaEventCreate(&t.e0[i]);
        cudaEventCreate(&t.e1[i]);
    }
    c->timed.push_back(t);
    return &c->timed.back();
}

struct ScopedTimer {
    sdrb_chain* c;
    cudaStream_t st;
    Timed* t = nullptr;
    int slot = 0;
    ScopedTimer(sdrb_chain* c_, const char* name, cudaStream_t st_) : c(c_), st(st_) {
        if (c->profiling) {
            t = timer_for(c, name);
            slot = (int)(t->n % kTimedSlots);
            t->n++;
            cudaEventRecord(t->e0[slot], st);
        }
    }
    ~ScopedTimer() {
        if (t) cudaEventRecord(t->e1[slot], st);
    }
};

int check_launch(sdrb_chain* c, const char* what, cudaStream_t st) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    c->launches++;
    static const bool debug_sync = getenv("SDRB_DEBUG_SYNC") != nullptr;  // attribute an asynchronous fault to its kernel
    if (debug_sync) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return cuda_fail(e, what);
    }
    return SDRB_OK;
}

template <int DECIM>
int launch_rf(sdrb_chain* c, const RfArgs& a, cudaStream_t st) {
    dim3 grid((c->info.if_block + kRfTile - 2) / (kRfTile - 1), c->S);
    k_rf_frontend<DECIM><<<grid, kRfThreads, 0, st>>>(c->rf_h, a);
    return check_launch(c, "k_rf_frontend", st);
}

template <int DOWN>
int launch_audio_decim(sdrb_chain* c, const AudioArgs& a, cudaStream_t st) {
    dim3 grid((a.n_out + kAudTile - 1) / kAudTile, c->S);
    if (c->stereo) k_audio_decim<DOWN, true><<<grid, kAudThreads, 0, st>>>(c->audio_h, a);
    else k_audio_decim<DOWN, false><<<grid, kAudThreads, 0, st>>>(c->audio_h, a);
    return check_launch(c, "k_audio_decim", st);
}

// Host -> device copy of one block of every stream.  When the host rows are laid out with the device pitch (what a caller
// gets by allocating its pinned ring with info.block_bytes rounded up to 256) the whole batch is ONE contiguous copy; a
// pitched copy of n_streams rows otherwise.
cudaError_t copy_rows_h2d(sdrb_chain* c, uint8_t* dst, const uint8_t* h_iq, size_t iq_pitch, cudaStream_t st) {
    if (iq_pitch == c->iq_pitch)  // (the last row is only read up to its block bytes: the caller's buffer may end there)
        return cudaMemcpyAsync(dst, h_iq, c->iq_pitch * (size_t)(c->S - 1) + (size_t)c->info.block_bytes, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, c->iq_pitch, h_iq, iq_pitch, c->info.block_bytes, c->S, cudaMemcpyHostToDevice, st);
}

size_t rds_backend_smem(int n_if, int n_out) {
    const int rrc_tiles = (n_out + kRrcTile - 1) / kRrcTile;
    const size_t nfilt = (size_t)rrc_tiles * kRrcTile + kState;
    return sizeof(float) * ((size_t)rds_sdc_len(n_if) + nfilt + nfilt / kRrcR + 8);
}

// Makes the caller-visible stream wait for everything issued on the internal streams (no host blocking).
int join_main(sdrb_chain* c) {
    if (!c->pending) return SDRB_OK;
    for (cudaStream_t st : {c->s_front, c->s_pll, c->s_back, c->s_h2d, c->s_d2h}) {
        CU(cudaEventRecord(c->ev_join, st));
        CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    c->pending = false;
    return SDRB_OK;
}

// One block.  In overlap mode the three phases run on three streams chained by events:
//   front(b) = RF front end + band filters   waits: input (ev_in or the H2D copy), back(b-2)  [ring slots, see sdr_kernels.cuh]
//   pll(b)                                    waits: front(b)          (and pll(b-1): same stream)
//   back(b)  = mixers + audio + RDS           waits: pll(b)            (and back(b-1): same stream)
// so the FIR-heavy front end of block b+1 fills the machine while the latency-bound PLL of block b runs on a few SMs.
int process_block_impl(sdrb_chain* c, const uint8_t* d_iq, size_t iq_pitch, cudaEvent_t input_ready) {
    const long long b = c->block;
    const int S = c->S, n_if = c->info.if_block;
    const bool keep = c->cfg.keep_stages != 0;
    const bool ov = c->overlap;
    const cudaStream_t sf = ov ? c->s_front : c->stream, sp = ov ? c->s_pll : c->stream, sb = ov ? c->s_back : c->stream;
    int rc;
    if (ov) {
        if (input_ready) {
            CU(cudaStreamWaitEvent(sf, input_ready, 0));
        } else {
            CU(cudaEventRecord(c->ev_in, c->stream));
            CU(cudaStreamWaitEvent(sf, c->ev_in, 0));
        }
        if (b >= 2) CU(cudaStreamWaitEvent(sf, c->ev_back[(b - 2) % kNRing], 0));
        c->pending = true;
    } else if (input_ready) {
        CU(cudaStreamWaitEvent(sf, input_ready, 0));
    }
    {
        ScopedTimer tm(c, "rf_frontend", sf);
        RfArgs a{};
        a.iq = d_iq;
        a.iq_pitch = iq_pitch;
        a.halo_in = c->d_iq_halo[b & 1];
        a.halo_out = c->d_iq_halo[(b + 1) & 1];
        a.block_pairs = c->info.block_pairs;
        a.if_block = n_if;
        a.fm = c->fm.view(b);
        a.i_ds = keep ? c->d_ids : nullptr;
        a.q_ds = keep ? c->d_qds : nullptr;
        const size_t bytes16 = round_up((size_t)c->info.block_bytes, 16);
        a.tma_ok = (reinterpret_cast<uintptr_t>(d_iq) % 16 == 0) && (iq_pitch % 16 == 0) && (iq_pitch >= bytes16);
        a.row_bytes16 = a.tma_ok ? (int)bytes16 : c->info.block_bytes;
        switch (c->cfg.rf_decim) {
            case 10: rc = launch_rf<10>(c, a, sf); break;
            case 4: rc = launch_rf<4>(c, a, sf); break;
            case 3: rc = launch_rf<3>(c, a, sf); break;
            default: return fail(SDRB_ERR_INVALID, "rf_decim must be 10, 4 or 3");
        }
        if (rc) return rc;
        if (!c->input_is_host) CU(cudaEventRecord(c->ev_consumed[b & 1], sf));  // the caller's device buffer has been read
    }
    const int tiles = (n_if + kBankTile - 1) / kBankTile;
    const int bank_blocks = (int)(((long long)S * tiles + kBankWarps - 1) / kBankWarps);
    if (c->stereo && c->rds) {
        ScopedTimer tm(c, "if_bands", sf);
        BankArgs<3> a{};
        a.x = c->fm.cur(b); a.x_pitch = c->fm.pitch; a.n = n_if; a.tiles = tiles; a.n_streams = S;
        a.y[0] = c->pilot.view(b); a.y[1] = c->sband.view(b); a.y[2] = c->rband.view(b);
        a.taps[0] = c->pilot_h; a.taps[1] = c->stereo_h; a.taps[2] = c->rds_h;
        k_fir_bank<3, false><<<bank_blocks, 32 * kBankWarps, 0, sf>>>(a);
        if ((rc = check_launch(c, "k_fir_bank<3>", sf))) return rc;
    } else if (c->stereo) {
        ScopedTimer tm(c, "if_bands", sf);
        BankArgs<2> a{};
        a.x = c->fm.cur(b); a.x_pitch = c->fm.pitch; a.n = n_if; a.tiles = tiles; a.n_streams = S;
        a.y[0] = c->pilot.view(b); a.y[1] = c->sband.view(b);
        a.taps[0] = c->pilot_h; a.taps[1] = c->stereo_h;
        k_fir_bank<2, false><<<bank_blocks, 32 * kBankWarps, 0, sf>>>(a);
        if ((rc = check_launch(c, "k_fir_bank<2>", sf))) return rc;
    }
    if (c->rds) {
        ScopedTimer tm(c, "rds_carrier_bpf", sf);
        BankArgs<1> a{};
        a.x = c->rband.cur(b); a.x_pitch = c->rband.pitch; a.n = n_if; a.tiles = tiles; a.n_streams = S;
        a.y[0] = c->gpilot.view(b);
        a.taps[0] = c->rds114_h;
        k_fir_bank_scalar<1, true><<<bank_blocks, 32 * kBankWarps, 0, sf>>>(a);
        if ((rc = check_launch(c, "k_fir_bank<1,sq>", sf))) return rc;
    }
    if (ov) {
        CU(cudaEventRecord(c->ev_front[b % kNRing], sf));
        CU(cudaStreamWaitEvent(sp, c->ev_front[b % kNRing], 0));
    }
    if (c->stereo) {
        ScopedTimer tm(c, "pll", sp);
        PllArgs a{};
        a.n = n_if; a.n_streams = S;
        const float if_fs = (float)(c->cfg.rf_Fs / c->cfg.rf_decim);
        a.loop[0].x = c->pilot.cur(b); a.loop[0].x_pitch = c->pilot.pitch; a.loop[0].trig = c->trig19.view(b);
        a.loop[0].st = c->d_pll[0];
        a.loop[0].redo = c->d_pll_redo;
        a.loop[1].redo = c->d_pll_redo + 1;
        a.loop[0].coef = cr::pll_coef(19e3f, if_fs, 2.0f, 0.0f, 0.01f);        // src/stereo.cpp:77
        if (c->rds) {
            a.loop[1].x = c->gpilot.cur(b); a.loop[1].x_pitch = c->gpilot.pitch; a.loop[1].trig = c->trig114.view(b);
            a.loop[1].st = c->d_pll[1];
            a.loop[1].coef = cr::pll_coef(114e3f, (float)c->cfg.if_Fs, 0.5f, 0.0f, 0.001f);  // src/rds.cpp:119
        }
        const int loops = c->rds ? 2 : 1;
        auto launch = [&](auto threads_tag) {
            constexpr int T = decltype(threads_tag)::value;
            dim3 grid((S + T - 1) / T, loops);
            k_pll<T><<<grid, T, kPllSmemBytes, sp>>>(a);
        };
        // the smallest CTA that keeps the PLL on at most pll_max_ctas SMs (see kPllMaxCtas)
        const int cap = c->pll_max_ctas;
        if ((S + 31) / 32 * loops <= cap) launch(std::integral_constant<int, 32>{});
        else if ((S + 63) / 64 * loops <= cap) launch(std::integral_constant<int, 64>{});
        else if ((S + 127) / 128 * loops <= cap) launch(std::integral_constant<int, 128>{});
        else launch(std::integral_constant<int, 256>{});
        if ((rc = check_launch(c, "k_pll", sp))) return rc;
    }
    if (ov) {
        CU(cudaEventRecord(c->ev_pll[b % kNRing], sp));
        CU(cudaStreamWaitEvent(sb, c->ev_pll[b % kNRing], 0));
    }
    if (c->stereo) {
        ScopedTimer tm(c, "mix", sb);
        MixArgs a{};
        a.n = n_if; a.n_streams = S; a.do_stereo = 1;
        a.band = c->sband.cur(b); a.band_pitch = c->sband.pitch;
        a.trig19 = c->trig19.cur(b); a.trig19_pitch = c->trig19.pitch;
        a.stereo_dc = c->sdc.view(b);
        a.carrier_out = keep ? c->d_carrier : nullptr;
        a.scale19 = 2.0f; a.adjust19 = 0.0f; a.scale114 = 0.5f; a.adjust114 = 0.0f;
        if (c->rds) {
            a.rds_band = c->rband.cur(b); a.rds_band_pitch = c->rband.pitch;
            a.trig114 = c->trig114.cur(b); a.trig114_pitch = c->trig114.pitch;
            a.rds_dc = c->rdc.view(b);
            a.ipll_out = keep ? c->d_ipll : nullptr;
            a.delay_out = keep ? c->d_rdelay : nullptr;
        }
        dim3 grid((n_if + 1 + 256 * kMixPer - 1) / (256 * kMixPer), S);
        // k_mix needs no shared memory, which would let its CTAs slip into the few KB the PLL kernel leaves free on the SMs
        // it reserves; asking for 8 KB keeps them off those SMs (the other kernels of the chain already use more).
        if (c->rds && !keep) k_mix<true><<<grid, 256, 8192, sb>>>(a);
        else k_mix<false><<<grid, 256, 8192, sb>>>(a);
        if ((rc = check_launch(c, "k_mix", sb))) return rc;
    }
    {
        ScopedTimer tm(c, "audio", sb);
        AudioArgs a{};
        // stereo(): mono path = 50-sample all-pass delay of fm_demod (src/stereo.cpp:88); mono(): fm_demod itself
        a.mono_x = c->fm.cur(b) - (c->stereo ? 50 : 0);
        a.mono_pitch = c->fm.pitch;
        a.dc_x = c->stereo ? c->sdc.cur(b) : nullptr;
        a.dc_pitch = c->sdc.pitch;
        a.n_in = n_if; a.n_out = c->info.audio_block; a.up = c->up; a.down = c->down;
        a.taps_pm = c->d_audio_pm;
        a.taps_tm = c->d_audio_tm;
        a.thread_res = c->d_audio_res;
        a.pcm = c->d_pcm[b & 1]; a.pcm_pitch = c->pcm_pitch;
        a.mono_out = keep ? c->d_mono : nullptr;
        a.dc_out = keep ? c->d_sfilt : nullptr;
        if (c->up == 1 && c->down == 5) rc = launch_audio_decim<5>(c, a, sb);
        else if (c->up == 1 && c->down == 9) rc = launch_audio_decim<9>(c, a, sb);
        else {
            if (c->d_audio_tm) {  // phase classes: one CTA per stream, the block staged whole in shared memory
                const size_t smem = sizeof(float) * (size_t)(n_if + kState) * (c->stereo ? 2 : 1);
                if (c->stereo) k_audio_updown_pc<true><<<S, kUpdThreads, smem, sb>>>(a);
                else k_audio_updown_pc<false><<<S, kUpdThreads, smem, sb>>>(a);
                rc = check_launch(c, "k_audio_updown_pc", sb);
            } else {
                dim3 grid((a.n_out + 127) / 128, S);
                if (c->stereo) k_audio_updown<true><<<grid, 128, 0, sb>>>(a);
                else k_audio_updown<false><<<grid, 128, 0, sb>>>(a);
                rc = check_launch(c, "k_audio_updown", sb);
            }
        }
        if (rc) return rc;
    }
    if (c->rds) {
        ScopedTimer tm(c, "rds_backend", sb);
        RdsArgs a{};
        a.dc = c->rdc.cur(b); a.dc_pitch = c->rdc.pitch;
        a.n_in = n_if; a.n_out = c->info.rds_block; a.sps = 39; a.rds_on = c->cfg.rds_on;
        a.res_taps = c->d_rds_res_taps;
        a.res_lane = c->d_rds_res_lane;
        a.rrc = c->rrc_h;
        a.filt_state_in = c->d_filt_state[b & 1];
        a.filt_state_out = c->d_filt_state[(b + 1) & 1];
        a.st = c->d_rds_state;
        a.rec = c->d_rec[b & 1];
        a.filt_out = keep ? c->d_rfilt : nullptr;
        a.clean_out = keep ? c->d_rclean : nullptr;
        a.overflow = c->d_rds_overflow;
        const size_t smem = rds_backend_smem(n_if, a.n_out);
        k_rds_backend<<<S, kRdsThreads, smem, sb>>>(a);
        if ((rc = check_launch(c, "k_rds_backend", sb))) return rc;
    }
    if (ov) CU(cudaEventRecord(c->ev_back[b % kNRing], sb));
    c->block = b + 1;
    c->blocks_since_load++;
    return SDRB_OK;
}

// A failure after the first kernel of a block was enqueued leaves halos, rings and PLL state partly advanced while
// c->block is not: such a chain cannot continue (or be retried) bit-exactly, so it is marked unusable.
int process_block(sdrb_chain* c, const uint8_t* d_iq, size_t iq_pitch, cudaEvent_t input_ready) {
    if (c->poisoned) return fail(SDRB_ERR_STATE, "chain is unusable: a kernel launch failed in the middle of an earlier block");
    const long long launches0 = c->launches;
    const int rc = process_block_impl(c, d_iq, iq_pitch, input_ready);
    if (rc != SDRB_OK && (c->launches != launches0 || rc == SDRB_ERR_CUDA)) c->poisoned = 1;
    return rc;
}

}  // namespace

extern "C" {

const char* sdrb_last_error(void) { return g_err.c_str(); }
int sdrb_version(void) { return 100; }

int sdrb_config_for_mode(int mode, int type, int n_streams, sdrb_config* cfg) {
    if (!cfg) return fail(SDRB_ERR_INVALID, "cfg is null");
    if (type != 'm' && type != 's' && type != 'r') return fail(SDRB_ERR_INVALID, "type must be 'm', 's' or 'r'");
    memset(cfg, 0, sizeof(*cfg));
    // defaults, /root/reference/src/project.cpp:31-44
    cfg->rf_Fs = 2400000; cfg->rf_Fc = 100000; cfg->rf_taps = kTaps; cfg->rf_decim = 10;
    cfg->audio_decim = 5; cfg->audio_upsample = 1; cfg->if_Fs = 240000; cfg->audio_Fc = 16000;
    cfg->audio_Fs = 48000; cfg->symbol_Fs = 39;
    switch (mode) {  // :67-108
        case 0: break;
        case 1: cfg->rf_Fs = 1440000; cfg->rf_decim = 4; cfg->audio_decim = 9; cfg->if_Fs = 360000; cfg->audio_Fs = 40000; break;
        case 2: cfg->audio_decim = 800; cfg->audio_upsample = 147; cfg->audio_Fs = 44100; cfg->symbol_Fs = 20; break;
        case 3: cfg->rf_Fs = 1152000; cfg->rf_decim = 3; cfg->audio_decim = 1280; cfg->audio_upsample = 147;
                cfg->if_Fs = 384000; cfg->audio_Fs = 44100; cfg->symbol_Fs = 20; break;
        default: return fail(SDRB_ERR_INVALID, "mode must be 0..3");
    }
    cfg->rds_on = (type == 'r');  // :111-132
    cfg->type = type;
    cfg->n_streams = n_streams;
    cfg->device = 0;
    cfg->keep_stages = 0;
    return SDRB_OK;
}

int sdrb_chain_destroy(sdrb_chain* c) {
    if (!c) return SDRB_OK;
    cudaSetDevice(c->cfg.device);
    for (cudaStream_t st : {c->s_front, c->s_pll, c->s_back, c->s_h2d, c->s_d2h, c->stream})
        if (st) cudaStreamSynchronize(st);
    for (void* p : c->allocs) cudaFree(p);
    for (cudaStream_t st : {c->s_front, c->s_pll, c->s_back, c->s_h2d, c->s_d2h})
        if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : {c->ev_in, c->ev_join, c->ev_h2d[0], c->ev_h2d[1], c->ev_consumed[0], c->ev_consumed[1]})
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < kNRing; i++)
        for (cudaEvent_t e : {c->ev_front[i], c->ev_pll[i], c->ev_back[i]})
            if (e) cudaEventDestroy(e);
    for (auto& t : c->timed)
        for (int i = 0; i < kTimedSlots; i++) {
            cudaEventDestroy(t.e0[i]);
            cudaEventDestroy(t.e1[i]);
        }
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
    for (void* g : {c->gctx_pll, c->gctx_fir})
        if (g) green_api().GreenCtxDestroy((CUgreenCtx)g);
    delete c;
    return SDRB_OK;
}

int sdrb_chain_create(const sdrb_config* cfg, sdrb_chain** out) {
    if (!cfg || !out) return fail(SDRB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->n_streams < 1 || cfg->n_streams > 65535) return fail(SDRB_ERR_INVALID, "n_streams must be in 1..65535");
    if (cfg->rf_taps != kTaps) return fail(SDRB_ERR_INVALID, "rf_taps must be 101");
    if (cfg->type != 'm' && cfg->type != 's' && cfg->type != 'r') return fail(SDRB_ERR_INVALID, "bad type");
    if (cfg->rf_decim != 10 && cfg->rf_decim != 4 && cfg->rf_decim != 3) return fail(SDRB_ERR_INVALID, "rf_decim must be 10, 4 or 3");
    if (cfg->audio_upsample < 1 || cfg->audio_decim < 1) return fail(SDRB_ERR_INVALID, "bad audio resampling ratio");
    if (cfg->rf_Fs <= 0 || cfg->if_Fs <= 0 || cfg->audio_Fc <= 0 || cfg->rf_Fc <= 0) return fail(SDRB_ERR_INVALID, "rates and cut-offs must be positive");
    if ((long long)kTaps * cfg->audio_upsample > 65535) return fail(SDRB_ERR_INVALID, "101 * audio_upsample must fit an unsigned short (include/filter.h:19)");
    if ((1470LL * cfg->audio_decim) / cfg->audio_upsample < kState + 12) return fail(SDRB_ERR_INVALID, "block too short for the 101-tap filters");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = "no CUDA device available (libsdr_b200 has no CPU fallback)";
        if (e != cudaSuccess) g_err += std::string(": ") + cudaGetErrorString(e);
        return SDRB_ERR_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SDRB_ERR_INVALID, "bad device ordinal");
    CU(cudaSetDevice(cfg->device));

    sdrb_chain* c = new sdrb_chain();
    c->cfg = *cfg;
    c->S = cfg->n_streams;
    // SMs given to k_pll: 32 (two warps per SM at 1024 stereo+RDS stations).  Round 1 used 64 (one warp per SM) up to 1024
    // stations; with the rotated PLL loop of round 2 the FIR kernels on the remaining 84 SMs had become the longer side of
    // the step (profiles/README.md: 1.058 -> 0.979 ms per step at 1024 stations with 32; the PLL alone loses 1 %).
    {
        const int loops = cfg->type == 'r' ? 2 : 1;
        (void)loops;
        c->pll_max_ctas = kPllMaxCtas / 2;
    }
    if (const char* e = getenv("SDRB_PLL_MAX_CTAS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 148) c->pll_max_ctas = v;
    }
    if (const char* e = getenv("SDRB_GUARD")) c->guard = atoi(e) != 0;
    c->up = cfg->audio_upsample;
    c->down = cfg->audio_decim;
    c->stereo = cfg->type != 'm';
    // The RDS back end is built for the one configuration in which the reference's rds() decodes anything: 240 kHz IF and
    // 39 samples per symbol (mode 0, src/project.cpp:67-74).  `project <1|2|3> r` still plays stereo audio while its rds
    // thread runs the 57 kHz chain at rates it was not designed for and prints nothing (checked against the threaded
    // reference binary in tests/); here such a chain is the stereo chain with every RDS record gated.
    c->rds = cfg->type == 'r' && cfg->if_Fs == 240000 && cfg->audio_upsample == 1 && cfg->rf_Fs / cfg->rf_decim == 240000;
    c->rds_gated = cfg->type == 'r' && !c->rds;
    sdrb_chain_info& I = c->info;
    I.block_pairs = (1470 * cfg->rf_decim * c->down) / c->up;  // src/rffrontend.cpp:21
    I.block_bytes = 2 * I.block_pairs;
    I.if_block = (1470 * c->down) / c->up;                     // src/mono.cpp:19
    I.audio_block = (int)(((long long)I.if_block * c->up) / c->down);
    I.pcm_per_block = I.audio_block * (c->stereo ? 2 : 1);
    I.rds_block = c->rds ? (int)(((long long)I.if_block * kRdsUp) / kRdsDown) : 0;
    I.max_bits = kRdsMaxBits;
    I.max_groups = kRdsMaxGroups;
    const int S = c->S, n_if = I.if_block;

#define TRY(x)                        \
    do {                              \
        int rc_ = (x);                \
        if (rc_) {                    \
            sdrb_chain_destroy(c);    \
            return rc_;               \
        }                             \
    } while (0)
#define TRYCU(x)                                   \
    do {                                           \
        cudaError_t e_ = (x);                      \
        if (e_ != cudaSuccess) {                   \
            int rc_ = cuda_fail(e_, #x);           \
            sdrb_chain_destroy(c);                 \
            return rc_;                            \
        }                                          \
    } while (0)

    TRYCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    {
        // The PLL kernel needs whole SMs (it reserves their shared memory) and is the longest dependency chain of a step:
        // its stream gets the highest priority so the block scheduler stops refilling SMs with FIR CTAs while PLL CTAs wait.
        int prio_lo = 0, prio_hi = 0;
        TRYCU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (cudaStream_t* st : {&c->s_h2d, &c->s_d2h}) TRYCU(cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, prio_lo));
        // SM partition for the overlap-mode streams (see make_sm_partition); SDRB_SM_PARTITION=0 keeps plain priority streams
        // Only while the PLL is the longer side of a step (at most two of its warps per SM of the partition: 1024 stereo+RDS
        // stations).  Larger batches are bound by the FIR kernels, which then want the PLL's SMs whenever it is idle
        // (4096 stations: 2.47 ms per step without the partition, 2.68 with it).  SDRB_SM_PARTITION=0 / 1 forces it off / on.
        const char* pe = getenv("SDRB_SM_PARTITION");
        const int pll_warps = (cfg->n_streams + 31) / 32 * (cfg->type == 'r' ? 2 : 1);
        const bool want_part = pe ? atoi(pe) != 0 : pll_warps <= 2 * c->pll_max_ctas;
        if (!(want_part && make_sm_partition(c, c->pll_max_ctas, prio_lo, prio_hi))) {
            for (cudaStream_t* st : {&c->s_front, &c->s_back}) TRYCU(cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, prio_lo));
            TRYCU(cudaStreamCreateWithPriority(&c->s_pll, cudaStreamNonBlocking, prio_hi));
        }
    }
    for (cudaEvent_t* e : {&c->ev_in, &c->ev_join, &c->ev_h2d[0], &c->ev_h2d[1], &c->ev_consumed[0], &c->ev_consumed[1]}) TRYCU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < kNRing; i++)
        for (cudaEvent_t* e : {&c->ev_front[i], &c->ev_pll[i], &c->ev_back[i]}) TRYCU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));

    // ---- taps (all designed on the host, same libm as the reference build)
    std::vector<float> h(kTaps);
    const float if_fs_f = (float)(cfg->rf_Fs / cfg->rf_decim);
    TRY(sdrb_design_lpf((float)cfg->rf_Fs, (float)cfg->rf_Fc, kTaps, h.data()));            // src/rffrontend.cpp:24
    copy_taps(c->rf_h, h);
    TRY(sdrb_design_bpf(if_fs_f, 18.5e3f, 19.5e3f, kTaps, h.data()));                       // src/stereo.cpp:65
    copy_taps(c->pilot_h, h);
    TRY(sdrb_design_bpf(if_fs_f, 22e3f, 54e3f, kTaps, h.data()));                           // src/stereo.cpp:67
    copy_taps(c->stereo_h, h);
    TRY(sdrb_design_bpf((float)cfg->if_Fs, 54e3f, 60e3f, kTaps, h.data()));                 // src/rds.cpp:62
    copy_taps(c->rds_h, h);
    TRY(sdrb_design_bpf((float)cfg->if_Fs, 113.5e3f, 114.5e3f, kTaps, h.data()));           // src/rds.cpp:63
    copy_taps(c->rds114_h, h);
    TRY(sdrb_design_rrc((float)(2375 * 39), kTaps, h.data()));                              // src/rds.cpp:65 (sps = 39)
    copy_taps(c->rrc_h, h);
    {
        const int nh = kTaps * c->up;
        std::vector<float> ah(nh);
        TRY(sdrb_design_lpf_gain((float)cfg->if_Fs * (float)c->up, (float)cfg->audio_Fc, nh, c->up, ah.data()));  // src/mono.cpp:22
        if (c->up == 1) {
            copy_taps(c->audio_h, ah);
        } else {
            std::vector<float> pm((size_t)c->up * kTaps);
            for (int p = 0; p < c->up; p++)
                for (int j = 0; j < kTaps; j++) pm[(size_t)p * kTaps + j] = ah[p + c->up * j];
            TRY(dalloc(c, (void**)&c->d_audio_pm, pm.size() * sizeof(float)));
            TRYCU(cudaMemcpyAsync(c->d_audio_pm, pm.data(), pm.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            // phase-class resampler (k_audio_updown_pc): applicable when every residue class has the same number of outputs
            const int n_if_ = (1470 * c->down) / c->up, n_out_ = (int)(((long long)n_if_ * c->up) / c->down);
            const size_t smem_pc = sizeof(float) * (size_t)(n_if_ + kState) * (c->stereo ? 2 : 1);
            if (c->up <= kUpdThreads && n_out_ % c->up == 0 && n_out_ / c->up <= kUpdMaxQ && smem_pc <= 200 * 1024) {
                // thread -> residue: a half-warp's loads (64-bit for stereo, 32-bit for mono: then a whole warp) hit distinct
                // banks when the input offsets floor(r down / up) differ modulo 16 (32) among its lanes
                const int group = c->stereo ? 16 : 32;
                std::vector<int> res(kUpdThreads, -1), left;
                std::vector<int> next(group);
                for (int k = 0; k < group; k++) next[k] = k;
                for (int r = 0; r < c->up; r++) {
                    int& sl = next[(int)(((long long)r * c->down) / c->up) % group];
                    if (sl < kUpdThreads) { res[sl] = r; sl += group; }
                    else left.push_back(r);
                }
                for (int th = 0; th < kUpdThreads && !left.empty(); th++)
                    if (res[th] < 0) { res[th] = left.back(); left.pop_back(); }
                std::vector<float> tm((size_t)kTaps * kUpdThreads, 0.0f);
                for (int th = 0; th < kUpdThreads; th++) {
                    if (res[th] < 0) continue;
                    const int phase = (int)(((long long)res[th] * c->down) % c->up);
                    for (int j = 0; j < kTaps; j++) tm[(size_t)j * kUpdThreads + th] = ah[phase + c->up * j];
                }
                TRY(dalloc(c, (void**)&c->d_audio_tm, tm.size() * sizeof(float)));
                TRY(dalloc(c, (void**)&c->d_audio_res, kUpdThreads * sizeof(int)));
                TRYCU(cudaMemcpyAsync(c->d_audio_tm, tm.data(), tm.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
                TRYCU(cudaMemcpyAsync(c->d_audio_res, res.data(), kUpdThreads * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                TRYCU(cudaFuncSetAttribute(k_audio_updown_pc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pc));
                TRYCU(cudaFuncSetAttribute(k_audio_updown_pc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pc));
            }
            TRYCU(cudaStreamSynchronize(c->stream));
        }
    }
    if (c->rds) {
        const int nh = kTaps * kRdsUp;
        std::vector<float> lh(nh);
        TRY(sdrb_design_lpf_gain((float)(cfg->if_Fs * kRdsUp), 3e3f, nh, kRdsUp, lh.data()));  // src/rds.cpp:61
        // resampler lanes and their tap table (res_lanes.h)
        static_assert(sizeof(ResLane) == sizeof(int4) && sizeof(float2) == 2 * sizeof(float), "uploaded as int4 / float2");
        std::vector<ResLane> lanes(kResLanes);
        std::vector<float> rtaps((size_t)kResIter * kResLanes * 2);
        if (build_res_lanes(lh.data(), lanes.data(), rtaps.data()) < 0) { sdrb_chain_destroy(c); return fail(SDRB_ERR_INVALID, "resampler lanes: no free lane"); }
        TRY(dalloc(c, (void**)&c->d_rds_res_taps, rtaps.size() * sizeof(float)));
        TRY(dalloc(c, (void**)&c->d_rds_res_lane, lanes.size() * sizeof(ResLane)));
        TRYCU(cudaMemcpyAsync(c->d_rds_res_taps, rtaps.data(), rtaps.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        TRYCU(cudaMemcpyAsync(c->d_rds_res_lane, lanes.data(), lanes.size() * sizeof(ResLane), cudaMemcpyHostToDevice, c->stream));
        TRYCU(cudaStreamSynchronize(c->stream));
    }

    // ---- input staging and halo (bytes; 128 unpacks to 0.0f = the reference's zero-initialised state)
    c->iq_pitch = round_up((size_t)I.block_bytes, 256);
    for (int i = 0; i < 2; i++) {
        TRY(dalloc(c, (void**)&c->d_iq[i], c->iq_pitch * S));
        TRY(dalloc(c, (void**)&c->d_iq_halo[i], (size_t)2 * kIqHaloPairs * S, 128));
    }
    // ---- rings
    TRY(ring_alloc(c, c->fm, n_if, 160));
    if (c->stereo) {
        TRY(ring_alloc(c, c->pilot, n_if, 0));
        TRY(ring_alloc(c, c->sband, n_if, 0));
        TRY(ring_alloc(c, c->trig19, n_if, 4));
        TRY(ring_alloc(c, c->sdc, n_if, 112));
        TRY(dalloc(c, (void**)&c->d_pll[0], sizeof(PllStateDev) * S));
        TRY(dalloc(c, (void**)&c->d_pll_redo, 20 * sizeof(unsigned long long)));  // [2] totals + 9 x [2] per-test counters of diagnostic builds
    }
    if (c->rds) {
        TRY(ring_alloc(c, c->rband, n_if, 160));
        TRY(ring_alloc(c, c->gpilot, n_if, 0));
        TRY(ring_alloc(c, c->trig114, n_if, 4));
        TRY(ring_alloc(c, c->rdc, n_if, 112));
        TRY(dalloc(c, (void**)&c->d_pll[1], sizeof(PllStateDev) * S));
        for (int i = 0; i < 2; i++) TRY(dalloc(c, (void**)&c->d_filt_state[i], sizeof(float) * kState * S));
        TRY(dalloc(c, (void**)&c->d_rds_state, sizeof(RdsStreamState) * S));
        for (int i = 0; i < 2; i++) TRY(dalloc(c, (void**)&c->d_rec[i], sizeof(RdsRecord) * S));
        TRY(dalloc(c, (void**)&c->d_rds_overflow, 3 * sizeof(unsigned int)));
    }
    {   // PLL initial state: feedbackI = 1, rest 0 (src/stereo.cpp:51-57, src/rds.cpp:51-56)
        std::vector<PllStateDev> init(S, PllStateDev{1.0f, 0.0f, 0.0f, 0.0f, 0.0});
        for (int i = 0; i < 2; i++)
            if (c->d_pll[i]) TRYCU(cudaMemcpyAsync(c->d_pll[i], init.data(), sizeof(PllStateDev) * S, cudaMemcpyHostToDevice, c->stream));
        if (c->d_rds_state) {
            std::vector<RdsStreamState> rs(S);
            memset(rs.data(), 0, sizeof(RdsStreamState) * S);
            for (auto& r : rs) r.first_time = 1;
            TRYCU(cudaMemcpyAsync(c->d_rds_state, rs.data(), sizeof(RdsStreamState) * S, cudaMemcpyHostToDevice, c->stream));
        }
        TRYCU(cudaStreamSynchronize(c->stream));
    }
    c->pcm_pitch = round_up((size_t)I.pcm_per_block, 8);
    for (int i = 0; i < 2; i++) TRY(dalloc(c, (void**)&c->d_pcm[i], sizeof(int16_t) * c->pcm_pitch * S));
    if (cfg->keep_stages) {
        TRY(dalloc(c, (void**)&c->d_ids, sizeof(float) * n_if * S));
        TRY(dalloc(c, (void**)&c->d_qds, sizeof(float) * n_if * S));
        TRY(dalloc(c, (void**)&c->d_mono, sizeof(float) * I.audio_block * S));
        if (c->stereo) {
            TRY(dalloc(c, (void**)&c->d_carrier, sizeof(float) * (n_if + 1) * S));
            TRY(dalloc(c, (void**)&c->d_sfilt, sizeof(float) * I.audio_block * S));
        }
        if (c->rds) {
            TRY(dalloc(c, (void**)&c->d_ipll, sizeof(float) * (n_if + 1) * S));
            TRY(dalloc(c, (void**)&c->d_rdelay, sizeof(float) * n_if * S));
            TRY(dalloc(c, (void**)&c->d_rfilt, sizeof(float) * I.rds_block * S));
            TRY(dalloc(c, (void**)&c->d_rclean, sizeof(float) * I.rds_block * S));
        }
    }
    static_assert(pll_tile_bytes(256) <= kPllSmemBytes, "PLL input ring must fit the reserved shared memory");
    if (c->stereo) {
        TRYCU(cudaFuncSetAttribute(k_pll<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPllSmemBytes));
        TRYCU(cudaFuncSetAttribute(k_pll<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPllSmemBytes));
        TRYCU(cudaFuncSetAttribute(k_pll<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPllSmemBytes));
        TRYCU(cudaFuncSetAttribute(k_pll<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPllSmemBytes));
    }
    if (c->rds) {
        if (I.rds_block > kResQ * kRdsUp) { sdrb_chain_destroy(c); return fail(SDRB_ERR_INVALID, "RDS block too long for the resampler kernel"); }
        TRYCU(cudaFuncSetAttribute(k_rds_backend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rds_backend_smem(n_if, I.rds_block)));
    }
    TRYCU(cudaStreamSynchronize(c->stream));
#undef TRY
#undef TRYCU
    *out = c;
    return SDRB_OK;
}

int sdrb_chain_get_info(const sdrb_chain* c, sdrb_chain_info* info) {
    if (!c || !info) return fail(SDRB_ERR_INVALID, "null argument");
    *info = c->info;
    return SDRB_OK;
}

int sdrb_chain_process_device(sdrb_chain* c, const uint8_t* d_iq, size_t iq_pitch) {
    if (!c || !d_iq) return fail(SDRB_ERR_INVALID, "null argument");
    if (iq_pitch < (size_t)c->info.block_bytes || (iq_pitch & 1)) return fail(SDRB_ERR_INVALID, "iq_pitch too small or odd");
    if ((reinterpret_cast<uintptr_t>(d_iq) & 1)) return fail(SDRB_ERR_INVALID, "d_iq must be 2-byte aligned");
    CU(cudaSetDevice(c->cfg.device));
    c->input_is_host = false;
    return process_block(c, d_iq, iq_pitch, nullptr);
}

int sdrb_chain_process_host(sdrb_chain* c, const uint8_t* h_iq, size_t iq_pitch) {
    if (!c || !h_iq) return fail(SDRB_ERR_INVALID, "null argument");
    if (iq_pitch < (size_t)c->info.block_bytes) return fail(SDRB_ERR_INVALID, "iq_pitch too small");
    CU(cudaSetDevice(c->cfg.device));
    const long long b = c->block;
    uint8_t* dst = c->d_iq[b & 1];
    if (c->poisoned) return fail(SDRB_ERR_STATE, "chain is unusable: a kernel launch failed in the middle of an earlier block");
    c->input_is_host = true;
    if (!c->overlap) {
        CU(copy_rows_h2d(c, dst, h_iq, iq_pitch, c->stream));
        CU(cudaEventRecord(c->ev_consumed[b & 1], c->stream));
        return process_block(c, dst, c->iq_pitch, nullptr);
    }
    // copy engine stream: the staging buffer of parity b was last read by the front end of block b-2
    if (b >= 2) CU(cudaStreamWaitEvent(c->s_h2d, c->ev_front[(b - 2) % kNRing], 0));
    CU(cudaEventRecord(c->ev_in, c->stream));
    CU(cudaStreamWaitEvent(c->s_h2d, c->ev_in, 0));
    CU(copy_rows_h2d(c, dst, h_iq, iq_pitch, c->s_h2d));
    CU(cudaEventRecord(c->ev_h2d[b & 1], c->s_h2d));
    CU(cudaEventRecord(c->ev_consumed[b & 1], c->s_h2d));
    c->pending = true;
    return process_block(c, dst, c->iq_pitch, c->ev_h2d[b & 1]);
}

int sdrb_chain_input_consumed(sdrb_chain* c, int lag) {
    if (!c || lag < 0 || lag > 1) return -(int)fail(SDRB_ERR_INVALID, "bad argument");
    if (c->blocks_since_load <= lag) return -(int)fail(SDRB_ERR_STATE, "that block has not been issued");
    if (cudaSetDevice(c->cfg.device) != cudaSuccess) return -(int)SDRB_ERR_CUDA;
    const cudaError_t e = cudaEventQuery(c->ev_consumed[(c->block - 1 - lag) & 1]);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    return -cuda_fail(e, "cudaEventQuery");
}

int sdrb_chain_join(sdrb_chain* c) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    return join_main(c);
}

int sdrb_chain_sync(sdrb_chain* c) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    int rc = join_main(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return SDRB_OK;
}

int sdrb_chain_set_overlap(sdrb_chain* c, int on) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    int rc = join_main(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));  // switching modes between blocks: nothing may be in flight
    c->overlap = on != 0;
    return SDRB_OK;
}

int sdrb_chain_set_stream(sdrb_chain* c, void* stream) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    int rc = join_main(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    if (c->own_stream) {
        cudaStreamDestroy(c->stream);
        c->own_stream = false;
    }
    c->stream = (cudaStream_t)stream;
    return SDRB_OK;
}

int sdrb_pinned_alloc(size_t bytes, void** h_ptr) {
    if (!h_ptr || bytes == 0) return fail(SDRB_ERR_INVALID, "bad argument");
    CU(cudaHostAlloc(h_ptr, bytes, cudaHostAllocDefault));
    return SDRB_OK;
}

int sdrb_pinned_free(void* h_ptr) {
    if (h_ptr) CU(cudaFreeHost(h_ptr));
    return SDRB_OK;
}

// Results of block (most recent - lag), lag 0 or 1.  With lag 1 only the back end of that block is waited for, so the
// device-to-host copies overlap the block in flight (outputs are double buffered by block parity).
int sdrb_chain_read_results(sdrb_chain* c, int lag, int16_t* h_pcm, size_t pcm_pitch, sdrb_rds_record* h_records) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    if (lag < 0 || lag > 1) return fail(SDRB_ERR_INVALID, "lag must be 0 or 1");
    if (c->poisoned) return fail(SDRB_ERR_STATE, "chain is unusable: a kernel launch failed in the middle of an earlier block");
    if (c->block - lag <= 0 || c->blocks_since_load <= lag) return fail(SDRB_ERR_STATE, "that block has not been processed yet");
    if (h_pcm && pcm_pitch < (size_t)c->info.pcm_per_block) return fail(SDRB_ERR_INVALID, "pcm_pitch too small");
    if (h_records && !c->rds && !c->rds_gated) return fail(SDRB_ERR_STATE, "chain was created without RDS (type != 'r')");
    if (h_records && c->rds_gated) {  // no RDS back end in this mode: every block reads as gated (cdr_offset = -1)
        memset(h_records, 0, sizeof(sdrb_rds_record) * c->S);
        for (int s = 0; s < c->S; s++) h_records[s].cdr_offset = -1;
        h_records = nullptr;
    }
    CU(cudaSetDevice(c->cfg.device));
    const long long b = c->block - 1 - lag;
    cudaStream_t st = c->stream;
    if (c->overlap) {
        st = c->s_d2h;
        CU(cudaStreamWaitEvent(st, c->ev_back[b % kNRing], 0));
    }
    if (h_pcm)
        CU(cudaMemcpy2DAsync(h_pcm, pcm_pitch * sizeof(int16_t), c->d_pcm[b & 1], c->pcm_pitch * sizeof(int16_t),
                             c->info.pcm_per_block * sizeof(int16_t), c->S, cudaMemcpyDeviceToHost, st));
    if (h_records) CU(cudaMemcpyAsync(h_records, c->d_rec[b & 1], sizeof(RdsRecord) * c->S, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SDRB_OK;
}

int sdrb_chain_read_pcm(sdrb_chain* c, int16_t* h_pcm, size_t pcm_pitch) {
    if (!c || !h_pcm) return fail(SDRB_ERR_INVALID, "null argument");
    return sdrb_chain_read_results(c, 0, h_pcm, pcm_pitch, nullptr);
}

int sdrb_chain_pcm_device(sdrb_chain* c, const int16_t** d_pcm, size_t* pcm_pitch) {
    if (!c || !d_pcm || !pcm_pitch) return fail(SDRB_ERR_INVALID, "null argument");
    *d_pcm = c->d_pcm[(c->block > 0 ? c->block - 1 : 0) & 1];
    *pcm_pitch = c->pcm_pitch;
    return SDRB_OK;
}

int sdrb_chain_read_rds(sdrb_chain* c, sdrb_rds_record* h_records) {
    if (!c || !h_records) return fail(SDRB_ERR_INVALID, "null argument");
    return sdrb_chain_read_results(c, 0, nullptr, 0, h_records);
}

// parse(), /root/reference/src/rds_utilities.cpp:172-199 (+ stringify :111-119): "PI: " in lower-case hex
// without padding (std::hex), "PTY: " by name, on every group; type-0 groups fill the 8-character PS
// buffer and print it when segment 3 arrives and the buffer changed (printed as a C string).
int sdrb_rds_parse(uint64_t group, uint64_t* chars, uint64_t* output, char* text, int text_cap) {
    static const char* const pty_names[32] = {  // :137-170
        "Undefined", "News", "Information", "Sports", "Talk", "Rock", "Classic Rock", "Adult Hits", "Soft Rock",
        "Top 40", "Country", "Oldies", "Soft", "Nostalgia", "Jazz", "Classical", "Rhythm & Blues",
        "Soft Rhythm & Blues", "Language", "Religious Music", "Religious Talk", "Personality", "Public", "College",
        "Spanish Talk", "Spanish Music", "Hip Hop", "Unassigned", "Unassigned", "Weather", "Emergency Test",
        "Emergency"};
    if (!chars || !output || !text || text_cap < 1) return 0;
    const unsigned group_type = (unsigned)(group >> 44) & 0xF;
    const unsigned segment = (unsigned)(group >> 32) & 0x3;
    const unsigned pi = (unsigned)(group >> 48) & 0xFFFF;
    const unsigned pty = (unsigned)(group >> 37) & 0x1F;
    std::string out;
    char line[96];
    snprintf(line, sizeof line, "PI: %x\nPTY: %s\n", pi, pty_names[pty]);
    out += line;
    if (group_type == 0) {
        const int sh = 16 * (3 - (int)segment);
        *chars = (*chars & ~((uint64_t)0xFFFF << sh)) | ((group & 0xFFFFull) << sh);
        if (segment == 3 && *chars != *output) {
            *output = *chars;
            char ps[9];
            for (int i = 0; i < 8; i++) ps[i] = (char)((*chars >> (8 * (7 - i))) & 0xFF);
            ps[8] = 0;
            out += "Program Service: ";
            out += ps;  // C-string semantics: stops at the first NUL
            out += "\n";
        }
    }
    int n = (int)out.size() < text_cap - 1 ? (int)out.size() : text_cap - 1;
    memcpy(text, out.data(), n);
    text[n] = 0;
    return n;
}

int sdrb_chain_stage(sdrb_chain* c, const char* name, float* h_out, int cap_per_stream, int* count) {
    if (!c || !name || !h_out || !count) return fail(SDRB_ERR_INVALID, "null argument");
    if (c->block == 0 || c->blocks_since_load == 0) return fail(SDRB_ERR_STATE, "no block processed yet");
    const long long b = c->block - 1;
    const int n_if = c->info.if_block;
    const float* src = nullptr;
    size_t pitch = 0;
    int n = 0;
    std::string s(name);
    auto ring = [&](const Ring& r) { src = r.base ? r.cur(b) : nullptr; pitch = r.pitch; n = r.n; };
    // ring-backed stages always exist; the flat ones are dumps that only a keep_stages chain writes
    auto flat = [&](const float* p, int cnt) { src = c->cfg.keep_stages ? p : nullptr; pitch = cnt; n = cnt; };
    if (s == "fm_demod") ring(c->fm);
    else if (s == "pilot") ring(c->pilot);
    else if (s == "stereo_band") ring(c->sband);
    else if (s == "stereo_dc") ring(c->sdc);
    else if (s == "rds_band") ring(c->rband);
    else if (s == "gen_pilot") ring(c->gpilot);
    else if (s == "rds_dc") ring(c->rdc);
    else if (s == "I_ds") flat(c->d_ids, n_if);
    else if (s == "Q_ds") flat(c->d_qds, n_if);
    else if (s == "carrier") flat(c->d_carrier, n_if + 1);
    else if (s == "IPLL") flat(c->d_ipll, n_if + 1);
    else if (s == "rds_band_delay") flat(c->d_rdelay, n_if);
    else if (s == "mono_filt" || s == "audio_filt") flat(c->d_mono, c->info.audio_block);
    else if (s == "stereo_filt") flat(c->d_sfilt, c->info.audio_block);
    else if (s == "rds_filt") flat(c->d_rfilt, c->info.rds_block);
    else if (s == "rds_clean") flat(c->d_rclean, c->info.rds_block);
    else return fail(SDRB_ERR_INVALID, "unknown stage name: " + s);
    if (!src) return fail(SDRB_ERR_STATE, "stage not produced in this configuration (flat stages need keep_stages = 1): " + s);
    *count = n;
    if (cap_per_stream < n) return fail(SDRB_ERR_INVALID, "cap_per_stream too small");
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaMemcpy2DAsync(h_out, (size_t)cap_per_stream * sizeof(float), src, pitch * sizeof(float), (size_t)n * sizeof(float),
                         c->S, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDRB_OK;
}

// ---- carried state (checkpoint / resume) ---------------------------------------------------------
// Blob = header {magic, n_streams, if_block, type, block} followed by, per item, the raw device arrays
// that carry information from one block to the next: the IQ halo, the halo of every ring (as seen by the
// next block), both PLL states, the RDS resampler state, the per-stream decoder state.
namespace {
struct StateItem {
    void* ptr;        // start (device)
    size_t row_bytes; // bytes per stream row to save
    size_t pitch_bytes;
    int rows;
    const char* name; // for sdrb_chain_state_item_offset
};
std::vector<StateItem> state_items(sdrb_chain* c, long long b /* block that will be processed next */) {
    std::vector<StateItem> v;
    const int S = c->S;
    v.push_back({c->d_iq_halo[b & 1], (size_t)2 * kIqHaloPairs, (size_t)2 * kIqHaloPairs, S, "iq_halo"});
    for (Ring* r : {&c->fm, &c->pilot, &c->sband, &c->rband, &c->gpilot, &c->trig19, &c->trig114, &c->sdc, &c->rdc})
        if (r->base && r->halo > 0)
            v.push_back({r->base + (size_t)(b % kNRing) * r->slot, (size_t)r->halo * sizeof(float), r->pitch * sizeof(float), S, "ring_halo"});
    for (int i = 0; i < 2; i++)
        if (c->d_pll[i]) v.push_back({c->d_pll[i], sizeof(PllStateDev) * S, sizeof(PllStateDev) * S, 1, i ? "pll114" : "pll19"});
    if (c->d_filt_state[0]) v.push_back({c->d_filt_state[b & 1], sizeof(float) * kState * S, sizeof(float) * kState * S, 1, "rds_filt_state"});
    if (c->d_rds_state) v.push_back({c->d_rds_state, sizeof(RdsStreamState) * S, sizeof(RdsStreamState) * S, 1, "rds_decoder"});
    return v;
}
struct StateHeader {
    uint32_t magic, n_streams, if_block, type;
    long long block;
    unsigned long long total_bytes;  // of the whole blob, header included
};
constexpr uint32_t kStateMagic = 0x53445244u;
}  // namespace

size_t sdrb_chain_state_bytes(const sdrb_chain* c) {
    if (!c) return 0;
    size_t total = sizeof(StateHeader);
    for (auto& it : state_items(const_cast<sdrb_chain*>(c), c->block)) total += it.row_bytes * it.rows;
    return total;
}

long long sdrb_chain_state_item_offset(const sdrb_chain* c, const char* name) {
    if (!c || !name) return -1;
    size_t off = sizeof(StateHeader);
    for (auto& it : state_items(const_cast<sdrb_chain*>(c), c->block)) {
        if (strcmp(it.name, name) == 0) return (long long)off;
        off += it.row_bytes * it.rows;
    }
    return -1;
}

int sdrb_chain_state_save(sdrb_chain* c, void* h_blob) {
    if (!c || !h_blob) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaStreamSynchronize(c->stream));
    if (c->poisoned) return fail(SDRB_ERR_STATE, "chain is unusable: a kernel launch failed in the middle of an earlier block");
    StateHeader hd{kStateMagic, (uint32_t)c->S, (uint32_t)c->info.if_block, (uint32_t)c->cfg.type, c->block, sdrb_chain_state_bytes(c)};
    char* p = static_cast<char*>(h_blob);
    memcpy(p, &hd, sizeof hd);
    p += sizeof hd;
    for (auto& it : state_items(c, c->block)) {
        CU(cudaMemcpy2D(p, it.row_bytes, it.ptr, it.pitch_bytes, it.row_bytes, it.rows, cudaMemcpyDeviceToHost));
        p += it.row_bytes * it.rows;
    }
    return SDRB_OK;
}

int sdrb_chain_state_load(sdrb_chain* c, const void* h_blob) { return sdrb_chain_state_load_n(c, h_blob, (size_t)-1); }

int sdrb_chain_state_load_n(sdrb_chain* c, const void* h_blob, size_t blob_bytes) {
    if (!c || !h_blob) return fail(SDRB_ERR_INVALID, "null argument");
    if (blob_bytes < sizeof(StateHeader)) return fail(SDRB_ERR_INVALID, "state blob is shorter than its header");
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaStreamSynchronize(c->stream));
    StateHeader hd;
    const char* p = static_cast<const char*>(h_blob);
    memcpy(&hd, p, sizeof hd);
    p += sizeof hd;
    if (hd.magic != kStateMagic || hd.n_streams != (uint32_t)c->S || hd.if_block != (uint32_t)c->info.if_block ||
        hd.type != (uint32_t)c->cfg.type)
        return fail(SDRB_ERR_INVALID, "state blob does not match this chain");
    {   // the size the blob claims, the size this chain writes for that block index, and the size the caller holds must agree
        const long long keep = c->block;
        c->block = hd.block;
        const size_t want = sdrb_chain_state_bytes(c);
        c->block = keep;
        if (hd.total_bytes != want || (blob_bytes != (size_t)-1 && blob_bytes < want))
            return fail(SDRB_ERR_INVALID, "state blob is truncated or was written by a different configuration");
    }
    c->block = hd.block;
    c->blocks_since_load = 0;
    c->poisoned = 0;  // a complete consistent state replaces whatever a failed block left behind
    for (auto& it : state_items(c, c->block)) {
        CU(cudaMemcpy2D(it.ptr, it.pitch_bytes, p, it.row_bytes, it.row_bytes, it.rows, cudaMemcpyHostToDevice));
        p += it.row_bytes * it.rows;
    }
    return SDRB_OK;
}

int sdrb_chain_set_profiling(sdrb_chain* c, int on) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    c->profiling = on != 0;
    if (c->profiling)
        for (auto& t : c->timed) t.n = 0;  // a new measurement window
    return SDRB_OK;
}

int sdrb_chain_kernel_times(sdrb_chain* c, const char** names, float* ms, int cap, int* n) {
    if (!c || !names || !ms || !n) return fail(SDRB_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaStreamSynchronize(c->stream));
    int k = 0;
    for (auto& t : c->timed) {
        if (t.n == 0 || k >= cap) continue;
        const int cnt = (int)(t.n < kTimedSlots ? t.n : kTimedSlots);
        double sum = 0;
        for (int i = 0; i < cnt; i++) {
            float v = 0;
            CU(cudaEventElapsedTime(&v, t.e0[i], t.e1[i]));
            sum += v;
        }
        names[k] = t.name;
        ms[k] = (float)(sum / cnt);
        k++;
    }
    *n = k;
    return SDRB_OK;
}

int sdrb_chain_sm_partition(const sdrb_chain* c, int sms[2]) {
    if (!c || !sms) return fail(SDRB_ERR_INVALID, "null argument");
    sms[0] = c->part_pll_sms;
    sms[1] = c->part_fir_sms;
    return SDRB_OK;
}

long long sdrb_chain_launch_count(const sdrb_chain* c) { return c ? c->launches : 0; }

// diagnostic builds (-DSDRB_PLL_DIAG): counts[2 * (1 + test) + loop], see pllmath.cuh
int sdrb_chain_check_guards(sdrb_chain* c, int* n_checked) {
    if (!c) return fail(SDRB_ERR_INVALID, "null argument");
    if (n_checked) *n_checked = 0;
    if (!c->guard) return fail(SDRB_ERR_STATE, "chain was created without SDRB_GUARD=1");
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaStreamSynchronize(c->stream));
    std::vector<unsigned char> lo(kGuardBytes), hi(kGuardBytes + 256);
    int idx = 0;
    for (auto& g : c->guards) {
        const size_t body = round_up(g.bytes, 256);
        const size_t tail = kGuardBytes + (body - g.bytes);  // the alignment slack behind the buffer is canary too
        CU(cudaMemcpy(lo.data(), g.base, kGuardBytes, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hi.data(), g.base + kGuardBytes + g.bytes, tail, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < kGuardBytes; i++)
            if (lo[i] != kGuardByte) return fail(SDRB_ERR_STATE, "guard zone BEFORE allocation #" + std::to_string(idx) + " (" + std::to_string(g.bytes) + " bytes) was overwritten");
        for (size_t i = 0; i < tail; i++)
            if (hi[i] != kGuardByte) return fail(SDRB_ERR_STATE, "guard zone AFTER allocation #" + std::to_string(idx) + " (" + std::to_string(g.bytes) + " bytes) was overwritten at +" + std::to_string(i));
        idx++;
    }
    if (n_checked) *n_checked = idx;
    return SDRB_OK;
}

int sdrb_chain_pll_redo_detail(sdrb_chain* c, unsigned long long counts[20]) {
    if (!c || !counts) return fail(SDRB_ERR_INVALID, "null argument");
    memset(counts, 0, 20 * sizeof(unsigned long long));
    if (!c->d_pll_redo) return SDRB_OK;
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaMemcpyAsync(counts, c->d_pll_redo, 20 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDRB_OK;
}

int sdrb_chain_pll_redos(sdrb_chain* c, unsigned long long counts[2]) {
    if (!c || !counts) return fail(SDRB_ERR_INVALID, "null argument");
    counts[0] = counts[1] = 0;
    if (!c->d_pll_redo) return SDRB_OK;
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaMemcpyAsync(counts, c->d_pll_redo, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDRB_OK;
}

int sdrb_chain_rds_overflows(sdrb_chain* c, unsigned int counts[3]) {
    if (!c || !counts) return fail(SDRB_ERR_INVALID, "null argument");
    counts[0] = counts[1] = counts[2] = 0;
    if (!c->d_rds_overflow) return SDRB_OK;
    CU(cudaSetDevice(c->cfg.device));
    if (int rcj = join_main(c)) return rcj;
    CU(cudaMemcpyAsync(counts, c->d_rds_overflow, 3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SDRB_OK;
}

// ---- stand-alone batched primitives ---------------------------------------------------------------
namespace {
// Taps of the stand-alone FIR calls live in a small per-process cache keyed by device and content: the reference-shaped
// callers pass the same host vector block after block (src/rffrontend.cpp:66-67, src/stereo.cpp:74-97, src/rds.cpp:105-133),
// so a device allocation, an upload and a free per call (round 1) were pure overhead.  An entry is uploaded once with a
// blocking copy and never changes afterwards, which makes it safe to read from any stream.
struct TapEntry {
    int dev;
    std::vector<float> host;
    float* d;
};
std::mutex g_tap_mutex;
std::vector<TapEntry> g_tap_cache;
constexpr size_t kTapCacheEntries = 64;

// *d: device taps; *owned: 1 if the caller has to cudaFreeAsync them on `st` (cache full: the round-1 path)
int upload_taps(const float* h_taps, int nh, float** d, cudaStream_t st, int* owned) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    *owned = 0;
    {
        std::lock_guard<std::mutex> lock(g_tap_mutex);
        for (const TapEntry& e : g_tap_cache)
            if (e.dev == dev && (int)e.host.size() == nh && memcmp(e.host.data(), h_taps, sizeof(float) * nh) == 0) {
                *d = e.d;
                return SDRB_OK;
            }
        if (g_tap_cache.size() < kTapCacheEntries) {
            float* p = nullptr;
            CU(cudaMalloc((void**)&p, sizeof(float) * nh));
            cudaError_t e = cudaMemcpy(p, h_taps, sizeof(float) * nh, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                cudaFree(p);
                return cuda_fail(e, "cudaMemcpy(taps)");
            }
            g_tap_cache.push_back(TapEntry{dev, std::vector<float>(h_taps, h_taps + nh), p});
            *d = p;
            return SDRB_OK;
        }
    }
    CU(cudaMallocAsync((void**)d, sizeof(float) * nh, st));
    CU(cudaMemcpyAsync(*d, h_taps, sizeof(float) * nh, cudaMemcpyHostToDevice, st));
    *owned = 1;
    return SDRB_OK;
}
}  // namespace

int sdrb_fir_decim(const float* d_x, size_t x_pitch, int nx, const float* h_taps, int nh, float* d_state, float* d_y,
                   size_t y_pitch, int decim, int n_streams, void* stream) {
    if (!d_x || !h_taps || !d_state || !d_y || nx < 0 || nh < 1 || decim < 1 || n_streams < 1)
        return fail(SDRB_ERR_INVALID, "bad argument");
    const int nstate = nh - 1;
    if (nx < nstate) return fail(SDRB_ERR_INVALID, "nx must be >= nh-1 (the reference reads out of bounds otherwise)");
    cudaStream_t st = (cudaStream_t)stream;
    float* d_h = nullptr;
    int owned = 0;
    int rc = upload_taps(h_taps, nh, &d_h, st, &owned);
    if (rc) return rc;
    const int ny = nx / decim;
    if (ny > 0) {
        dim3 grid((ny + 127) / 128, n_streams);
        k_fir_decim_generic<<<grid, 128, 0, st>>>(d_x, x_pitch, nx, d_h, nh, d_state, nstate, d_y, y_pitch, decim);
    }
    if (nstate > 0) {
        dim3 grid((nstate + 127) / 128, n_streams);
        k_state_update<<<grid, 128, 0, st>>>(d_x, x_pitch, nx, d_state, nstate);
    }
    CU(cudaGetLastError());
    if (owned) CU(cudaFreeAsync(d_h, st));
    return SDRB_OK;
}

int sdrb_fir_updown(const float* d_x, size_t x_pitch, int nx, const float* h_taps, int nh, float* d_state, int nstate,
                    float* d_y, size_t y_pitch, int up, int down, int n_streams, void* stream) {
    if (!d_x || !h_taps || !d_state || !d_y || nx < 0 || nh < 1 || up < 1 || down < 1 || n_streams < 1 || nstate < 0)
        return fail(SDRB_ERR_INVALID, "bad argument");
    if (nx < nstate) return fail(SDRB_ERR_INVALID, "nx must be >= nstate");
    if (nstate < (nh - 1) / up) return fail(SDRB_ERR_INVALID, "nstate must be >= (nh-1)/up");
    cudaStream_t st = (cudaStream_t)stream;
    float* d_h = nullptr;
    int owned = 0;
    int rc = upload_taps(h_taps, nh, &d_h, st, &owned);
    if (rc) return rc;
    const int ny = (int)(((long long)nx * up) / down);
    if (ny > 0) {
        dim3 grid((ny + 127) / 128, n_streams);
        k_fir_updown_generic<<<grid, 128, 0, st>>>(d_x, x_pitch, nx, d_h, nh, d_state, nstate, d_y, y_pitch, up, down);
    }
    if (nstate > 0) {
        dim3 grid((nstate + 127) / 128, n_streams);
        k_state_update<<<grid, 128, 0, st>>>(d_x, x_pitch, nx, d_state, nstate);
    }
    CU(cudaGetLastError());
    if (owned) CU(cudaFreeAsync(d_h, st));
    return SDRB_OK;
}

int sdrb_fm_demod(const float* d_I, const float* d_Q, size_t iq_pitch, int n, float* d_prev, float* d_out, size_t out_pitch,
                  int n_streams, void* stream) {
    if (!d_I || !d_Q || !d_prev || !d_out || n < 1 || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((n + 127) / 128, n_streams);
    k_fm_demod_generic<<<grid, 128, 0, st>>>(d_I, d_Q, iq_pitch, n, d_prev, d_out, out_pitch);
    k_fm_prev_update<<<(n_streams + 127) / 128, 128, 0, st>>>(d_I, d_Q, iq_pitch, n, d_prev, n_streams);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_pll(const float* d_in, size_t in_pitch, int n, float freq, float Fs, float ncoScale, float phaseAdjust,
             float normBandwidth, sdrb_pll_state* d_state, float* d_out, size_t out_pitch, int n_streams, void* stream) {
    if (!d_in || !d_state || !d_out || n < 0 || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    cr::PllCoef k = cr::pll_coef(freq, Fs, ncoScale, phaseAdjust, normBandwidth);
    k_pll_generic<<<(n_streams + kPllThreads - 1) / kPllThreads, kPllThreads, 0, st>>>(
        d_in, in_pitch, n, k, reinterpret_cast<PllStateAbi*>(d_state), d_out, out_pitch, n_streams);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_cdr(const float* d_x, size_t x_pitch, int n, int sps, int* d_offset, int n_streams, void* stream) {
    if (!d_x || !d_offset || n < 0 || sps < 1 || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    k_cdr_generic<<<n_streams, 64, 0, (cudaStream_t)stream>>>(d_x, x_pitch, n, sps, d_offset);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_manchester_decode(const int32_t* d_symbols, size_t sym_pitch, const int32_t* d_nsym, int block_count,
                           sdrb_manchester_state* d_state, int32_t* d_bits, size_t bits_pitch, int32_t* d_nbits, int n_streams, void* stream) {
    if (!d_symbols || !d_nsym || !d_state || !d_bits || !d_nbits || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    static_assert(sizeof(sdrb_manchester_state) == sizeof(ManchesterState), "ABI struct");
    k_manchester_generic<<<n_streams, 32, 0, (cudaStream_t)stream>>>(d_symbols, sym_pitch, d_nsym, block_count,
                                                                      reinterpret_cast<ManchesterState*>(d_state), d_bits, bits_pitch, d_nbits);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_differential_decode(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int block_num, int32_t* d_last_bit,
                             int32_t* d_decoded, size_t dec_pitch, int n_streams, void* stream) {
    if (!d_bits || !d_nbits || !d_last_bit || !d_decoded || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    k_differential_generic<<<n_streams, 32, 0, (cudaStream_t)stream>>>(d_bits, bits_pitch, d_nbits, block_num, d_last_bit, d_decoded, dec_pitch);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_frame_sync(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int max_nbits, sdrb_framesync_state* d_state,
                    uint64_t* d_groups, size_t groups_pitch, int32_t* d_ngroups, int max_groups, int n_streams, void* stream) {
    if (!d_bits || !d_nbits || !d_state || !d_groups || !d_ngroups || max_groups < 0 || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    if (max_nbits < 0 || max_nbits + 64 > kFrameSyncMaxBits) return fail(SDRB_ERR_INVALID, "sdrb_frame_sync: at most 8128 new bits per call");
    static_assert(sizeof(sdrb_framesync_state) == sizeof(FrameSyncState), "ABI struct");
    k_frame_sync_generic<<<n_streams, 32, 0, (cudaStream_t)stream>>>(d_bits, bits_pitch, d_nbits, reinterpret_cast<FrameSyncState*>(d_state),
                                                                      reinterpret_cast<unsigned long long*>(d_groups), groups_pitch, d_ngroups, max_groups);
    CU(cudaGetLastError());
    return SDRB_OK;
}

int sdrb_rds_sync(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int max_nbits, sdrb_rds_sync_state* d_state,
                  sdrb_rds_sync_event* d_events, size_t events_pitch, int32_t* d_nevents, int max_events, uint16_t* d_syndromes,
                  int n_streams, void* stream) {
    if (!d_bits || !d_nbits || !d_state || !d_events || !d_nevents || max_events < 0 || n_streams < 1) return fail(SDRB_ERR_INVALID, "bad argument");
    if (max_nbits < 0 || max_nbits + 32 > kFrameSyncMaxBits) return fail(SDRB_ERR_INVALID, "sdrb_rds_sync: at most 8160 bits per call");
    static_assert(sizeof(sdrb_rds_sync_state) == sizeof(RdsSyncState), "ABI struct");
    static_assert(sizeof(sdrb_rds_sync_event) == sizeof(RdsSyncEvent), "ABI struct");
    k_rds_sync_generic<<<n_streams, 32, 0, (cudaStream_t)stream>>>(d_bits, bits_pitch, d_nbits, reinterpret_cast<RdsSyncState*>(d_state),
                                                                    reinterpret_cast<RdsSyncEvent*>(d_events), events_pitch, d_nevents, max_events,
                                                                    reinterpret_cast<unsigned short*>(d_syndromes));
    CU(cudaGetLastError());
    return SDRB_OK;
}

}  // extern "C"
