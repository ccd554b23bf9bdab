// Device kernels of the B200 receive chain (sm_100a).  Host orchestration and the C ABI live in
// sdr_chain.cu; the correctly-rounded PLL math is pllmath.cuh.
//
// Arithmetic contract: every FIR accumulates  acc = RN(acc + RN(h[k]*x[n-k]))  for
// k = 0..K-1 in that order, starting from +0, exactly like the reference's scalar loops
// (/root/reference/src/filter.cpp:106-147, compiled without FMA).  All float arithmetic that feeds a
// result goes through __fmul_rn/__fadd_rn (never contracted) and the file is compiled with -fmad=false.
//
// Data layout: every intermediate is a "ring" of NRING slots, each [n_streams][pitch] floats, where a
// stream row is  [halo | block]  : index -halo..-1 holds the tail of the previous block (what the
// reference keeps in its `state` vectors), 0..n-1 the current block.  A producer writes block b into
// slot b%NRING and the tail of that block into the halo of slot (b+1)%NRING, so a consumer always
// reads one contiguous row and block b+1 can be produced while block b is still being consumed.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pllmath.cuh"
#include "res_lanes.h"

namespace sdrb {

constexpr int kTaps = 101;     // rf_taps, /root/reference/src/project.cpp:34
constexpr int kState = 100;    // kTaps-1 samples of carried FIR state
constexpr int kNRing = 3;

struct Taps101 {
    float h[kTaps];
};

// One ring (see the layout note above).  `cur` points at sample 0 of stream 0 in the slot being
// produced/consumed; `nxt` at sample 0 of the following slot (its halo receives the tail).
struct RingView {
    float* cur;
    float* nxt;
    size_t pitch;  // floats per stream row (halo included)
    int halo;
    int n;         // samples per block
};

__device__ __forceinline__ void ring_store(const RingView& r, int s, int i, float v) {
    r.cur[(size_t)s * r.pitch + i] = v;
    int back = r.n - i;  // 1..halo for the tail
    if (back <= r.halo) r.nxt[(size_t)s * r.pitch - back] = v;
}

// ------------------------------------------------------------------------------------------------
// FIR cores.  A thread owns R consecutive outputs and slides a register window over the input, so a
// tap step costs one shared-memory load per input phase instead of one per MAC.
//
// Shared-memory layout: logical sample u (u = 0 is the oldest sample the tile needs) lives at
// P(u) = u + u / L with L = D*R (the lane stride), which makes the lane stride L+1 (odd) words and all
// window loads bank-conflict free.  With u = L*lane + c (c a compile-time constant) this is
// (L+1)*lane + c + c/L, i.e. an immediate offset from a per-lane base.
// ------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ int pad_pos(int u) {
    return u + u / L;
}

__device__ __forceinline__ float mac(float acc, float h, float x) { return __fadd_rn(acc, __fmul_rn(h, x)); }
// Two channels at once on Blackwell's packed FP32 pipe (FFMA2 / FADD2), keeping the two roundings of the bit-exact MAC.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 (one rounding) even under -fmad=false and even when the
// multiply is spelled fma(h, x, -0) with a literal -0, so the product is formed as fma(h, x, nz) with nz = (-0, -0) read
// from a run-time constant the compiler cannot fold: RN(h*x - 0) = RN(h*x) exactly (and the sign of a zero product is
// kept), followed by a separate FADD2.  Two instructions per two MACs instead of four.
__device__ __constant__ unsigned long long c_neg_zero2 = 0x8000000080000000ull;

__device__ __forceinline__ float2 mac(float2 acc, float h, float2 x) {
    unsigned long long a, xx, hh, p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(hh), "l"(xx), "l"(c_neg_zero2));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(a), "l"(p));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(a));
    return r;
}

// y[m] = sum_k h[k] * x[D*m - k]   for the R outputs m = R*lane + j of this lane.
// sx_lane = tile base + (L+1)*lane; logical sample index of x[D*m - k] inside the tile is
// u = D*(R*lane + j) - k + kState.
template <int D, int R, typename V>
__device__ __forceinline__ void fir_core(const V* __restrict__ sx_lane, const Taps101& t, V (&acc)[R]) {
    constexpr int L = D * R;
    constexpr int A = (kTaps + D - 1) / D;
    V w[D][R];
#pragma unroll
    for (int j = 0; j < R; j++)
#pragma unroll
        for (int b = 0; b < D; b++) {
            const int c = D * j - b + kState;
            w[b][j] = sx_lane[c + c / L];
        }
#pragma unroll
    for (int a = 0; a < A; a++) {
        if (a > 0) {
            const int slot = ((-a) % R + R) % R;
#pragma unroll
            for (int b = 0; b < D; b++) {
                if (D * a + b < kTaps) {
                    const int c = -D * a - b + kState;
                    w[b][slot] = sx_lane[c + c / L];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < D; b++) {
            const int k = D * a + b;
            if (k < kTaps) {
                const float hk = t.h[k];
#pragma unroll
                for (int j = 0; j < R; j++) acc[j] = mac(acc[j], hk, w[b][((j - a) % R + R) % R]);
            }
        }
    }
}

// NF filters over the same input, D = 1, two adjacent outputs per packed accumulator.
//
// Output pair p of a lane is (y[R*lane + 2p], y[R*lane + 2p + 1]); tap k needs the input pair that starts at
// x[R*lane + 2p - k].  For even k that is an even-aligned pair, for odd k an odd-aligned one, so the tile is kept
// twice in shared memory: E[m] = (x[2m], x[2m+1]) and O[m] = (x[2m+1], x[2m+2]) (u = 0 is the oldest sample,
// kState before the first output).  Each alignment has its own sliding window of R/2 register pairs that moves by
// one pair every second tap: two LDS.64 per two taps, all MACs packed (see mac(float2, ...)).  The tap loop is a
// real loop (period R/2 in q = k/2, the window slots rotate by compile-time indices inside the body): fully
// unrolled, the 3-filter kernel was 80 KB of code and spent half its cycles waiting for instructions.
// Pair m lives at position m + m/(R/2): lane stride R/2+1 pairs, conflict free for LDS.64.
template <int NF, int R>
__device__ __forceinline__ void fir_bank_core(const float2* __restrict__ sE_lane, const float2* __restrict__ sO_lane,
                                              const Taps101* t, float2 (&acc)[NF][R / 2]) {
    constexpr int HP = R / 2;
    constexpr int HS = kState / 2;
    constexpr int NQ = kTaps / 2 + 1;  // q = 0 .. 50: taps 2q and 2q+1
    static_assert(R % 2 == 0 && kState % 2 == 0, "pairs");
    float2 wE[HP], wO[HP];
#pragma unroll
    for (int p = 0; p < HP; p++) {
        const int cE = HS + p, cO = HS - 1 + p;
        wE[p] = sE_lane[cE + cE / HP];
        wO[p] = sO_lane[cO + cO / HP];
    }
#pragma unroll 1
    for (int q0 = 0; q0 < NQ; q0 += HP) {
#pragma unroll
        for (int qq = 0; qq < HP; qq++) {
            const int q = q0 + qq;
            if (2 * q < kTaps) {
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    const float hk = t[f].h[2 * q];
#pragma unroll
                    for (int p = 0; p < HP; p++) acc[f][p] = mac(acc[f][p], hk, wE[((p - qq) % HP + HP) % HP]);
                }
            }
            if (2 * q + 1 < kTaps) {
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    const float hk = t[f].h[2 * q + 1];
#pragma unroll
                    for (int p = 0; p < HP; p++) acc[f][p] = mac(acc[f][p], hk, wO[((p - qq) % HP + HP) % HP]);
                }
            }
            // windows for q+1: one new (oldest) pair each, into the slot the newest pair just left
            const int slot = ((-(qq + 1)) % HP + HP) % HP;
            const int cE = HS - (q + 1), cO = HS - 2 - q;
            if (cE >= 0) wE[slot] = sE_lane[cE + cE / HP];
            if (cO >= 0) wO[slot] = sO_lane[cO + cO / HP];
        }
    }
}

// Scalar form of the bank core (one output per accumulator, fully unrolled): best for a single filter, where the
// packed form has too few independent accumulators per thread; its code stays small enough for the I-cache (NF = 1).
template <int NF, int R>
__device__ __forceinline__ void fir_bank_core_scalar(const float* __restrict__ sx_lane, const Taps101* t, float (&acc)[NF][R]) {
    constexpr int L = R;
    float w[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int c = j + kState;
        w[j] = sx_lane[c + c / L];
    }
#pragma unroll
    for (int k = 0; k < kTaps; k++) {
        if (k > 0) {
            const int c = -k + kState;
            w[((-k) % R + R) % R] = sx_lane[c + c / L];
        }
#pragma unroll
        for (int f = 0; f < NF; f++) {
            const float hk = t[f].h[k];
#pragma unroll
            for (int j = 0; j < R; j++) acc[f][j] = mac(acc[f][j], hk, w[((j - k) % R + R) % R]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1  RF front-end: u8 IQ -> unpack -> 101-tap LPF / DECIM on I and Q -> FM discriminator.
// /root/reference/src/rffrontend.cpp:58-71, src/filter.cpp:106-121, src/demod.cpp:3-24.
//
// One CTA = one stream x one tile.  The tile computes kRfTile FIR outputs; the first one only serves
// as the discriminator's "previous" sample, so tiles advance by kRfTile-1 demodulated samples and the
// reference's carried prev_I/prev_Q never has to be stored: it is recomputed from the carried input
// halo (110 IQ pairs, kept as raw bytes; 128 = the byte that unpacks to 0.0f, the initial state).
// ------------------------------------------------------------------------------------------------
#if !defined(SDRB_RF_R)
#define SDRB_RF_R 4
#define SDRB_RF_THREADS 64
#endif
constexpr int kRfR = SDRB_RF_R;            // FIR outputs per thread (tools/ab_run.sh variants: -DSDRB_RF_R=2 -DSDRB_RF_THREADS=128)
constexpr int kRfThreads = SDRB_RF_THREADS;
constexpr int kRfTile = kRfR * kRfThreads;  // 256 FIR outputs, 255 discriminator outputs
constexpr int kIqHaloPairs = 112;           // >= kState + max DECIM; 224 bytes, keeps 16-byte alignment

struct RfArgs {
    const uint8_t* iq;       // [n_streams][iq_pitch] bytes of the current block
    size_t iq_pitch;
    const uint8_t* halo_in;  // [n_streams][2*kIqHaloPairs] last pairs of the previous block
    uint8_t* halo_out;       // same, for the next block (ping-pong)
    int block_pairs;
    int if_block;            // block_pairs / DECIM
    RingView fm;             // fm_demod
    float* i_ds;             // optional [n_streams][if_block] (parity tests)
    float* q_ds;
    int tma_ok;              // iq base and pitch are 16-byte aligned and every row has round_up(block_bytes, 16) readable bytes
    int row_bytes16;         // round_up(2*block_pairs, 16) when tma_ok, else 2*block_pairs
};

// (u8 - 128)/128 exactly: 0x4B000000|u is the float 2^23 + u; *2^-7 - 65537 is exact in one FMA.
__device__ __forceinline__ float unpack_u8(uint32_t u) {
    return __fmaf_rn(__uint_as_float(0x4B000000u | u), 0.0078125f, -65537.0f);
}

// IQ pair `WHICH` (0 or 1) of a 32-bit word of raw bytes I0 Q0 I1 Q1 -> (float I, float Q): two byte permutes build
// 2^23 + u for both bytes, one packed FMA scales and recentres both (exact, as in unpack_u8).
template <int WHICH>
__device__ __forceinline__ float2 unpack_iq(uint32_t word) {
    const uint32_t i = __byte_perm(word, 0x4B000000u, WHICH ? 0x7652u : 0x7650u);
    const uint32_t q = __byte_perm(word, 0x4B000000u, WHICH ? 0x7653u : 0x7651u);
    unsigned long long v, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(i), "r"(q));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(v), "l"(0x3C0000003C000000ull), "l"(0xC7800080C7800080ull));
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
}

// (float)(num / (double)I^2 + (double)Q^2), /root/reference/src/demod.cpp:9-18
// The division: single-precision reciprocal + one double correction, accepted when the float rounding is certain
// (cr::fm_quotient_fast, pllmath.cuh; checked against the exact form on the host, tests/test_pllmath.py); the exact double
// division only for the rest (a tie window of 8e-6, and denominators or quotients outside the float-normal range).
__device__ __noinline__ float fm_divide_exact(float num, double den) { return __double2float_rn(__ddiv_rn((double)num, den)); }
__device__ __forceinline__ float fm_discriminate(float I, float Q, float pI, float pQ) {
    if (I == 0.0f && Q == 0.0f) return 0.0f;
    float num = __fadd_rn(__fmul_rn(I, __fadd_rn(Q, -pQ)), -__fmul_rn(Q, __fadd_rn(I, -pI)));
    double den = __dadd_rn(__dmul_rn((double)I, (double)I), __dmul_rn((double)Q, (double)Q));
#if !defined(SDRB_FM_EXACT_DIVIDE)
    if (cr::fm_den_in_range(den)) {
        float rc, out;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(cr::fm_den_approx_f(den)));
        if (cr::fm_quotient_fast(num, den, rc, out)) return out;
    }
    return fm_divide_exact(num, den);
#else
    return __double2float_rn(__ddiv_rn((double)num, den));
#endif
}

// ---- TMA (bulk async copy) helpers: global -> shared, completion on an mbarrier -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the async proxy must see the initialised barrier
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int DECIM>
__global__ void __launch_bounds__(kRfThreads) k_rf_frontend(const __grid_constant__ Taps101 taps, const RfArgs a) {
    constexpr int L = DECIM * kRfR;
    constexpr int NS = DECIM * (kRfTile - 1) + kTaps;  // input pairs a tile needs
    constexpr int kRawBytes = (2 * NS + 15 + 16) / 16 * 16;  // the tile's bytes, widened to 16-byte boundaries on both sides
    constexpr int kRawAlloc = kRawBytes > (int)(sizeof(float2) * kRfTile) ? kRawBytes : (int)(sizeof(float2) * kRfTile);
    __shared__ float2 sx[NS + NS / L + 2];
    __shared__ __align__(16) uint8_t raw[kRawAlloc];  // later reused for the FIR outputs (sy)
    __shared__ __align__(8) unsigned long long bar;
    float2* sy = reinterpret_cast<float2*>(raw);  // FIR outputs; the raw bytes are dead by then
    const int s = blockIdx.y;
    const int tile = blockIdx.x;
    const int m0 = tile * (kRfTile - 1) - 1;  // first FIR output of the tile (may be -1)
    const int g0 = DECIM * m0 - kState;       // first input pair (negative: halo)
    const uint8_t* blk = a.iq + (size_t)s * a.iq_pitch;
    const uint8_t* hal = a.halo_in + (size_t)s * (2 * kIqHaloPairs);
    // Byte window [A, A + span) of the stream row (negative = carried halo), A a multiple of 16.
    const int byte_lo = 2 * g0;
    const int A = (byte_lo >= 0) ? (byte_lo & ~15) : -((-byte_lo + 15) & ~15);
    const int byte_hi = min(2 * (g0 + NS), a.row_bytes16);  // never past the 16-byte-rounded end of the block
    if (a.tma_ok) {
        // Staging by TMA: one thread issues the bulk copies (carried halo bytes, then the block bytes), the CTA sleeps
        // on the mbarrier; no registers or issue slots are spent on the 5.3 KB of input.
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int lo = max(A, 0);
            const uint32_t nhalo = A < 0 ? (uint32_t)(-A) : 0u;
            const uint32_t nblk = byte_hi > lo ? (uint32_t)((byte_hi - lo + 15) & ~15) : 0u;
            mbar_expect_tx(&bar, nhalo + nblk);
            if (nhalo) tma_load_1d(raw, hal + (2 * kIqHaloPairs + A), nhalo, &bar);
            if (nblk) tma_load_1d(raw + (lo - A), blk + lo, nblk, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        // unaligned caller buffer: plain 16-bit loads into the same layout
        for (int o = 2 * threadIdx.x; o < kRawBytes; o += 2 * kRfThreads) {
            const int bo = A + o;
            uint16_t v = 0x8080u;
            if (bo < 0) { if (bo >= -2 * kIqHaloPairs) v = *reinterpret_cast<const uint16_t*>(hal + 2 * kIqHaloPairs + bo); }
            else if (bo < 2 * a.block_pairs) v = *reinterpret_cast<const uint16_t*>(blk + bo);
            *reinterpret_cast<uint16_t*>(raw + o) = v;
        }
        __syncthreads();
    }
    // unpack to float2 (I, Q) in the padded, bank-conflict-free layout the FIR core reads
    if (((byte_lo - A) & 3) == 0) {
        // Word path (always taken for even DECIM): one 32-bit word = two IQ pairs -> four PRMT, two packed FMAs, two
        // 64-bit stores (6.5 instructions per pair; the pair-at-a-time loop below is 17).  Bytes past the end of the
        // block (last tile only) are first overwritten with 128, the byte that unpacks to 0.0f.
        if (a.tma_ok && g0 + NS > a.block_pairs) {
            for (int o = 2 * a.block_pairs - A + 2 * (int)threadIdx.x; o < kRawBytes; o += 2 * kRfThreads)
                *reinterpret_cast<uint16_t*>(raw + o) = 0x8080u;
            __syncthreads();
        }
        // A sweep covers a whole number of padding blocks (L pairs = L/2 words each), so from one sweep to the next
        // both addresses advance by compile-time constants: no index arithmetic in the (unrolled) loop.
        constexpr int NW = (NS + 1) / 2;               // words in the tile
        constexpr int HW = L / 2;                       // words per padding block (L is even)
        constexpr int SW = (kRfThreads / HW) * HW;      // words per sweep (threads >= SW sit this phase out)
        constexpr int SP = 2 * SW + 2 * SW / L;         // float2 slots per sweep
        if (threadIdx.x < SW) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw) + ((byte_lo - A) >> 2) + threadIdx.x;
            float2* d = &sx[pad_pos<L>(2 * threadIdx.x)];
#pragma unroll
            for (int i = 0; i < (NW + SW - 1) / SW; i++) {
                if ((i + 1) * SW <= NW || (int)threadIdx.x + i * SW < NW) {
                    const uint32_t word = rw[i * SW];
                    d[i * SP] = unpack_iq<0>(word);
                    d[i * SP + 1] = unpack_iq<1>(word);  // (for odd NS the last one lands in the spare slot past the tile)
                }
            }
        }
    } else {
        for (int u = threadIdx.x; u < NS; u += kRfThreads) {
            const int g = g0 + u;
            uint32_t pr = 0x8080u;  // beyond the block: the byte that unpacks to 0.0f
            if (g < a.block_pairs) pr = *reinterpret_cast<const uint16_t*>(raw + (2 * g - A));
            sx[pad_pos<L>(u)] = make_float2(unpack_u8(pr & 0xFFu), unpack_u8(pr >> 8));
        }
    }
    // carry the last pairs of this block to the next block's halo (one tile per stream does it)
    if (tile == 0)
        for (int i = threadIdx.x; i < kIqHaloPairs; i += kRfThreads)
            reinterpret_cast<uint16_t*>(a.halo_out + (size_t)s * (2 * kIqHaloPairs))[i] =
                *reinterpret_cast<const uint16_t*>(blk + 2 * (size_t)(a.block_pairs - kIqHaloPairs + i));
    __syncthreads();
    float2 acc[kRfR];
#pragma unroll
    for (int j = 0; j < kRfR; j++) acc[j] = make_float2(0.0f, 0.0f);
    fir_core<DECIM, kRfR, float2>(sx + (L + 1) * threadIdx.x, taps, acc);
#pragma unroll
    for (int j = 0; j < kRfR; j++) sy[kRfR * threadIdx.x + j] = acc[j];
    __syncthreads();
    float* const fm_row = a.fm.cur + (size_t)s * a.fm.pitch;
    float* const fm_halo = a.fm.nxt + (size_t)s * a.fm.pitch - a.fm.n;  // fm_halo[m] = slot of sample m in the next block's halo
    const int halo_from = a.fm.n - a.fm.halo;
#pragma unroll
    for (int i = 0; i < kRfR; i++) {
        const int q = threadIdx.x + 1 + i * kRfThreads;
        const int m = m0 + q;
        if (q < kRfTile && m < a.if_block) {
            const float2 c = sy[q], p = sy[q - 1];
            const float v = fm_discriminate(c.x, c.y, p.x, p.y);
            fm_row[m] = v;
            if (m >= halo_from) fm_halo[m] = v;
            if (a.i_ds) {
                a.i_ds[(size_t)s * a.if_block + m] = c.x;
                a.q_ds[(size_t)s * a.if_block + m] = c.y;
            }
        }
    }
}

// (A warp-specialised, persistent form of K1 - a producer warp keeping a bulk copy in flight and unpacking tile t+1 while
// two consumer warps filter tile t, double-buffered float2 tiles, mbarrier hand-off - was built in round 2 and is
// bit-exact, but slower: 0.205-0.248 ms against 0.167 for 1024 stations with one or two producer warps and 2-15 tiles per
// CTA (profiles/README.md).  Its 53 KB of shared memory per CTA leave two consumer warps per scheduler, and a MAC loop needs
// more resident warps than that to keep the FMA pipe busy; the one-tile-per-CTA kernel above has four.)

// ------------------------------------------------------------------------------------------------
// K2  IF band filters: NF 101-tap FIRs (decim 1) over one input ring, e.g. pilot / stereo / RDS
// band-pass over fm_demod (/root/reference/src/stereo.cpp:74,80, src/rds.cpp:105), the 114 kHz
// band-pass over rds_band^2 (src/rds.cpp:111-116, SQUARE) and the RRC.  One warp = one tile of 256
// outputs of one stream.
// ------------------------------------------------------------------------------------------------
constexpr int kBankR = 8;
constexpr int kBankTile = 32 * kBankR;
constexpr int kBankWarps = 4;

template <int NF>
struct BankArgs {
    const float* x;  // input ring slot, sample 0 of stream 0 (halo >= kState)
    size_t x_pitch;
    int n;           // samples per block
    int tiles;       // tiles per stream
    int n_streams;
    RingView y[NF];
    Taps101 taps[NF];
};

template <int NF, bool SQUARE>
__global__ void __launch_bounds__(32 * kBankWarps) k_fir_bank(const __grid_constant__ BankArgs<NF> a) {
    constexpr int NS = kBankTile + kState;  // samples a tile needs
    constexpr int HP = kBankR / 2;
    constexpr int NP = NS / 2 + 1;          // pairs per alignment
    __shared__ float2 sE[kBankWarps][NP + NP / HP + 1];
    __shared__ float2 sO[kBankWarps][NP + NP / HP + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kBankWarps + warp;
    const int s = (int)(wid / a.tiles);
    const int tile = (int)(wid % a.tiles);
    if (s >= a.n_streams) return;
    const int n0 = tile * kBankTile;
    const float* xr = a.x + (size_t)s * a.x_pitch + n0 - kState;
    float* fE = reinterpret_cast<float*>(sE[warp]);
    float* fO = reinterpret_cast<float*>(sO[warp]);
    // Staging.  Inside the block (every tile but the last of a row) a lane loads one even-aligned pair per step, which
    // is E[m] as it stands, and gets x[2m+2] for O[m] from its neighbour by shuffle: 5 instructions per sample instead
    // of 17 for the element-wise form below (bounds test, two index computations with a division, two scalar stores).
    if (n0 + kBankTile + 2 <= a.n && (((uintptr_t)xr & 7u) == 0)) {
        static_assert(NS % 2 == 0, "pairs");
        const float2* xp = reinterpret_cast<const float2*>(xr);
        for (int m0 = 0; m0 < NP; m0 += 32) {
            const int m = m0 + lane;
            float2 v = make_float2(0.0f, 0.0f);
            if (m < NP) v = xp[m];  // m = NP-1 reads x[NS], x[NS+1]: inside the block by the test above
            if (SQUARE) v = make_float2(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
            float nx = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
            if (lane == 31 && m + 1 < NP) { nx = xr[2 * m + 2]; if (SQUARE) nx = __fmul_rn(nx, nx); }
            if (m < NP) {
                const int pos = m + m / HP;
                sE[warp][pos] = v;
                sO[warp][pos] = make_float2(v.y, nx);
            }
        }
    } else
    for (int u = lane; u < NS + 2; u += 32) {
        float v = (u < NS && n0 - kState + u < a.n) ? xr[u] : 0.0f;
        if (SQUARE) v = __fmul_rn(v, v);
        const int mE = u >> 1;
        if (mE < NP) fE[2 * (mE + mE / HP) + (u & 1)] = v;          // E[m] = (x[2m], x[2m+1])
        if (u >= 1) {
            const int mO = (u - 1) >> 1;
            if (mO < NP) fO[2 * (mO + mO / HP) + ((u - 1) & 1)] = v;  // O[m] = (x[2m+1], x[2m+2])
        }
    }
    __syncwarp();
    float2 acc[NF][HP];
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
        for (int p = 0; p < HP; p++) acc[f][p] = make_float2(0.0f, 0.0f);
    fir_bank_core<NF, kBankR>(&sE[warp][(HP + 1) * lane], &sO[warp][(HP + 1) * lane], a.taps, acc);
    // A tile that ends before the part of the block that is also the next block's halo (all but the last one or two
    // tiles of a row) stores pairs directly: no bounds test, no halo test, one 64-bit store per pair.
    int hmax = 0;
#pragma unroll
    for (int f = 0; f < NF; f++) hmax = max(hmax, a.y[f].halo);
    bool fast = n0 + kBankTile <= a.n - hmax;
#pragma unroll
    for (int f = 0; f < NF; f++) fast = fast && (((uintptr_t)(a.y[f].cur + (size_t)s * a.y[f].pitch + n0) & 7u) == 0);
    if (fast) {
#pragma unroll
        for (int f = 0; f < NF; f++) {
            float2* row = reinterpret_cast<float2*>(a.y[f].cur + (size_t)s * a.y[f].pitch + n0 + kBankR * lane);
#pragma unroll
            for (int p = 0; p < HP; p++) row[p] = acc[f][p];
        }
        return;
    }
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
        for (int p = 0; p < HP; p++) {
            const int n = n0 + kBankR * lane + 2 * p;
            if (n < a.n) ring_store(a.y[f], s, n, acc[f][p].x);
            if (n + 1 < a.n) ring_store(a.y[f], s, n + 1, acc[f][p].y);
        }
}

template <int NF, bool SQUARE>
__global__ void __launch_bounds__(32 * kBankWarps) k_fir_bank_scalar(const __grid_constant__ BankArgs<NF> a) {
    constexpr int NS = kBankTile + kState;
    __shared__ float sx[kBankWarps][NS + NS / kBankR + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kBankWarps + warp;
    const int s = (int)(wid / a.tiles);
    const int tile = (int)(wid % a.tiles);
    if (s >= a.n_streams) return;
    const int n0 = tile * kBankTile;
    const float* xr = a.x + (size_t)s * a.x_pitch + n0 - kState;
    for (int u = lane; u < NS; u += 32) {
        float v = (n0 - kState + u < a.n) ? xr[u] : 0.0f;
        if (SQUARE) v = __fmul_rn(v, v);
        sx[warp][pad_pos<kBankR>(u)] = v;
    }
    __syncwarp();
    float acc[NF][kBankR];
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
        for (int j = 0; j < kBankR; j++) acc[f][j] = 0.0f;
    fir_bank_core_scalar<NF, kBankR>(&sx[warp][(kBankR + 1) * lane], a.taps, acc);
    int hmax = 0;
#pragma unroll
    for (int f = 0; f < NF; f++) hmax = max(hmax, a.y[f].halo);
    bool fast = n0 + kBankTile <= a.n - hmax;  // see k_fir_bank
#pragma unroll
    for (int f = 0; f < NF; f++) fast = fast && (((uintptr_t)(a.y[f].cur + (size_t)s * a.y[f].pitch + n0) & 15u) == 0);
    if (fast) {
        static_assert(kBankR % 4 == 0, "float4 stores");
#pragma unroll
        for (int f = 0; f < NF; f++) {
            float4* row = reinterpret_cast<float4*>(a.y[f].cur + (size_t)s * a.y[f].pitch + n0 + kBankR * lane);
#pragma unroll
            for (int j = 0; j < kBankR; j += 4) row[j / 4] = make_float4(acc[f][j], acc[f][j + 1], acc[f][j + 2], acc[f][j + 3]);
        }
        return;
    }
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
        for (int j = 0; j < kBankR; j++) {
            int n = n0 + kBankR * lane + j;
            if (n < a.n) ring_store(a.y[f], s, n, acc[f][j]);
        }
}

// ------------------------------------------------------------------------------------------------
// K3  PLL: one lane per stream, strictly sequential in time.  /root/reference/src/pll.cpp:4-61.
// The NCO output cos(trigArg*ncoScale + phaseAdjust) (:52) is not part of the recurrence, so the
// kernel only stores trigArg; the consumers (k_mix, or the standalone ABI call) evaluate the cosine
// in parallel.  blockIdx.y selects one of two independent loops (19 kHz pilot, 114 kHz RDS carrier).
// ------------------------------------------------------------------------------------------------
struct PllStateDev {
    float feedbackI, feedbackQ, integrator, phaseEst;
    double trigOffset;
};

struct PllLoop {
    const float* x;  // input slot, sample 0 of stream 0
    size_t x_pitch;
    RingView trig;   // trig.cur[i] = trigArg after input sample i
    PllStateDev* st; // [n_streams]
    cr::PllCoef coef;
    unsigned long long* redo;  // health counter: lane-chunks (4 samples of one station) that took the careful path
};

struct PllArgs {
    PllLoop loop[2];
    int n;
    int n_streams;
};

// CTA size of k_pll: the launch uses the smallest CTA that keeps the PLL on at most `pll_max_ctas` SMs (sdr_chain.cu:
// 32 by default, i.e. two warps per SM at 1024 stereo+RDS stations) and leaves the rest to the FIR kernels.  Alone, one
// warp per SM is 1 % faster than two and 4 % faster than four (the warps of an SM share its instruction and constant
// caches), but the 32 SMs that costs the FIR side are worth more (profiles/README.md).  kPllMaxCtas bounds the knob.
constexpr int kPllThreads = 128;  // largest CTA; also the block size of the generic single-call kernel
constexpr int kPllMaxCtas = 64;

// 1/v for the rotated phase detector: w = u * (1/in) is a <= 2^-22 rad correction whose absolute error may be 2^-45,
// so a relative 2^-23 is enough: the FP32 reciprocal approximation (MUFU.RCP, 1 ulp) widened to double, unguarded:
// the speculative step itself rejects samples outside 2^-90 <= |in| < 2^90, and the careful path guards its own.
__device__ __forceinline__ double pll_recip(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(v)));
    return (double)r;
}

constexpr int kPllTileChunks = 8;  // 32 steps per staged tile

constexpr int kPllTileRow = 4 * kPllTileChunks + 4;
#ifndef SDRB_PLL_RING
#define SDRB_PLL_RING 3
#endif
#ifndef SDRB_PLL_CHUNK_UNROLL
#define SDRB_PLL_CHUNK_UNROLL 2  // chunks of four samples per loop iteration
#endif
constexpr int kPllChunkUnroll = SDRB_PLL_CHUNK_UNROLL;
constexpr int kPllRing = SDRB_PLL_RING;  // tiles in the input ring; kPllRing - 1 of them are in flight ahead of the loop
constexpr size_t pll_tile_bytes(int threads) { return sizeof(float) * kPllRing * threads * kPllTileRow; }
// The launch asks for (nearly) all of an SM's shared memory, far more than the input ring needs: no other CTA then fits
// on that SM, so in overlap mode the FIR kernels of the neighbouring blocks cannot steal issue slots from the
// latency-bound warps.
#ifndef SDRB_PLL_SMEM_KB
#define SDRB_PLL_SMEM_KB 225
#endif
constexpr size_t kPllSmemBytes = SDRB_PLL_SMEM_KB * 1024;

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_pll(const PllArgs a) {
    __shared__ cr::AtanTab tab;
    extern __shared__ __align__(16) float pll_dyn_smem[];
    float (*tile)[THREADS][kPllTileRow] = reinterpret_cast<float (*)[THREADS][kPllTileRow]>(pll_dyn_smem);
    {
        const cr::AtanTab init = SDRB_ATAN_TAB_INIT;
        if (threadIdx.x < 17) {
            tab.hi[threadIdx.x] = init.hi[threadIdx.x];
            tab.lo[threadIdx.x] = init.lo[threadIdx.x];
        }
    }
    __syncthreads();
    const PllLoop& lp = a.loop[blockIdx.y];
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= a.n_streams) return;
#if defined(SDRB_PLL_TIMESTAMPS)
    unsigned long long ts0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts0));
#endif
    PllStateDev sd = lp.st[s];
    cr::PllState st{sd.feedbackI, sd.feedbackQ, sd.integrator, sd.phaseEst, sd.trigOffset};
    const cr::PllCoef k = lp.coef;
    cr::PllFast f;
    cr::pll_fast_load(f, st, k);
    const float* x = lp.x + (size_t)s * lp.x_pitch;
    float* out = lp.trig.cur + (size_t)s * lp.trig.pitch;
    const int n4 = a.n & ~3;
    // Input staging: a warp issues in order, so a global load anywhere near its consumer stalls the whole
    // recurrence for the full memory latency.  The input therefore travels through a 3-tile shared-memory ring
    // filled with cp.async two tiles (64 steps, ~30 us) ahead; the loop only ever waits on shared memory.
    // Row stride 36 floats keeps the per-lane LDS.128 conflict free and the 16-byte cp.async chunks aligned.
    const int lane = threadIdx.x;
    const int nc = n4 >> 2;                          // chunks of 4 steps
    const int ntiles = (nc + kPllTileChunks - 1) / kPllTileChunks;
    auto issue_tile = [&](int t) {
        if (t < ntiles) {
            const float* src = x + t * (4 * kPllTileChunks);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[t % kPllRing][lane][0]);
#pragma unroll
            for (int j = 0; j < kPllTileChunks; j++)
                if (t * kPllTileChunks + j < nc)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * j), "l"(src + 4 * j) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int t0 = 0; t0 < kPllRing - 1; t0++) issue_tile(t0);
    float4 vc = make_float4(1.f, 1.f, 1.f, 1.f);
    double q0 = 1.0, q1 = 1.0, q2 = 1.0, q3 = 1.0;
    float4* o4 = reinterpret_cast<float4*>(out);
    cr::PllK kk;
#if defined(SDRB_PLL_UNROTATED)
    cr::pll_k_load(kk);
#else
    cr::pll_k_load_lean(kk);
#endif
    cr::PllHead hd{0.0, 0u, 0u};
    [[maybe_unused]] float4 op = make_float4(0.f, 0.f, 0.f, 0.f);  // NCO phases of the last accepted chunk, not stored yet
    [[maybe_unused]] bool pending = false;
    // Loop shape (measured, profiles/README.md): tiles outside and chunks inside, so the common path has no branch
    // around the tile staging (a taken branch costs ~16 cycles of instruction fetch on a warp that has its scheduler to
    // itself); the chunk loop unrolled by two (half the loop branches and state-rotation moves; by four or more the
    // body outgrows the instruction cache and the kernel slows down by 10-30 %).
    int g = 0;
    for (int t = 0; t < ntiles; t++) {
        issue_tile(t + kPllRing - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(kPllRing - 2) : "memory");  // tiles t and t+1 have landed
        __syncwarp();  // a lane only ever reads the row it filled itself
        const float4* cur = reinterpret_cast<const float4*>(&tile[t % kPllRing][lane][0]);
        const float4* nxt = reinterpret_cast<const float4*>(&tile[(t + 1) % kPllRing][lane][0]);
        if (t == 0) {
            vc = cur[0];
            q0 = pll_recip(vc.x); q1 = pll_recip(vc.y); q2 = pll_recip(vc.z); q3 = pll_recip(vc.w);
#if !defined(SDRB_PLL_UNROTATED)
            cr::pll_chunk4r_prime(vc.x, q0, f, kk, hd);
#endif
        }
        const int gend = min(nc, kPllTileChunks * (t + 1));
#if defined(SDRB_PLL_UNROTATED) || defined(SDRB_PLL_INLINE_REDO)
#pragma unroll 2
        for (; g < gend; g++) {
            // next chunk's samples and their reciprocals 1/in (needed by the rotated phase detector): independent of
            // the loop state, they fill the issue slots the dependent chain below leaves empty.  (Past the last chunk
            // of the block this reads stale bytes of the ring, which are never used.)
            const int j = g & (kPllTileChunks - 1);
            const float4 vn = *(j + 1 < kPllTileChunks ? cur + j + 1 : nxt);
            const double p0 = pll_recip(vn.x), p1 = pll_recip(vn.y), p2 = pll_recip(vn.z), p3 = pll_recip(vn.w);
            float4 o;
#if defined(SDRB_PLL_UNROTATED)
            cr::pll_chunk4(vc.x, vc.y, vc.z, vc.w, q0, q1, q2, q3, f, k, kk, tab, o.x, o.y, o.z, o.w, lp.redo);
#else
            cr::pll_chunk4r(vc.x, vc.y, vc.z, vc.w, q0, q1, q2, q3, vn.x, p0, f, hd, k, kk, tab, o.x, o.y, o.z, o.w, lp.redo);
#endif
            o4[g] = o;
            vc = vn;
            q0 = p0; q1 = p1; q2 = p2; q3 = p3;
        }
#else
        // Rotated chunks (pllmath.cuh: pll_chunk4r_spec): this chunk's first phase detector was evaluated by the previous
        // pass, the next chunk's is evaluated before the acceptance flag is looked at.  The careful repeat sits behind a
        // warp-uniform loop exit (a vote), not inside the loop: the common path has no taken forward branch and no
        // reconvergence barrier, which cost ~12 cycles per chunk on a warp that has its scheduler to itself.
        // (Passes of eight samples under one vote were built and measured 6 % slower, unrolled or not: profiles/README.md.)
        while (g < gend) {
            float4 o, vn;
            double p0, p1, p2, p3;
            cr::PllFast saved;
            cr::PllHead hn;
            unsigned bad = 0;
#pragma unroll kPllChunkUnroll
            for (; g < gend; g++) {
                // next chunk's samples and their reciprocals 1/in: independent of the loop state, they fill the issue slots
                // the dependent chain leaves empty.  (Past the last chunk of the block this reads stale bytes of the ring,
                // which are never used.)
                const int j = g & (kPllTileChunks - 1);
                vn = *(j + 1 < kPllTileChunks ? cur + j + 1 : nxt);
                p0 = pll_recip(vn.x); p1 = pll_recip(vn.y); p2 = pll_recip(vn.z); p3 = pll_recip(vn.w);
#if defined(SDRB_PLL_DEFER_STORE)
                if (pending) o4[g - 1] = op;
#endif
                bad = cr::pll_chunk4r_spec(vc.x, vc.y, vc.z, vc.w, q1, q2, q3, vn.x, p0, f, saved, hd, hn, k, kk, o.x, o.y, o.z, o.w);
                if (__any_sync(0xFFFFFFFFu, bad != 0u)) break;
#if defined(SDRB_PLL_DEFER_STORE)
                op = o; pending = true;
#else
                o4[g] = o;
#endif
                hd = hn;
                vc = vn;
                q0 = p0; q1 = p1; q2 = p2; q3 = p3;
            }
            if (g < gend) {  // left through the break: some lanes repeat chunk g on the careful path, the others keep their results
                if (bad) cr::pll_chunk4r_redo(bad, vc.x, vc.y, vc.z, vc.w, q0, q1, q2, q3, vn.x, p0, f, saved, hn, k, kk, tab, o.x, o.y, o.z, o.w, lp.redo);
#if defined(SDRB_PLL_DEFER_STORE)
                if (pending) o4[g - 1] = op;
                pending = false;
#endif
                o4[g] = o;
                hd = hn;
                vc = vn;
                q0 = p0; q1 = p1; q2 = p2; q3 = p3;
                g++;
            }
        }
#endif
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#if defined(SDRB_PLL_DEFER_STORE)
    if (pending) o4[g - 1] = op;
#endif
#if !defined(SDRB_PLL_UNROTATED)
    if (!f.generic_next) cr::pll_fast_resync(f, k);  // the rotated loop leaves sa / cr without their tie test
#endif
    for (int i = n4; i < a.n; i++) out[i] = cr::pll_step_fast(x[i], cr::pll_guard_recip(x[i], pll_recip(x[i])), f, k, tab);
    // tail -> halo of the next slot
    float* nh = lp.trig.nxt + (size_t)s * lp.trig.pitch;
    for (int b = 1; b <= lp.trig.halo && b <= a.n; b++) nh[-b] = out[a.n - b];
    cr::pll_fast_store(f, st);
    lp.st[s] = PllStateDev{st.feedbackI, st.feedbackQ, st.integrator, st.phaseEst, st.trigOffset};
#if defined(SDRB_PLL_TIMESTAMPS)
    if (threadIdx.x == 0) {
        unsigned long long ts1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        printf("PLLTS %llu %llu %d %d %u\n", ts0, ts1, (int)blockIdx.x, (int)blockIdx.y, smid);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// K3b  Mixers (elementwise, parallel): NCO cosines from the stored trigArg and the two products
//   stereo_dc[i] = (float)(2.0 * stereo_band[i] * carrier[i])        /root/reference/src/stereo.cpp:83-85
//   rds_dc[i]    = 2 * rds_band[i-50] * IPLL[i]   (float)              src/rds.cpp:122-127
// carrier[i] / IPLL[i] is the NCO output after input sample i-1 (src/pll.cpp:18,52), i.e.
// cos(ncoScale * trig[i-1] + phaseAdjust); trig[-1] of the very first block is 0 -> 1.0 (the seed the
// callers put in pllOut[N], src/stereo.cpp:45, src/rds.cpp:38).
// ------------------------------------------------------------------------------------------------
struct MixArgs {
    int n, n_streams;
    // stereo
    const float* band;  size_t band_pitch;
    const float* trig19; size_t trig19_pitch;
    RingView stereo_dc;
    float* carrier_out;  // optional [n_streams][n+1]
    // rds (rds_band == nullptr: skip)
    const float* rds_band; size_t rds_band_pitch;
    const float* trig114; size_t trig114_pitch;
    RingView rds_dc;
    float* ipll_out;     // optional [n_streams][n+1]
    float* delay_out;    // optional rds_band_delay [n_streams][n]
    float scale19, adjust19, scale114, adjust114;
    int do_stereo;
};

#ifndef SDRB_MIX_PER
#define SDRB_MIX_PER 4
#endif
#ifndef SDRB_MIX_MINB
#define SDRB_MIX_MINB 4
#endif
constexpr int kMixPer = SDRB_MIX_PER;  // consecutive samples per thread: all loads first, then the cosines

// One sample of the generic (bounds- and option-checked) form: the tail of a row and the stage-dumping parity runs.
__device__ __forceinline__ void mix_one_checked(const MixArgs& a, int s, int i, bool rds, const double* ctab) {
    const bool in_blk = i < a.n;
    if (a.do_stereo) {
        const float th = a.trig19[(size_t)s * a.trig19_pitch + i - 1];
        const float car = cr::cos_lean_f(__fadd_rn(__fmul_rn(th, a.scale19), a.adjust19), ctab);
        if (a.carrier_out) a.carrier_out[(size_t)s * (a.n + 1) + i] = car;
        if (in_blk) {
            const float band = a.band[(size_t)s * a.band_pitch + i];
            ring_store(a.stereo_dc, s, i, __double2float_rn(__dmul_rn(__dmul_rn(2.0, (double)band), (double)car)));
        }
    }
    if (rds) {
        const float th = a.trig114[(size_t)s * a.trig114_pitch + i - 1];
        const float ip = cr::cos_lean_f(__fadd_rn(__fmul_rn(th, a.scale114), a.adjust114), ctab);
        if (a.ipll_out) a.ipll_out[(size_t)s * (a.n + 1) + i] = ip;
        if (in_blk) {
            // the all-pass "delay" FIR (src/rds.cpp:122): 0 + 1*x[i-50] + 0*... == 0.0f + x[i-50]
            const float d = __fadd_rn(0.0f, a.rds_band[(size_t)s * a.rds_band_pitch + i - 50]);
            if (a.delay_out) a.delay_out[(size_t)s * a.n + i] = d;
            ring_store(a.rds_dc, s, i, __fmul_rn(__fmul_rn(2.0f, d), ip));
        }
    }
}

// FAST: no stage dumps are requested and both mixers run (the production configuration): a thread whose kMixPer
// samples lie inside the block and before the halo tail takes a path without a single bounds or option test.
template <bool FAST>
__global__ void __launch_bounds__(256, SDRB_MIX_MINB) k_mix(const MixArgs a) {
    // the cosine's coefficient table where every lane can index its own row with one shared-memory load per coefficient
    // (selecting twelve 32-bit immediates per cosine cost more instructions than the polynomial itself)
    __shared__ double ctab[16];
    if (threadIdx.x < 16) ctab[threadIdx.x] = cr::c_cos_lean_tab[threadIdx.x];
    __syncthreads();
    const int s = blockIdx.y;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * kMixPer;
    if (i0 > a.n) return;
    const bool rds = a.rds_band != nullptr;
    if (FAST && i0 + kMixPer <= a.n - max(a.stereo_dc.halo, a.rds_dc.halo)) {
        // Vector accesses (the rows are 16-byte aligned at sample 0 and i0 is a multiple of four): one float4 per band
        // and output, two per NCO-phase row (the phase of sample i-1: the aligned quad before this one supplies [i0-1]),
        // two float2 for the delayed RDS band (offset -50: 8-byte aligned).  Scalar accesses at a 16-byte lane stride kept
        // the load/store unit busier than the issue slots (ncu: Mem Busy 63 %).
        static_assert(kMixPer == 4, "the fast path of k_mix is written for four samples per thread");
        const float* t19 = a.trig19 + (size_t)s * a.trig19_pitch + i0;
        const float* t114 = a.trig114 + (size_t)s * a.trig114_pitch + i0;
        const float* rb = a.rds_band + (size_t)s * a.rds_band_pitch + i0 - 50;
        const float4 bq = *reinterpret_cast<const float4*>(a.band + (size_t)s * a.band_pitch + i0);
        const float4 p19 = *reinterpret_cast<const float4*>(t19 - 4), c19 = *reinterpret_cast<const float4*>(t19);
        const float4 p114 = *reinterpret_cast<const float4*>(t114 - 4), c114 = *reinterpret_cast<const float4*>(t114);
        const float2 r0 = *reinterpret_cast<const float2*>(rb), r1 = *reinterpret_cast<const float2*>(rb + 2);
        const float th19[4] = {p19.w, c19.x, c19.y, c19.z}, band[4] = {bq.x, bq.y, bq.z, bq.w};
        const float th114[4] = {p114.w, c114.x, c114.y, c114.z}, rbd[4] = {r0.x, r0.y, r1.x, r1.y};
        float so[4], ro[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float car = cr::cos_lean_f(__fadd_rn(__fmul_rn(th19[j], a.scale19), a.adjust19), ctab);
            so[j] = __double2float_rn(__dmul_rn(__dmul_rn(2.0, (double)band[j]), (double)car));
            const float ip = cr::cos_lean_f(__fadd_rn(__fmul_rn(th114[j], a.scale114), a.adjust114), ctab);
            ro[j] = __fmul_rn(__fmul_rn(2.0f, __fadd_rn(0.0f, rbd[j])), ip);
        }
        *reinterpret_cast<float4*>(a.stereo_dc.cur + (size_t)s * a.stereo_dc.pitch + i0) = make_float4(so[0], so[1], so[2], so[3]);
        *reinterpret_cast<float4*>(a.rds_dc.cur + (size_t)s * a.rds_dc.pitch + i0) = make_float4(ro[0], ro[1], ro[2], ro[3]);
        return;
    }
    for (int j = 0; j < kMixPer; j++) {
        const int i = i0 + j;
        if (i > a.n) break;
        mix_one_checked(a, s, i, rds, ctab);
    }
}

// ------------------------------------------------------------------------------------------------
// K4  Audio: 101-tap low-pass + decimate (up == 1 fast path) on mono and L-R, recombine, int16.
// /root/reference/src/mono.cpp:34-42, src/stereo.cpp:88-107, src/filter.cpp:123-147.
// ------------------------------------------------------------------------------------------------
// float -> short exactly as the x86-64 reference build does it: cvttss2si (INT_MIN when out of
// range or NaN), then the low 16 bits.
__device__ __forceinline__ int16_t to_pcm(float v) {
    int i = (v > -2147483904.0f && v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
    return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}

struct AudioArgs {
    const float* mono_x;  // ring slot of the mono path input, already offset by the delay (sample 0)
    size_t mono_pitch;
    const float* dc_x;    // stereo_dc ring slot (nullptr: mono only)
    size_t dc_pitch;
    int n_in;             // IF samples per block
    int n_out;            // audio frames per block
    int up, down;
    const float* taps_pm; // generic path: phase-major taps [up][kTaps] (taps_pm[p*kTaps+j] = h[p + up*j])
    // phase-class path (k_audio_updown_pc): thread-major taps [kTaps][kUpdThreads] and the output residue each thread owns
    const float* taps_tm;
    const int* thread_res;
    int16_t* pcm;         // [n_streams][pcm_pitch]
    size_t pcm_pitch;
    float* mono_out;      // optional [n_streams][n_out]
    float* dc_out;        // optional
};

#if !defined(SDRB_AUD_R)
#define SDRB_AUD_R 4
#define SDRB_AUD_THREADS 64
#endif
constexpr int kAudR = SDRB_AUD_R;
constexpr int kAudThreads = SDRB_AUD_THREADS;
constexpr int kAudTile = kAudR * kAudThreads;

template <int DOWN, bool STEREO>
__global__ void __launch_bounds__(kAudThreads) k_audio_decim(const __grid_constant__ Taps101 taps, const AudioArgs a) {
    constexpr int L = DOWN * kAudR;
    constexpr int NS = DOWN * (kAudTile - 1) + kTaps;
    __shared__ float2 sx[NS + NS / L + 2];
    const int s = blockIdx.y;
    const int m0 = blockIdx.x * kAudTile;
    const int g0 = DOWN * m0 - kState;
    const float* mx = a.mono_x + (size_t)s * a.mono_pitch;
    const float* dx = STEREO ? a.dc_x + (size_t)s * a.dc_pitch : nullptr;
    // (Loading a tile two samples at a time with all of a thread's loads issued before its first store was built and
    // measured: 0.0399 against 0.0402 ms, no change - profiles/README.md - so the plain loop stays.)
    for (int u = threadIdx.x; u < NS; u += kAudThreads) {
        int g = g0 + u;
        float2 v = make_float2(0.0f, 0.0f);
        if (g < a.n_in) {
            // stereo(): mono goes through the all-pass delay FIR first == 0.0f + x (src/stereo.cpp:88);
            // mono(): fed directly (src/mono.cpp:34).  Adding +0 is exact either way except for -0.
            v.x = STEREO ? __fadd_rn(0.0f, mx[g]) : mx[g];
            if (STEREO) v.y = dx[g];
        }
        sx[pad_pos<L>(u)] = v;
    }
    __syncthreads();
    float2 acc[kAudR];
#pragma unroll
    for (int j = 0; j < kAudR; j++) acc[j] = make_float2(0.0f, 0.0f);
    fir_core<DOWN, kAudR, float2>(sx + (L + 1) * threadIdx.x, taps, acc);
    const int m = m0 + kAudR * threadIdx.x;
#pragma unroll
    for (int j = 0; j < kAudR; j++) {
        if (m + j < a.n_out) {
            if (STEREO) {
                int16_t* p = a.pcm + (size_t)s * a.pcm_pitch + 2 * (m + j);
                p[0] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(acc[j].x, acc[j].y)));   // left, even index
                p[1] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(acc[j].x, -acc[j].y)));  // right, odd index
                if (a.mono_out) {
                    a.mono_out[(size_t)s * a.n_out + m + j] = acc[j].x;
                    a.dc_out[(size_t)s * a.n_out + m + j] = acc[j].y;
                }
            } else {
                a.pcm[(size_t)s * a.pcm_pitch + m + j] = to_pcm(__fmul_rn(16384.0f, acc[j].x));
                if (a.mono_out) a.mono_out[(size_t)s * a.n_out + m + j] = acc[j].x;
            }
        }
    }
}

// Generic rational resampler (any up/down; used for up = 147).  One thread per output frame.
// y[n] = sum_j h[phase + up*j] * x[(n*down - phase)/up - j],  phase = (n*down) % up.
template <bool STEREO>
__global__ void __launch_bounds__(128) k_audio_updown(const AudioArgs a) {
    const int s = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n_out) return;
    const long long nd = (long long)n * a.down;
    const int phase = (int)(nd % a.up);
    const int base = (int)(nd / a.up);
    const float* h = a.taps_pm + (size_t)phase * kTaps;
    const float* mx = a.mono_x + (size_t)s * a.mono_pitch + base;
    const float* dx = STEREO ? a.dc_x + (size_t)s * a.dc_pitch + base : nullptr;
    float am = 0.0f, ad = 0.0f;
#pragma unroll 4
    for (int j = 0; j < kTaps; j++) {
        float hj = __ldg(h + j);
        float xm = STEREO ? __fadd_rn(0.0f, mx[-j]) : mx[-j];
        am = mac(am, hj, xm);
        if (STEREO) ad = mac(ad, hj, dx[-j]);
    }
    if (STEREO) {
        int16_t* p = a.pcm + (size_t)s * a.pcm_pitch + 2 * n;
        p[0] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(am, ad)));
        p[1] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(am, -ad)));
        if (a.mono_out) {
            a.mono_out[(size_t)s * a.n_out + n] = am;
            a.dc_out[(size_t)s * a.n_out + n] = ad;
        }
    } else {
        a.pcm[(size_t)s * a.pcm_pitch + n] = to_pcm(__fmul_rn(16384.0f, am));
        if (a.mono_out) a.mono_out[(size_t)s * a.n_out + n] = am;
    }
}

// Rational resampler by phase classes (147/800 and 147/1280 of modes 2 and 3; any up <= kUpdThreads that divides n_out).
// Outputs n and n + up use the same polyphase branch h[phase + up j] and inputs `down` apart, so one thread owns the
// residue class n = r + up q, q = 0 .. n_out/up - 1 (10 outputs here), streams its 101 branch taps once (thread-major
// table: a warp's tap loads are one coalesced row) and reads its inputs from the block staged whole in shared memory.
// Stereo: the staged element is (mono, stereo_dc), so one 64-bit load feeds both lanes of a packed MAC; mono: outputs
// q and q+1 share the packed accumulator.  Which residue a thread owns is a host-made permutation that spreads a
// half-warp's loads over the banks.  (The one-thread-per-output kernel above did 101 scattered tap loads and 101-202
// global input loads per output: 0.23 ms per 1024 mode-2 blocks, 4 % of the MAC issue rate.)
constexpr int kUpdThreads = 160;
constexpr int kUpdMaxQ = 10;

template <bool STEREO>
__global__ void __launch_bounds__(kUpdThreads) k_audio_updown_pc(const AudioArgs a) {
    extern __shared__ __align__(16) float upd_smem[];
    const int s = blockIdx.x, t = threadIdx.x;
    const int nx = a.n_in + kState;
    const float* mx = a.mono_x + (size_t)s * a.mono_pitch - kState;
    if (STEREO) {
        float2* sx = reinterpret_cast<float2*>(upd_smem);
        const float* dx = a.dc_x + (size_t)s * a.dc_pitch - kState;
        for (int u = t; u < nx; u += kUpdThreads) sx[u] = make_float2(__fadd_rn(0.0f, mx[u]), dx[u]);  // 0.0f + x: src/stereo.cpp:88
    } else {
        for (int u = t; u < nx; u += kUpdThreads) upd_smem[u] = mx[u];
    }
    __syncthreads();
    const int r = a.thread_res[t];
    if (r < 0) return;
    const int Q = a.n_out / a.up;
    const long long rd = (long long)r * a.down;
    const int base = (int)(rd / a.up) + kState;  // index of x[(n*down - phase)/up] for q = 0
    if (STEREO) {
        const float2* xb = reinterpret_cast<const float2*>(upd_smem) + base;
        float2 acc[kUpdMaxQ];
#pragma unroll
        for (int q = 0; q < kUpdMaxQ; q++) acc[q] = make_float2(0.0f, 0.0f);
#pragma unroll 4
        for (int j = 0; j < kTaps; j++) {
            const float hj = __ldg(a.taps_tm + j * kUpdThreads + t);
#pragma unroll
            for (int q = 0; q < kUpdMaxQ; q++)
                if (q < Q) acc[q] = mac(acc[q], hj, xb[q * a.down - j]);
        }
#pragma unroll
        for (int q = 0; q < kUpdMaxQ; q++) {
            if (q >= Q) break;
            const int n = r + a.up * q;
            int16_t* p = a.pcm + (size_t)s * a.pcm_pitch + 2 * n;
            p[0] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(acc[q].x, acc[q].y)));
            p[1] = to_pcm(__fmul_rn(16384.0f, __fadd_rn(acc[q].x, -acc[q].y)));
            if (a.mono_out) {
                a.mono_out[(size_t)s * a.n_out + n] = acc[q].x;
                a.dc_out[(size_t)s * a.n_out + n] = acc[q].y;
            }
        }
    } else {
        const float* xb = upd_smem + base;
        float2 acc[kUpdMaxQ / 2];
#pragma unroll
        for (int q = 0; q < kUpdMaxQ / 2; q++) acc[q] = make_float2(0.0f, 0.0f);
#pragma unroll 4
        for (int j = 0; j < kTaps; j++) {
            const float hj = __ldg(a.taps_tm + j * kUpdThreads + t);
#pragma unroll
            for (int q = 0; q < kUpdMaxQ / 2; q++) {
                const float x0 = 2 * q < Q ? xb[2 * q * a.down - j] : 0.0f;
                const float x1 = 2 * q + 1 < Q ? xb[(2 * q + 1) * a.down - j] : 0.0f;
                acc[q] = mac(acc[q], hj, make_float2(x0, x1));
            }
        }
#pragma unroll
        for (int q = 0; q < kUpdMaxQ; q++) {
            if (q >= Q) break;
            const int n = r + a.up * q;
            const float y = (q & 1) ? acc[q / 2].y : acc[q / 2].x;
            a.pcm[(size_t)s * a.pcm_pitch + n] = to_pcm(__fmul_rn(16384.0f, y));
            if (a.mono_out) a.mono_out[(size_t)s * a.n_out + n] = y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5  RDS back end, one CTA per stream: 247/640 resampler + 3 kHz LPF, RRC, clock recovery, slicer,
// Manchester and differential decode, and (every 15 decode blocks) frame sync.
// /root/reference/src/rds.cpp:130-189, src/rds_utilities.cpp:4-88,313-400.
// ------------------------------------------------------------------------------------------------
constexpr int kRdsThreads = 256;
constexpr int kRdsMaxBits = 48;
constexpr int kRdsMaxGroups = 8;
constexpr int kBitBufWords = 26;  // 832 bits of storage; at most kBitBufCap are ever occupied
constexpr int kBitBufCap = 704;    // 15 blocks x <= 38 bits + carried leftover (< 64) fits with room to spare

struct RdsRecord {  // mirrors sdrb_rds_record
    int32_t cdr_offset, n_symbols, n_bits, n_groups;
    uint8_t bits[kRdsMaxBits];
    uint64_t groups[kRdsMaxGroups];
};

struct RdsStreamState {
    int32_t block_count;   // src/rds.cpp:28
    int32_t decoder_cont;  // :32
    int32_t half_symbol, start, last_bit;  // :29-31
    int32_t nbits;         // bits waiting in bitbuf (carried leftover + this period's blocks)
    int32_t window[4];     // last four matched offsets (0 A, 1 B, 2 C, 3 C', 4 D), src/rds.cpp:70
    int32_t nwindow;
    int32_t first_time;
    uint64_t reg;          // :67
    uint32_t bitbuf[kBitBufWords];  // bit i at word i/32, bit i%32
};

struct RdsArgs {
    const float* dc;       // rds_dc ring slot (halo >= kState)
    size_t dc_pitch;
    int n_in;              // IF samples per block
    int n_out;             // n_in*247/640
    int sps;
    int rds_on;
    const float2* res_taps;  // [kResIter][128]: the two branch taps of resampler lane `tl` at loop step j (see ResLane)
    const int4* res_lane;    // [128] ResLane as int4
    Taps101 rrc;
    float* filt_state_in;    // [n_streams][kState] last rds_filt samples of the previous block
    float* filt_state_out;
    RdsStreamState* st;      // [n_streams]
    RdsRecord* rec;          // [n_streams]
    float* filt_out;         // optional [n_streams][n_out]
    float* clean_out;        // optional
    unsigned int* overflow;  // [3] events in which a fixed capacity cut data off: bits per block, bit buffer, groups per block
};

// check_block's syndrome test (/root/reference/src/rds_utilities.cpp:357-366) on a 26-bit window held
// LSB-first (bit t = t-th received bit).  Masks are the parity_matrix rows of :122-133, the syndromes
// those of :135, both re-packed LSB-first.  Returns 0 A, 1 B, 2 C, 3 C', 4 D, -1 none.
__device__ __forceinline__ int rds_block_type(uint32_t w) {
    constexpr uint32_t rows[10] = {0x39BE401u, 0x337C802u, 0x1F47404u, 0x0730C08u, 0x0E61810u,
                                   0x257D420u, 0x3344C40u, 0x1F37C80u, 0x3E6F900u, 0x3CDF200u};
    uint32_t syn = 0;
#pragma unroll
    for (int c = 0; c < 10; c++) syn |= (uint32_t)(__popc(w & rows[c]) & 1) << c;
    int t = -1;
    if (syn == 0x06Fu) t = 0;
    if (syn == 0x0AFu) t = 1;
    if (syn == 0x0E9u) t = 2;
    if (syn == 0x0CFu) t = 3;
    if (syn == 0x069u) t = 4;
    return t;
}

__device__ __forceinline__ uint32_t bitbuf_window26(const uint32_t* buf, int idx) {
    int w = idx >> 5, sh = idx & 31;
    uint64_t two = (uint64_t)buf[w] | ((uint64_t)buf[w + 1] << 32);
    return (uint32_t)(two >> sh) & 0x3FFFFFFu;
}

constexpr int kRrcR = 12;
constexpr int kRrcTile = kRrcR * kRdsThreads;  // 3072 outputs per pass
static_assert(kResTaps == kTaps, "res_lanes.h");  // resampler lanes, loop steps and lags: res_lanes.h
// floats of the staged input: the block with its carried state, or (longer) what the loads of a lane may touch: outputs
// that do not exist (q = 11 for most residues) are computed on whatever lies there and never stored
__host__ __device__ constexpr int rds_sdc_len(int n_in) {
    const int need = kState + kRdsDown * kResQ + kResLagMax + 1, have = n_in + kState;
    return ((need > have ? need : have) + 3) / 4 * 4;
}


__global__ void __launch_bounds__(kRdsThreads) k_rds_backend(const __grid_constant__ RdsArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int s = blockIdx.x;
    const int t = threadIdx.x;
    const int n_in = a.n_in, n_out = a.n_out;
    const int rrc_tiles = (n_out + kRrcTile - 1) / kRrcTile;
    const int nfilt_pad = rrc_tiles * kRrcTile + kState;
    float* sdc = smem;                                // [n_in + kState]; later reused as sclean [n_out]
    float* sfilt = smem + rds_sdc_len(n_in);            // padded layout, pad_pos<kRrcR>
    __shared__ int ssum[64];
    __shared__ int ssym[160];
    __shared__ int8_t stype[kBitBufWords * 32];
    __shared__ int soff;

    // Carried decoder state: every thread reads it here, before the first barrier; thread 0 rewrites it only
    // after the last one (reading it later would race with that write across warps).
    RdsStreamState* st = a.st + s;
    RdsRecord* rec = a.rec + s;
    const int block_count = st->block_count;
    const int start_in = st->start, half_in = st->half_symbol, last_in = st->last_bit;
    const int nbits_in = st->nbits, decoder_cont_in = st->decoder_cont;

    // ---- stage rds_dc (with the carried 100-sample state in front): one bulk copy (TMA) issued by thread 0 when the row is
    // 16-byte aligned (it is for the chain's own rings), the few floats behind the last multiple of 16 bytes by plain loads
    static_assert(kRdsThreads == 2 * kResLanes, "two threads per resampler lane");
    __shared__ __align__(8) unsigned long long bar;
    const float* dc = a.dc + (size_t)s * a.dc_pitch - kState;
    const int n_stage = n_in + kState;
#if defined(SDRB_RDS_NO_TMA)  // measurement variant: plain loads only
    const int n_bulk = 0;
#else
    const int n_bulk = ((reinterpret_cast<uintptr_t>(dc) & 15u) == 0) ? (n_stage & ~3) : 0;
#endif
    if (n_bulk) {
        if (t == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (t == 0) {
            mbar_expect_tx(&bar, (uint32_t)n_bulk * 4u);
            tma_load_1d(sdc, dc, (uint32_t)n_bulk * 4u, &bar);
        }
    }
    for (int u = n_bulk + t; u < n_stage; u += kRdsThreads) sdc[u] = dc[u];
    for (int u = t; u < kState; u += kRdsThreads)
        sfilt[pad_pos<kRrcR>(u)] = a.filt_state_in[(size_t)s * kState + u];
    for (int u = n_out + kState + t; u < nfilt_pad; u += kRdsThreads) sfilt[pad_pos<kRrcR>(u)] = 0.0f;
    if (t < 64) ssum[t] = 0;
    if (n_bulk) mbar_wait(&bar, 0);
    __syncthreads();

    // ---- 247/640 resampler (/root/reference/src/filter.cpp:123-147, src/rds.cpp:130).
    // Outputs n and n + 247 share the polyphase branch (taps h[phase + 247 j]) and adjacent residues share the input
    // (see ResLane above): a thread owns two residues x six outputs n = tp + 247 q, loads per loop step one sample per q
    // and one float2 of taps, and does twelve MACs with them, outputs 2p and 2p+1 of a residue in one packed accumulator
    // (FFMA2 + FADD2).  Round 2 started from one residue per thread (one shared-memory load per MAC: the kernel was bound
    // by the load/store unit at twice the time of its floating-point work).
    {
        const int tl = t & (kResLanes - 1), qh = t / kResLanes;
        const int4 ln = a.res_lane[tl];  // x tp_hi (-1: idle lane), y tp_lo (-1: none), z base, w lag_hi | lag_lo << 8
        if (ln.x >= 0 && qh < 2) {
            constexpr int HQ = kResQ / 2;       // outputs per residue and thread
            const int lag_hi = ln.w & 0xFF, lag_lo = (ln.w >> 8) & 0xFF;
            const bool has_lo = ln.y >= 0;
            const float* xb = sdc + kState + ln.z + kRdsDown * HQ * qh;  // x[base + 640 q - j] for the thread's first q at j = 0
            const float2* tab = a.res_taps + tl;
            float2 aH[HQ / 2], aL[HQ / 2];
#pragma unroll
            for (int p = 0; p < HQ / 2; p++) aH[p] = aL[p] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < kResLagMax; j++) {  // first steps: a residue joins at its lag
                const float2 h = __ldg(tab + j * kResLanes);
                const bool vH = j >= lag_hi, vL = has_lo && j >= lag_lo;
#pragma unroll
                for (int p = 0; p < HQ / 2; p++) {
                    const float2 x = make_float2(xb[2 * kRdsDown * p - j], xb[2 * kRdsDown * p + kRdsDown - j]);
                    if (vH) aH[p] = mac(aH[p], h.x, x);
                    if (vL) aL[p] = mac(aL[p], h.y, x);
                }
            }
#pragma unroll 7  // 98 uniform steps; seven tap loads in flight ahead of their MACs
            for (int j = kResLagMax; j < kTaps; j++) {
                const float2 h = __ldg(tab + j * kResLanes);
#pragma unroll
                for (int p = 0; p < HQ / 2; p++) {
                    const float2 x = make_float2(xb[2 * kRdsDown * p - j], xb[2 * kRdsDown * p + kRdsDown - j]);
                    aH[p] = mac(aH[p], h.x, x);
                    aL[p] = mac(aL[p], h.y, x);  // (a lane without a partner has zero taps here and never stores the result)
                }
            }
#pragma unroll
            for (int j = kTaps; j < kResIter; j++) {  // last steps: a residue leaves after its 101st tap
                const float2 h = __ldg(tab + j * kResLanes);
                const bool vH = j < kTaps + lag_hi, vL = has_lo && j < kTaps + lag_lo;
#pragma unroll
                for (int p = 0; p < HQ / 2; p++) {
                    const float2 x = make_float2(xb[2 * kRdsDown * p - j], xb[2 * kRdsDown * p + kRdsDown - j]);
                    if (vH) aH[p] = mac(aH[p], h.x, x);
                    if (vL) aL[p] = mac(aL[p], h.y, x);
                }
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int tp = r ? ln.y : ln.x;
                if (tp < 0) continue;
#pragma unroll
                for (int qq = 0; qq < HQ; qq++) {
                    const int n = (HQ * qh + qq) * kRdsUp + tp;
                    const float2 v = r ? aL[qq / 2] : aH[qq / 2];
                    const float y = (qq & 1) ? v.y : v.x;
                    if (n < n_out) {
                        sfilt[pad_pos<kRrcR>(n + kState)] = y;
                        if (a.filt_out) a.filt_out[(size_t)s * n_out + n] = y;
                        if (n >= n_out - kState) a.filt_state_out[(size_t)s * kState + (n - (n_out - kState))] = y;
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- RRC matched filter (src/rds.cpp:133), 12 consecutive outputs per thread
    float* sclean = smem;
    for (int tile = 0; tile < rrc_tiles; tile++) {
        float acc[kRrcR];
#pragma unroll
        for (int j = 0; j < kRrcR; j++) acc[j] = 0.0f;
        fir_core<1, kRrcR, float>(sfilt + pad_pos<kRrcR>(tile * kRrcTile) + (kRrcR + 1) * t, a.rrc, acc);
#pragma unroll
        for (int j = 0; j < kRrcR; j++) {
            int n = tile * kRrcTile + kRrcR * t + j;
            if (n < n_out) {
                sclean[n] = acc[j];
                if (a.clean_out) a.clean_out[(size_t)s * n_out + n] = acc[j];
            }
        }
    }
    __syncthreads();

    const bool decode = block_count > 5 && a.rds_on;  // src/rds.cpp:135
    if (!decode) {
        if (t == 0) {
            rec->cdr_offset = -1; rec->n_symbols = 0; rec->n_bits = 0; rec->n_groups = 0;
            st->block_count = block_count + 1;
        }
        return;
    }

    // ---- clock recovery: argmax over the sps sampling phases of sum |(int)x| (src/rds_utilities.cpp:4-21)
    const int sps = a.sps;
    {
        const int parts = kRdsThreads / sps;
        const int i = t % sps, part = t / sps;
        const int nk = n_out / sps;
        if (part < parts) {
            int sum = 0;
            for (int k = part; k < nk; k += parts) sum += abs(__float2int_rz(sclean[k * sps + i]));
            atomicAdd(&ssum[i], sum);
        }
    }
    __syncthreads();
    if (t == 0) {
        int maxi = 0, maxv = 0;
        for (int i = 0; i < sps; i++)
            if (ssum[i] > maxv) { maxv = ssum[i]; maxi = i; }
        soff = maxi;
    }
    __syncthreads();
    const int off = soff;
    const int nsym = (n_out - off + sps - 1) / sps;  // src/rds.cpp:157-161
    if (t < nsym && t < 160) ssym[t] = sclean[off + t * sps] > 0.0f;
    __syncthreads();
    if (t >= 32) return;

    // ---- warp 0: Manchester pairing (src/rds_utilities.cpp:34-68) as two ballots, differential decode
    // (:70-88) as one shifted XOR on the 64-bit word.
    int nb = start_in + ((nsym - 1 - start_in) > 0 ? (nsym - 1 - start_in + 1) / 2 : 0);
    if (nb > kRdsMaxBits) {  // cannot happen with sps = 39 and 2836 samples per block (<= 37 bits); counted, never silent
        nb = kRdsMaxBits;
        if (t == 0) atomicAdd(a.overflow + 0, 1u);
    }
    uint32_t wlo, whi;
    {
        int q = t, bit = 0;
        if (q < nb) bit = (start_in && q == 0) ? half_in : ssym[start_in + 2 * (q - start_in)];
        wlo = __ballot_sync(0xFFFFFFFFu, bit != 0);
        q = t + 32; bit = 0;
        if (q < nb) bit = ssym[start_in + 2 * (q - start_in)];
        whi = __ballot_sync(0xFFFFFFFFu, bit != 0);
    }
    const uint64_t W = (uint64_t)wlo | ((uint64_t)whi << 32);
    const uint64_t Dm = (W ^ ((W << 1) | (uint64_t)(last_in & 1))) & (nb >= 64 ? ~0ull : ((1ull << nb) - 1));
    for (int q = t; q < kRdsMaxBits; q += 32) rec->bits[q] = (q < nb) ? (uint8_t)((Dm >> q) & 1) : 0;

    int nbits = nbits_in;
    int decoder_cont = decoder_cont_in + 1;
    if (t == 0) {
        // carried symbol state (:61-66), last bit (:87)
        if (((unsigned)nsym - (unsigned)start_in) & 1u) { st->half_symbol = ssym[nsym - 1]; st->start = 1; }
        else st->start = 0;
        if (nb > 0) st->last_bit = (int)((W >> (nb - 1)) & 1);
        // append to the stream buffer (src/rds.cpp:182)
        if (nbits + nb <= kBitBufCap) {
            int w = nbits >> 5, sh = nbits & 31;
            uint64_t lo = Dm << sh;
            st->bitbuf[w] = (st->bitbuf[w] & ((1u << sh) - 1u)) | (uint32_t)lo;
            st->bitbuf[w + 1] = (uint32_t)(lo >> 32);
            st->bitbuf[w + 2] = sh ? (uint32_t)(Dm >> (64 - sh)) : 0u;
            nbits += nb;
        } else {
            atomicAdd(a.overflow + 1, 1u);  // the block's bits are dropped (15 blocks x 37 bits + carry < kBitBufCap: unreachable)
        }
        rec->cdr_offset = off; rec->n_symbols = nsym; rec->n_bits = nb;
    }
    nbits = __shfl_sync(0xFFFFFFFFu, nbits, 0);
    __syncwarp();

    int ngroups = 0;
    if (decoder_cont == 15) {  // src/rds.cpp:184-189 -> start_frame_sync, src/rds_utilities.cpp:384-400
        decoder_cont = 0;
        const int total = nbits;
        const int end_range = total >= 26 ? total - 26 : 0;  // idx < size-26: the last full window waits
        for (int idx = t; idx < end_range; idx += 32) stype[idx] = (int8_t)rds_block_type(bitbuf_window26(st->bitbuf, idx));
        __syncwarp();
        if (t == 0) {
            uint64_t reg = st->reg;
            int win[4] = {st->window[0], st->window[1], st->window[2], st->window[3]};
            int nwin = st->nwindow;
            int idx = 0;
            while (idx < end_range) {
                int ty = stype[idx];
                if (ty >= 0) {
                    if (ty != 3) {  // "Cp" matches but copies nothing (:370)
                        int bt = (ty == 4) ? 3 : ty;
                        uint64_t word = (uint64_t)((__brev(bitbuf_window26(st->bitbuf, idx)) >> 16) & 0xFFFFu);
                        int shl = 48 - 16 * bt;
                        reg = (reg & ~((uint64_t)0xFFFF << shl)) | (word << shl);
                    }
                    if (nwin == 4) { win[0] = win[1]; win[1] = win[2]; win[2] = win[3]; nwin = 3; }
                    win[nwin++] = ty;
                    if (nwin == 4 && win[0] == 0 && win[1] == 1 && win[2] == 2 && win[3] == 4) {
                        if (ngroups < kRdsMaxGroups) rec->groups[ngroups] = reg;
                        ngroups++;
                        st->first_time = 0;
                    }
                    idx += 26;
                } else {
                    idx += 1;
                }
            }
            st->reg = reg;
            st->window[0] = win[0]; st->window[1] = win[1]; st->window[2] = win[2]; st->window[3] = win[3];
            st->nwindow = nwin;
            // keep the unread tail as the carry (:398-399): shift the buffer down by idx bits
            int keep = total - idx;
            if (keep < 0) keep = 0;
            uint32_t tmp[kBitBufWords];
            int w0 = idx >> 5, sh = idx & 31;
            for (int w = 0; w < kBitBufWords; w++) {
                uint32_t lo = (w0 + w < kBitBufWords) ? st->bitbuf[w0 + w] : 0u;
                uint32_t hi = (w0 + w + 1 < kBitBufWords) ? st->bitbuf[w0 + w + 1] : 0u;
                tmp[w] = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
            }
            for (int w = 0; w < kBitBufWords; w++) {
                int lim = keep - 32 * w;
                uint32_t m = lim >= 32 ? 0xFFFFFFFFu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
                st->bitbuf[w] = tmp[w] & m;
            }
            nbits = keep;
        }
    }
    if (t == 0) {
        rec->n_groups = ngroups < kRdsMaxGroups ? ngroups : kRdsMaxGroups;
        if (ngroups > kRdsMaxGroups) atomicAdd(a.overflow + 2, 1u);  // 15 blocks hold at most 5.3 groups of 104 bits
        st->nbits = nbits;
        st->decoder_cont = decoder_cont;
        st->block_count = block_count + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone batched primitives behind the per-function ABI (sdrb_fir_decim, sdrb_fir_updown,
// sdrb_fm_demod, sdrb_pll, sdrb_cdr).  Straightforward one-thread-per-output kernels: they serve the
// std::vector shim (batch = 1 drop-in use) and cross-check the fused kernels; the chain does not use them.
// ------------------------------------------------------------------------------------------------
__global__ void k_fir_decim_generic(const float* x, size_t x_pitch, int nx, const float* h, int nh, const float* state,
                                    int nstate, float* y, size_t y_pitch, int decim) {
    const int s = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nx / decim) return;
    const float* xs = x + (size_t)s * x_pitch;
    const float* ss = state + (size_t)s * nstate;
    const int n = m * decim;
    float acc = 0.0f;
    for (int k = 0; k < nh; k++) {
        int j = n - k;
        float xv = (j < 0) ? ((j + nstate >= 0) ? ss[j + nstate] : 0.0f) : xs[j];
        acc = mac(acc, __ldg(h + k), xv);
    }
    y[(size_t)s * y_pitch + m] = acc;
}

__global__ void k_fir_updown_generic(const float* x, size_t x_pitch, int nx, const float* h, int nh, const float* state,
                                     int nstate, float* y, size_t y_pitch, int up, int down) {
    const int s = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (int)(((long long)nx * up) / down)) return;
    const float* xs = x + (size_t)s * x_pitch;
    const float* ss = state + (size_t)s * nstate;
    const long long nd = (long long)n * down;
    const int phase = (int)(nd % up);
    float acc = 0.0f;
    for (int k = phase; k < nh; k += up) {
        int xi = (int)((nd - k) / up);
        float xv = (xi < 0) ? ((xi + nstate >= 0) ? ss[xi + nstate] : 0.0f) : xs[xi];
        acc = mac(acc, __ldg(h + k), xv);
    }
    y[(size_t)s * y_pitch + n] = acc;
}

// state <- last nstate inputs (src/filter.cpp:119,145); runs after the FIR kernel on the same stream
__global__ void k_state_update(const float* x, size_t x_pitch, int nx, float* state, int nstate) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstate) return;
    float* ss = state + (size_t)s * nstate;
    const float* xs = x + (size_t)s * x_pitch;
    // nx >= nstate at every reference call site; the host entry points reject shorter inputs (the reference reads
    // out of bounds there)
    ss[i] = xs[nx - nstate + i];
}

__global__ void k_fm_demod_generic(const float* I, const float* Q, size_t pitch, int n, const float* prev, float* out,
                                   size_t out_pitch) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* Is = I + (size_t)s * pitch;
    const float* Qs = Q + (size_t)s * pitch;
    float pI = i ? Is[i - 1] : prev[2 * s], pQ = i ? Qs[i - 1] : prev[2 * s + 1];
    out[(size_t)s * out_pitch + i] = fm_discriminate(Is[i], Qs[i], pI, pQ);
}
__global__ void k_fm_prev_update(const float* I, const float* Q, size_t pitch, int n, float* prev, int n_streams) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams || n <= 0) return;
    prev[2 * s] = I[(size_t)s * pitch + n - 1];
    prev[2 * s + 1] = Q[(size_t)s * pitch + n - 1];
}

struct PllStateAbi {  // mirrors sdrb_pll_state
    float feedbackI, feedbackQ, integrator, phaseEst;
    double trigOffset;
    float lastCarrier;
    float last_out;
};

__global__ void __launch_bounds__(kPllThreads) k_pll_generic(const float* in, size_t in_pitch, int n, cr::PllCoef k,
                                                              PllStateAbi* stp, float* out, size_t out_pitch, int n_streams) {
    __shared__ cr::AtanTab tab;
    {
        const cr::AtanTab init = SDRB_ATAN_TAB_INIT;
        if (threadIdx.x < 17) {
            tab.hi[threadIdx.x] = init.hi[threadIdx.x];
            tab.lo[threadIdx.x] = init.lo[threadIdx.x];
        }
    }
    __syncthreads();
    const int s = blockIdx.x * kPllThreads + threadIdx.x;
    if (s >= n_streams) return;
    PllStateAbi sa = stp[s];
    cr::PllState st{sa.feedbackI, sa.feedbackQ, sa.integrator, sa.phaseEst, sa.trigOffset};
    const float* x = in + (size_t)s * in_pitch;
    float* o = out + (size_t)s * out_pitch;
    o[0] = sa.last_out;  // src/pll.cpp:18
    float last = sa.last_out;
    for (int i = 0; i < n; i++) {
        float th = cr::pll_step_trig(x[i], st, k, tab);
        last = cr::cos_f(__fadd_rn(__fmul_rn(th, k.ncoScale), k.phaseAdjust));
        o[i + 1] = last;
    }
    stp[s] = PllStateAbi{st.feedbackI, st.feedbackQ, st.integrator, st.phaseEst, st.trigOffset, last, last};
}

__global__ void __launch_bounds__(64) k_cdr_generic(const float* x, size_t x_pitch, int n, int sps, int* offset) {
    __shared__ int ssum[64];
    const int s = blockIdx.x;
    const float* xs = x + (size_t)s * x_pitch;
    int best = 0, bestv = 0;
    for (int i0 = 0; i0 < sps; i0 += 64) {
        int i = i0 + threadIdx.x;
        int sum = 0;
        if (i < sps)
            for (int k = 0; k < n / sps; k++) sum += abs(__float2int_rz(xs[k * sps + i]));
        ssum[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int j = 0; j < 64 && i0 + j < sps; j++)
                if (ssum[j] > bestv) { bestv = ssum[j]; best = i0 + j; }
        __syncthreads();
    }
    if (threadIdx.x == 0) offset[s] = best;
}

// ------------------------------------------------------------------------------------------------
// Stand-alone batched bit-level primitives (sdrb_manchester_decode, sdrb_differential_decode, sdrb_frame_sync):
// one warp per stream.  The fused chain does the same work inside k_rds_backend.
// ------------------------------------------------------------------------------------------------
struct ManchesterState {  // mirrors sdrb_manchester_state
    int32_t half_symbol, start;
};

// manchester_decode, /root/reference/src/rds_utilities.cpp:34-68.  bits[q] = carried half symbol (if the carried
// `start` is set) followed by symbols[i] for i = start, start+2, ... < n-1, where `start` is first re-estimated from the
// pair agreement scores when block_count == 0 (:42-51); then the carry for the next call (:61-66).
__global__ void __launch_bounds__(32) k_manchester_generic(const int32_t* symbols, size_t sym_pitch, const int32_t* nsym, int block_count,
                                                           ManchesterState* state, int32_t* bits, size_t bits_pitch, int32_t* nbits) {
    const int s = blockIdx.x, lane = threadIdx.x;
    const int32_t* sym = symbols + (size_t)s * sym_pitch;
    int32_t* out = bits + (size_t)s * bits_pitch;
    const int n = nsym[s];
    const ManchesterState st = state[s];
    const int lead = st.start ? 1 : 0;  // the carried half symbol is emitted first (:38-40), before any re-estimate
    int start = st.start;
    if (block_count == 0) {
        int score = 0;
        for (int i = 2 * lane; i < n - 1; i += 64) score += sym[i] ^ sym[i + 1];
        for (int j = 2 * lane + 1; j < n - 1; j += 64) score -= sym[j] ^ sym[j + 1];
#pragma unroll
        for (int o = 16; o; o >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, o);
        start = score < 0;
    }
    const int npairs = (n - 1 > start) ? (n - 1 - start + 1) / 2 : 0;
    const int nb = lead + npairs;
    for (int q = lane; q < nb; q += 32) out[q] = (lead && q == 0) ? st.half_symbol : sym[start + 2 * (q - lead)];
    if (lane == 0) {
        nbits[s] = nb;
        ManchesterState o = st;
        if (n > 0) {  // (n == 0: the reference reads symbols[-1]; the carried state is left as it is)
            if ((n - start) & 1) { o.half_symbol = sym[n - 1]; o.start = 1; }
            else o.start = 0;
        }
        state[s] = o;
    }
}

// differential_decode, /root/reference/src/rds_utilities.cpp:70-88: decoded[0] = bits[0] (^ last_bit unless
// block_num == 0), decoded[i] = bits[i] ^ bits[i-1] (the neighbour comes by warp shuffle, the chunk edge from the
// previous chunk's last lane), last_bit = bits[n-1].  n == 0 (undefined in the reference) leaves last_bit alone.
__global__ void __launch_bounds__(32) k_differential_generic(const int32_t* bits, size_t bits_pitch, const int32_t* nbits, int block_num,
                                                             int32_t* last_bit, int32_t* decoded, size_t dec_pitch) {
    const int s = blockIdx.x, lane = threadIdx.x;
    const int32_t* b = bits + (size_t)s * bits_pitch;
    int32_t* d = decoded + (size_t)s * dec_pitch;
    const int n = nbits[s];
    int edge = block_num == 0 ? 0 : last_bit[s];  // the bit before this chunk's lane 0
    for (int q0 = 0; q0 < n; q0 += 32) {
        const int q = q0 + lane;
        const int v = q < n ? b[q] : 0;
        int prev = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if (lane == 0) prev = edge;
        if (q < n) d[q] = v ^ prev;
        edge = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
    if (lane == 0 && n > 0) last_bit[s] = b[n - 1];
}

struct FrameSyncState {  // mirrors sdrb_framesync_state
    uint64_t reg;
    int32_t window[4];
    int32_t nwindow;
    int32_t ncarry;
    uint8_t carry[64];
};
constexpr int kFrameSyncMaxBits = 8192;  // carried tail + new bits per call

// start_frame_sync + check_block, /root/reference/src/rds_utilities.cpp:352-400: the new bits are appended to the
// carried tail, every 26-bit window's syndrome is classified in parallel, then one lane walks the stream with the
// reference's stepping rule (26 after a match, 1 otherwise, while idx < size-26), keeps the group register and the
// window of the last four matched offsets, and reports `reg` at every A,B,C,D completion.
__global__ void __launch_bounds__(32) k_frame_sync_generic(const int32_t* bits, size_t bits_pitch, const int32_t* nbits, FrameSyncState* state,
                                                           unsigned long long* groups, size_t groups_pitch, int32_t* ngroups, int max_groups) {
    __shared__ uint32_t sbuf[kFrameSyncMaxBits / 32 + 2];
    __shared__ int8_t stype[kFrameSyncMaxBits];
    const int s = blockIdx.x, lane = threadIdx.x;
    const int32_t* b = bits + (size_t)s * bits_pitch;
    FrameSyncState* st = state + s;
    const int ncarry = st->ncarry;
    const int total = min(ncarry + max(nbits[s], 0), kFrameSyncMaxBits);
    for (int w = lane; w < kFrameSyncMaxBits / 32 + 2; w += 32) sbuf[w] = 0u;
    __syncwarp();
    for (int q0 = 0; q0 < total; q0 += 32) {  // pack 32 bits per ballot
        const int q = q0 + lane;
        int v = 0;
        if (q < total) v = q < ncarry ? st->carry[q] : b[q - ncarry];
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v != 0);
        if (lane == 0) sbuf[q0 >> 5] = word;
    }
    __syncwarp();
    const int end_range = total >= 26 ? total - 26 : 0;  // idx < size-26: the last full window waits for more bits
    for (int idx = lane; idx < end_range; idx += 32) stype[idx] = (int8_t)rds_block_type(bitbuf_window26(sbuf, idx));
    __syncwarp();
    if (lane == 0) {
        uint64_t reg = st->reg;
        int win[4] = {st->window[0], st->window[1], st->window[2], st->window[3]};
        int nwin = st->nwindow;
        int idx = 0, ng = 0;
        while (idx < end_range) {
            const int ty = stype[idx];
            if (ty >= 0) {
                if (ty != 3) {  // "Cp" matches but copies nothing (:370)
                    const int bt = (ty == 4) ? 3 : ty;
                    const uint64_t word = (uint64_t)((__brev(bitbuf_window26(sbuf, idx)) >> 16) & 0xFFFFu);
                    const int shl = 48 - 16 * bt;
                    reg = (reg & ~((uint64_t)0xFFFF << shl)) | (word << shl);
                }
                if (nwin == 4) { win[0] = win[1]; win[1] = win[2]; win[2] = win[3]; nwin = 3; }
                win[nwin++] = ty;
                if (nwin == 4 && win[0] == 0 && win[1] == 1 && win[2] == 2 && win[3] == 4) {
                    if (ng < max_groups) groups[(size_t)s * groups_pitch + ng] = reg;
                    ng++;
                }
                idx += 26;
            } else {
                idx += 1;
            }
        }
        ngroups[s] = ng;
        st->reg = reg;
        st->window[0] = win[0]; st->window[1] = win[1]; st->window[2] = win[2]; st->window[3] = win[3];
        st->nwindow = nwin;
        int keep = total - idx;  // the unread tail is the carry (:398-399)
        if (keep < 0) keep = 0;
        if (keep > 64) keep = 64;  // cannot happen (keep <= 26); keeps the store in bounds whatever the input
        for (int i = 0; i < keep; i++) st->carry[i] = (uint8_t)((sbuf[(idx + i) >> 5] >> ((idx + i) & 31)) & 1u);
        st->ncarry = keep;
    }
}

// ------------------------------------------------------------------------------------------------
// error_detection: the sync-state-machine RDS decoder the reference declares (include/rds_utilities.h:14) and defines
// (/root/reference/src/rds_utilities.cpp:202-311, calc_syndrome :90-109) but never calls (src/rds.cpp:177-179 is commented
// out).  One warp per stream: the two linear maps of every 26-bit window of the call - the syndrome calc_syndrome(reg, 26)
// = w(z) z^10 mod g(z) that the search for sync compares with the five offset syndromes, and the remainder w(z) mod g(z)
// that the block check of the locked state needs (checkword ^ offset == calc_syndrome(dataword, 16)  <=>  w mod g == offset)
// - are taken in parallel as popcount parities; lane 0 then walks the bits through the reference's state machine, one
// table look-up per bit.  Besides the reference's own observable events (sync found / kept / lost, its one parse() call)
// the walk assembles groups the way the reference meant to (its `registr` is a local that restarts at zero for every
// block): event 5 carries A|B|C|D of every group whose four blocks passed the check in order.
// ------------------------------------------------------------------------------------------------
struct RdsSyncState {  // mirrors sdrb_rds_sync_state
    uint64_t reg;
    int32_t sync, prevsync, lastseen_offset, rds_bit_cont, lastseen_offset_cont, block_distance, block_number, block_bit_cont,
        blocks_cont, wrong_blocks_cont, group_assembly_started, group_good_blocks_cont;
    uint64_t ext_reg;   // extension: group register that persists across the blocks of a group
    int32_t ext_good;   // extension: blocks 0..ext_good-1 of the current group passed the check
    int32_t reserved;
};
struct RdsSyncEvent {  // mirrors sdrb_rds_sync_event
    int32_t type, bit, a, b;
    uint64_t value;
};

__host__ __device__ constexpr uint32_t rds_zpow_mod_g(int e) {  // z^e mod g(z), g = 0x5B9
    uint32_t v = 1;
    for (int i = 0; i < e; i++) {
        v <<= 1;
        if (v & 0x400u) v ^= 0x5B9u;
    }
    return v;
}
// parity mask of output bit c of the map  w -> w(z) z^shift mod g(z)  over the 26 window bits (bit t of w = coefficient of z^t)
__host__ __device__ constexpr uint32_t rds_row_mask(int c, int shift) {
    uint32_t m = 0;
    for (int t = 0; t < 26; t++) m |= ((rds_zpow_mod_g(t + shift) >> c) & 1u) << t;
    return m;
}
template <int SHIFT>
__device__ __forceinline__ uint32_t rds_window_map(uint32_t w) {
    constexpr uint32_t rows[10] = {rds_row_mask(0, SHIFT), rds_row_mask(1, SHIFT), rds_row_mask(2, SHIFT), rds_row_mask(3, SHIFT),
                                   rds_row_mask(4, SHIFT), rds_row_mask(5, SHIFT), rds_row_mask(6, SHIFT), rds_row_mask(7, SHIFT),
                                   rds_row_mask(8, SHIFT), rds_row_mask(9, SHIFT)};
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 10; c++) v |= (uint32_t)(__popc(w & rows[c]) & 1) << c;
    return v;
}

__global__ void __launch_bounds__(32) k_rds_sync_generic(const int32_t* bits, size_t bits_pitch, const int32_t* nbits, RdsSyncState* state,
                                                         RdsSyncEvent* events, size_t events_pitch, int32_t* nevents, int max_events,
                                                         unsigned short* syndromes) {
    __shared__ uint32_t sbuf[kFrameSyncMaxBits / 32 + 3];
    __shared__ unsigned short sS[kFrameSyncMaxBits], sR[kFrameSyncMaxBits];
    const int s = blockIdx.x, lane = threadIdx.x;
    const int32_t* b = bits + (size_t)s * bits_pitch;
    RdsSyncState* stp = state + s;
    const int n = min(max(nbits[s], 0), kFrameSyncMaxBits - 32);
    const uint64_t reg0 = stp->reg;
    // bit p of the buffer: p < 25 the 25 bits before this call (oldest first), then the new bits
    for (int w = lane; w < kFrameSyncMaxBits / 32 + 3; w += 32) sbuf[w] = 0u;
    __syncwarp();
    for (int p0 = 0; p0 < n + 25; p0 += 32) {
        const int p = p0 + lane;
        int v = 0;
        if (p < 25) v = (int)((reg0 >> (24 - p)) & 1ull);
        else if (p < n + 25) v = b[p - 25] != 0;
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v != 0);
        if (lane == 0) sbuf[p0 >> 5] = word;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        // window of new bit i: buffer bits i .. i+25, oldest first; as a polynomial the newest bit is the constant term
        const int wd = i >> 5, sh = i & 31;
        const uint64_t two = (uint64_t)sbuf[wd] | ((uint64_t)sbuf[wd + 1] << 32);
        const uint32_t w = __brev((uint32_t)(two >> sh) & 0x3FFFFFFu) >> 6;
        const uint32_t S = rds_window_map<10>(w), R = rds_window_map<0>(w);
        sS[i] = (unsigned short)S;
        sR[i] = (unsigned short)R;
        if (syndromes) syndromes[(size_t)s * bits_pitch + i] = (unsigned short)S;
    }
    __syncwarp();
    if (lane != 0) return;
    RdsSyncState st = *stp;
    RdsSyncEvent* ev = events + (size_t)s * events_pitch;
    int nev = 0;
    auto emit = [&](int type, int a, int bb, uint64_t value) {
        if (nev < max_events) ev[nev] = RdsSyncEvent{type, st.rds_bit_cont, a, bb, value};
        nev++;
    };
    const int syn_of[5] = {383, 14, 303, 663, 748};          // A, B, C, D, C'   (:205)
    const int offset_word[5] = {252, 408, 360, 436, 848};    // (:206)
    const int offset_pos[5] = {0, 1, 2, 3, 2};               // (:207)
    for (int i = 0; i < n; i++) {
        st.reg = (st.reg << 1) | (uint64_t)((sbuf[(i + 25) >> 5] >> ((i + 25) & 31)) & 1u);
        if (!st.sync) {
            const int S = sS[i];
            for (int j = 0; j < 5; j++) {
                if (S != syn_of[j]) continue;
                if (!st.prevsync) {
                    st.lastseen_offset = j;
                    st.lastseen_offset_cont = st.rds_bit_cont;
                    st.prevsync = 1;
                } else {
                    st.block_distance = offset_pos[st.lastseen_offset] >= offset_pos[j] ? offset_pos[j] + 4 - offset_pos[st.lastseen_offset]
                                                                                       : offset_pos[j] - offset_pos[st.lastseen_offset];
                    if (st.block_distance * 26 != st.rds_bit_cont - st.lastseen_offset_cont) {
                        st.prevsync = 0;
                    } else {
                        st.wrong_blocks_cont = 0;
                        st.blocks_cont = 0;
                        st.block_bit_cont = 0;
                        st.block_number = (j + 1) & 3;
                        st.group_assembly_started = 0;
                        st.sync = 1;
                        st.ext_good = 0;
                        emit(1, j, st.block_number, 0);
                    }
                }
                break;
            }
        } else if (st.block_bit_cont < 25) {
            st.block_bit_cont++;
        } else {
            const int R = sR[i];
            const uint64_t dataword = (st.reg >> 10) & 0xFFFFull;
            const int bn = st.block_number;
            const bool good = R == offset_word[bn] || (bn == 2 && R == offset_word[4]);
            if (!good) st.wrong_blocks_cont++;
            uint64_t registr = 0;
            if (bn == 0 && good) {
                st.group_assembly_started = 1;
                st.group_good_blocks_cont++;
            }
            if (st.group_assembly_started) {
                if (!good) {
                    st.group_assembly_started = 0;
                } else {
                    registr |= dataword << (48 - bn * 16);
                    st.group_good_blocks_cont++;
                }
                if (st.group_good_blocks_cont == 5) emit(4, 0, 0, registr);
            }
            // extension: the group register the reference meant to keep
            if (!good) st.ext_good = 0;
            else if (bn == 0) { st.ext_reg = dataword << 48; st.ext_good = 1; }
            else if (st.ext_good == bn) { st.ext_reg |= dataword << (48 - bn * 16); st.ext_good++; }
            else st.ext_good = 0;
            if (bn == 3 && st.ext_good == 4) { emit(5, 0, 0, st.ext_reg); st.ext_good = 0; }
            st.block_bit_cont = 0;
            st.block_number = (bn + 1) & 3;
            st.blocks_cont++;
            if (st.blocks_cont == 50) {
                if (st.wrong_blocks_cont > 40) {
                    emit(2, st.wrong_blocks_cont, st.blocks_cont, 0);
                    st.sync = 0;
                    st.prevsync = 0;
                } else {
                    emit(3, st.wrong_blocks_cont, st.blocks_cont, 0);
                }
                st.blocks_cont = 0;
                st.wrong_blocks_cont = 0;
            }
        }
        st.rds_bit_cont++;
    }
    *stp = st;
    nevents[s] = nev;
}

}  // namespace sdrb
