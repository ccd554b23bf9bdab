// Host build of real-time-sdr_b200/csrc/pllmath.cuh for the CPU test-suite (no GPU needed).
// Exposes the correctly-rounded trig tiers and scan helpers that compare them with this machine's
// glibc, i.e. with what the reference build computes ((float)sin((double)x) etc.).
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../real-time-sdr_b200/csrc/pllmath.cuh"
#include "../../real-time-sdr_b200/csrc/res_lanes.h"

using namespace sdrb::cr;
static const AtanTab kTab = SDRB_ATAN_TAB_INIT;

static float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t f2bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

extern "C" {

// the lean sine / cosine kernels of the speculative PLL step on reduced arguments (|r| <= pi/4), as doubles
void crh_poly_lean(const double* r, int n, double* s, double* c) {
    PllK kk;
    pll_k_load_lean(kk);
    for (int i = 0; i < n; i++) sincos_poly2_lean(r[i], fabs(r[i]), s[i], c[i], kk);
}
// Error of the speculative phase detector against atan2 in long double, in units of the tolerance its acceptance test
// grants it (2^-42 sa + 2^-48): out[0] = max ratio, out[1] = samples used, out[2] = max |e - ref|.
void crh_head_error(const float* theta, const float* in, int n, double* out) {
    PllK kk;
    pll_k_load_lean(kk);
    double worst = 0, worst_abs = 0, used = 0;
    for (int i = 0; i < n; i++) {
        PllFast f{};
        pll_fast_sincos(theta[i], f);
        if (f.generic_next || in[i] == 0.0f) continue;
        const PllHead h = pll_spec_head(in[i], 1.0 / fabs((double)in[i]), f, kk);
        if (!(fabs(h.e) < 3.14159)) continue;  // the wrap test sends these to the careful path
        const float eI = fmul(in[i], f.fbI), eQ = fmul(in[i], -f.fbQ);
        const long double ref = atan2l((long double)eQ, (long double)eI);
        const double err = (double)fabsl((long double)h.e - ref);
        const double tol = f.sa * 0x1p-42 + 0x1p-48;
        if (err / tol > worst) worst = err / tol;
        if (err > worst_abs) worst_abs = err;
        used += 1;
    }
    out[0] = worst; out[1] = used; out[2] = worst_abs;
}
void crh_sincos(const float* t, int n, float* s, float* c) { for (int i = 0; i < n; i++) sincos_f(t[i], s[i], c[i]); }
void crh_cos(const float* t, int n, float* c) { for (int i = 0; i < n; i++) c[i] = cos_f(t[i]); }
void crh_cos_lean(const float* t, int n, float* c) { for (int i = 0; i < n; i++) c[i] = cos_lean_f(t[i]); }
void crh_atan2(const float* y, const float* x, int n, float* o) { for (int i = 0; i < n; i++) o[i] = atan2_f(y[i], x[i], kTab); }

// force a tier: 0 fast only (no boundary check), 1 slow only
void crh_sincos_tier(const float* t, int n, int tier, float* s, float* c) {
    for (int i = 0; i < n; i++) {
        double ds, dc; bool tiny;
        if (tier == 0) sincos_fast((double)t[i], ds, dc, tiny); else sincos_slow((double)t[i], ds, dc);
        s[i] = (float)ds; c[i] = (float)dc;
    }
}
void crh_atan2_tier(const float* y, const float* x, int n, int tier, float* o) {
    for (int i = 0; i < n; i++) {
        double ax = fabs((double)x[i]), ay = fabs((double)y[i]);
        double v = tier == 0 ? atan2_fast(ax, ay, x[i] < 0, kTab) : atan2_slow(ax, ay, x[i] < 0, kTab);
        float r = (float)v; o[i] = y[i] < 0 ? -r : r;
    }
}

// Scan float bit patterns [lo, hi) with the given stride; compare sincos_f with glibc.
// out[0] mismatches sin, out[1] mismatches cos, out[2] slow-tier entries, out[3] first bad pattern,
// out[4] = min |r| seen as float bits of (float)|r| (for the reduction bound), out[5] count
void crh_scan_sincos(uint32_t lo, uint32_t hi, uint32_t stride, int nthreads, uint64_t* out) {
    std::atomic<uint64_t> bad_s{0}, bad_c{0}, slow{0}, first{0}, cnt{0};
    std::vector<double> minr(nthreads, 1.0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            uint64_t bs = 0, bc = 0, sl = 0, n = 0; double mr = 1.0;
            for (uint64_t b = (uint64_t)lo + (uint64_t)t * stride; b < hi; b += (uint64_t)stride * nthreads) {
                float x = bits2f((uint32_t)b);
                float s, c; sincos_f(x, s, c);
                float rs = (float)sin((double)x), rc = (float)cos((double)x);
                if (f2bits(s) != f2bits(rs)) { bs++; if (!first) first = b; }
                if (f2bits(c) != f2bits(rc)) { bc++; if (!first) first = b; }
                double ds, dc; bool tiny; sincos_fast((double)x, ds, dc, tiny);
                if (tiny || near_float_boundary(ds) || near_float_boundary(dc)) sl++;
                double kd = rint((double)x * kTwoOverPi);
                if (kd != 0) { double r = fabs(fmin(fabs(ds), fabs(dc))); if (r < mr) mr = r; }
                n++;
            }
            bad_s += bs; bad_c += bc; slow += sl; cnt += n; minr[t] = mr;
        });
    for (auto& x : th) x.join();
    double mr = 1.0; for (double v : minr) if (v < mr) mr = v;
    out[0] = bad_s; out[1] = bad_c; out[2] = slow; out[3] = first; memcpy(&out[4], &mr, 8); out[5] = cnt;
}

// Random (y, x) pairs from a 64-bit LCG seeded per thread; mode 0: arbitrary floats with exponents in a
// band, mode 1: PLL-like (y = in*-sin, x = in*cos).  out[0] mismatches, out[1] slow entries, out[2] count
void crh_scan_atan2(uint64_t seed, uint64_t n_per_thread, int mode, int nthreads, uint64_t* out) {
    std::atomic<uint64_t> bad{0}, slow{0}, cnt{0}, firsty{0}, firstx{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            uint64_t s = seed * 0x9E3779B97F4A7C15ull + (uint64_t)t * 0xD1B54A32D192ED03ull + 1;
            auto next = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 32); };
            uint64_t b = 0, sl = 0;
            for (uint64_t i = 0; i < n_per_thread; i++) {
                float y, x;
                if (mode == 0) {
                    uint32_t a = next(), c = next();
                    // sign | exponent in [100, 140] | mantissa
                    y = bits2f((a & 0x807FFFFFu) | ((100u + (a >> 23) % 41u) << 23));
                    x = bits2f((c & 0x807FFFFFu) | ((100u + (c >> 23) % 41u) << 23));
                } else {
                    float in = bits2f((next() & 0x807FFFFFu) | ((110u + next() % 17u) << 23));
                    float th_ = (float)(next() * (3.0e6 / 4294967296.0));
                    float fs = (float)sin((double)th_), fc = (float)cos((double)th_);
                    y = in * -fs; x = in * fc;
                }
                float r = atan2_f(y, x, kTab);
                float g = (float)atan2((double)y, (double)x);
                if (f2bits(r) != f2bits(g)) { b++; if (!firsty) { firsty = f2bits(y); firstx = f2bits(x); } }
                if (y != 0 && x != 0) {
                    double v = atan2_fast(fabs((double)x), fabs((double)y), x < 0, kTab);
                    if (near_float_boundary(v)) sl++;
                }
            }
            bad += b; slow += sl; cnt += n_per_thread;
        });
    for (auto& x : th) x.join();
    out[0] = bad; out[1] = slow; out[2] = cnt; out[3] = firsty; out[4] = firstx;
}

// The PLL loop built from pll_step; same calling convention as the reference's fmpll.
void crh_pll(const float* in, int n, float freq, float Fs, float scale, float adjust, float bw, float* out,
             float* st4, double* trig) {
    PllCoef k = pll_coef(freq, Fs, scale, adjust, bw);
    PllState st{st4[0], st4[1], st4[2], st4[3], *trig};
    out[0] = out[n];
    for (int i = 0; i < n; i++) out[i + 1] = pll_step(in[i], st, k, kTab);
    st4[0] = st.feedbackI; st4[1] = st.feedbackQ; st4[2] = st.integrator; st4[3] = st.phaseEst; *trig = st.trigOffset;
}

// The short-chain recurrence the batched kernel runs (pll_step_fast); also returns how many steps fell back
// to the general atan2 / sincos (stats[0], stats[1]).
void crh_pll_fast(const float* in, int n, float freq, float Fs, float scale, float adjust, float bw, float* out,
                  float* st4, double* trig, uint64_t* stats) {
    const bool unrotated = getenv("SDRB_PLL_UNROTATED") != nullptr;  // the round-1 chunk (pll_chunk4), kept for comparison
    PllCoef k = pll_coef(freq, Fs, scale, adjust, bw);
    PllState st{st4[0], st4[1], st4[2], st4[3], *trig};
    PllFast f;
    pll_fast_load(f, st, k);
    out[0] = out[n];
    uint64_t gen_atan = 0, gen_sc = 0;
    PllK kk;
    if (unrotated) pll_k_load(kk); else pll_k_load_lean(kk);
    int i = 0;
    auto recip = [](float v) { return 1.0 / fabs((double)v); };  // device: rcp.approx.ftz of |v|; the chunk itself rejects out-of-range samples
    if (unrotated) {
        for (; i + 4 <= n; i += 4) {  // round-1 chunking: 4 speculative steps, verified once
            float c[4] = {in[i], in[i + 1], in[i + 2], in[i + 3]}, th[4];
            double r[4] = {recip(c[0]), recip(c[1]), recip(c[2]), recip(c[3])};
            PllFast probe = f;
            unsigned bad = f.generic_next ? 1u : 0u;
            for (int j = 0; j < 4; j++) pll_step_spec(c[j], r[j], probe, k, kk, bad);
            if (bad) gen_atan++;  // chunks that needed the careful path
            pll_chunk4(c[0], c[1], c[2], c[3], r[0], r[1], r[2], r[3], f, k, kk, kTab, th[0], th[1], th[2], th[3]);
            if (f.generic_next) gen_sc++;
            for (int j = 0; j < 4; j++) out[i + j + 1] = nco_out(th[j], k);
        }
    } else if (n >= 4) {  // the kernel's loop: rotated chunks, the next chunk's first phase detector evaluated ahead
        PllHead h;
        pll_chunk4r_prime(in[0], recip(in[0]), f, kk, h);
        for (; i + 4 <= n; i += 4) {
            float c[4] = {in[i], in[i + 1], in[i + 2], in[i + 3]}, th[4];
            double r[4] = {recip(c[0]), recip(c[1]), recip(c[2]), recip(c[3])};
            const float n0 = i + 4 < n ? in[i + 4] : 1.0f;
            unsigned long long redo = 0;
            pll_chunk4r(c[0], c[1], c[2], c[3], r[0], r[1], r[2], r[3], n0, recip(n0), f, h, k, kk, kTab, th[0], th[1], th[2], th[3], &redo);
            gen_atan += redo;
            if (f.generic_next) gen_sc++;
            for (int j = 0; j < 4; j++) out[i + j + 1] = nco_out(th[j], k);
        }
        if (!f.generic_next) pll_fast_resync(f, k);
    }
    for (; i < n; i++) {
        float th = pll_step_fast(in[i], pll_guard_recip(in[i], 1.0 / fabs((double)in[i])), f, k, kTab);
        out[i + 1] = nco_out(th, k);
    }
    pll_fast_store(f, st);
    st4[0] = st.feedbackI; st4[1] = st.feedbackQ; st4[2] = st.integrator; st4[3] = st.phaseEst; *trig = st.trigOffset;
    if (stats) { stats[0] += gen_atan; stats[1] += gen_sc; }
}

// The 247/640 RDS resampler of k_rds_backend replayed on the host: the lane tables of res_lanes.h (the same function the
// chain uploads) and the kernel's per-thread loop, statement by statement (real-time-sdr_b200/csrc/sdr_kernels.cuh).
// x: [100 + n_in] the carried state followed by the block; every float the kernel may load beyond that is NaN here, so a
// load that must not matter shows up if it does.  y: [n_out].  info[0] = pairs without a conflict-free lane, info[1] = the
// largest number of lanes of one warp whose loads fall into the same bank, info[2] = outputs written more than once,
// info[3] = outputs never written.  Returns 0, or -1 if the lanes ran out.
int crh_resampler_lanes(const float* lh, const float* x, int n_in, int n_out, float* y, int* info) {
    using namespace sdrb;
    constexpr int kState = kResTaps - 1;
    std::vector<ResLane> lanes(kResLanes);
    std::vector<float> rtaps((size_t)kResIter * kResLanes * 2);
    const int left = build_res_lanes(lh, lanes.data(), rtaps.data());
    if (left < 0) return -1;
    info[0] = left; info[1] = 0; info[2] = 0; info[3] = 0;
    for (int w = 0; w < kResLanes / 32; w++) {
        int per_bank[32] = {0};
        for (int l = 0; l < 32; l++)
            if (lanes[32 * w + l].tp_hi >= 0) per_bank[lanes[32 * w + l].base % 32]++;
        for (int b = 0; b < 32; b++) info[1] = std::max(info[1], per_bank[b]);
    }
    const int need = kState + kRdsDown * kResQ + kResLagMax + 1, have = n_in + kState;
    std::vector<float> sdc((std::max(need, have) + 3) / 4 * 4, NAN);
    memcpy(sdc.data(), x, sizeof(float) * have);
    std::vector<int> written(n_out, 0);
    constexpr int HQ = kResQ / 2;
    for (int t = 0; t < 2 * kResLanes; t++) {
        const int tl = t & (kResLanes - 1), qh = t / kResLanes;
        const ResLane ln = lanes[tl];
        if (ln.tp_hi < 0) continue;
        const int lag_hi = ln.lags & 0xFF, lag_lo = (ln.lags >> 8) & 0xFF;
        const bool has_lo = ln.tp_lo >= 0;
        const float* xb = sdc.data() + kState + ln.base + kRdsDown * HQ * qh;
        float aH[HQ], aL[HQ];
        for (int q = 0; q < HQ; q++) aH[q] = aL[q] = 0.0f;
        for (int j = 0; j < kResIter; j++) {
            const float hx = rtaps[((size_t)j * kResLanes + tl) * 2], hy = rtaps[((size_t)j * kResLanes + tl) * 2 + 1];
            bool vH = true, vL = true;  // the uniform steps do both MACs whatever the lane is
            if (j < kResLagMax) { vH = j >= lag_hi; vL = has_lo && j >= lag_lo; }
            if (j >= kResTaps) { vH = j < kResTaps + lag_hi; vL = has_lo && j < kResTaps + lag_lo; }
            for (int q = 0; q < HQ; q++) {
                const float xv = xb[kRdsDown * q - j];
                if (vH) aH[q] = fadd(aH[q], fmul(hx, xv));
                if (vL) aL[q] = fadd(aL[q], fmul(hy, xv));
            }
        }
        for (int r = 0; r < 2; r++) {
            const int tp = r ? ln.tp_lo : ln.tp_hi;
            if (tp < 0) continue;
            for (int q = 0; q < HQ; q++) {
                const int n = (HQ * qh + q) * kRdsUp + tp;
                if (n < n_out) { y[n] = r ? aL[q] : aH[q]; written[n]++; }
            }
        }
    }
    for (int n = 0; n < n_out; n++) { info[2] += written[n] > 1; info[3] += written[n] == 0; }
    return 0;
}

// The discriminator's division (fm_quotient_fast) against the exact form (float)((double)num / den),
// /root/reference/src/demod.cpp:11,17.  num[], I[], Q[]: n operands; den = RN((double)I^2 + (double)Q^2) as the kernel forms
// it.  The reciprocal seed is the correctly rounded float reciprocal moved by `ulps` float-ulps (the device's
// rcp.approx is within one).  out[0] = quotients the fast path accepted, out[1] = accepted and different from the exact
// form (must be 0), out[2] = operands outside the fast path's range.
void crh_fm_quotient_scan(const float* num, const float* I, const float* Q, int n, int ulps, uint64_t* out) {
    uint64_t acc = 0, bad = 0, outside = 0;
    for (int i = 0; i < n; i++) {
        const double den = dadd(dmul((double)I[i], (double)I[i]), dmul((double)Q[i], (double)Q[i]));
        const float exact = (float)((double)num[i] / den);
        if (!fm_den_in_range(den)) { outside++; continue; }
        const float df = fm_den_approx_f(den);
        float rc = 1.0f / df;
        rc = bits2f(f2bits(rc) + (uint32_t)ulps);
        float got;
        if (!fm_quotient_fast(num[i], den, rc, got)) continue;
        acc++;
        if (f2bits(got) != f2bits(exact)) bad++;
    }
    out[0] = acc; out[1] = bad; out[2] = outside;
}
}
