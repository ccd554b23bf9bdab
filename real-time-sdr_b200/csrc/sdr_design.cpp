// Host-side tap designers of libsdr_b200 (init time).  Bit-exact replacements for
// impulseResponseLPF / BPF / APF / RRC, /root/reference/include/filter.h:18-23, src/filter.cpp:13-102.
//
// The reference mixes float and double freely; the tap VALUES are part of the parity contract because
// every later stage is compared bit-for-bit, so each expression below keeps the reference's operand
// types: the normalised cut-off / centre / pass-band are floats, trigonometry is glibc double, and a tap
// is rounded to float BEFORE the window multiplies it (the reference stores h[i] and then does h[i] *= w).
// Compiled without FMA contraction (the reference build has none).
#include <cmath>

#include "../../include/sdr_b200.h"

namespace {
constexpr double kPi = 3.14159265358979323846;  // PI, /root/reference/include/dy4.h:14

// sin^2(i*pi/N) with N taken as a float, the window of src/filter.cpp:27,48,69
inline float windowed(float tap, int i, int num_taps) {
    double w = std::sin(i * kPi / static_cast<float>(num_taps));
    return static_cast<float>(tap * w * w);
}

int lowpass(float Fs, float Fc, int num_taps, bool with_gain, int u, float* h) {
    if (!h || num_taps < 1 || num_taps > 65535 || !(Fs > 0)) return SDRB_ERR_INVALID;
    const float cutoff = static_cast<float>(Fc / (Fs / 2.0));
    const float peak = with_gain ? u * cutoff : cutoff;  // float product u*nc, src/filter.cpp:42,44
    const double centre = (num_taps - 1.0) / 2.0;
    for (int i = 0; i < num_taps; i++) {
        float tap;
        if (i == centre) {
            tap = peak;
        } else {
            const double arg = kPi * cutoff * (i - centre);
            tap = static_cast<float>(peak * std::sin(arg) / arg);
        }
        h[i] = windowed(tap, i, num_taps);
    }
    return SDRB_OK;
}
}  // namespace

extern "C" {

int sdrb_design_lpf(float Fs, float Fc, int num_taps, float* h) { return lowpass(Fs, Fc, num_taps, false, 1, h); }

int sdrb_design_lpf_gain(float Fs, float Fc, int num_taps, int u, float* h) {
    return lowpass(Fs, Fc, num_taps, true, u, h);
}

int sdrb_design_bpf(float Fs, float f_lo, float f_hi, int num_taps, float* h) {
    if (!h || num_taps < 1 || num_taps > 65535 || !(Fs > 0)) return SDRB_ERR_INVALID;
    const float half = Fs / 2;
    const float centre = ((f_hi + f_lo) / 2) / half;
    const float pass = (f_hi - f_lo) / half;
    const int mid_int = (num_taps - 1) / 2;  // integer division inside the sinc, src/filter.cpp:66
    for (int i = 0; i < num_taps; i++) {
        float tap;
        if (i == (num_taps - 1.0) / 2.0) {
            tap = pass;
        } else {
            const double arg = kPi * (pass / 2) * (i - mid_int);
            tap = static_cast<float>(pass * (std::sin(arg) / arg));
        }
        tap = static_cast<float>(tap * std::cos(i * kPi * centre));  // modulated by i, not i-mid (:68)
        h[i] = windowed(tap, i, num_taps);
    }
    return SDRB_OK;
}

int sdrb_design_apf(float gain, int num_taps, float* h) {
    if (!h || num_taps < 1) return SDRB_ERR_INVALID;
    for (int i = 0; i < num_taps; i++) h[i] = 0.0f;
    h[static_cast<int>((num_taps - 1.0) / 2.0)] = gain;
    return SDRB_OK;
}

int sdrb_design_rrc(float Fs, int num_taps, float* h) {
    if (!h || num_taps < 1 || !(Fs > 0)) return SDRB_ERR_INVALID;
    const float Tsym = static_cast<float>(1 / 2375.0);
    const float beta = 0.90f;
    for (int i = 0; i < num_taps; i++) {
        const float t = static_cast<float>((i - static_cast<float>(num_taps) / 2.0) / Fs);
        if (t == 0.0) {
            h[i] = static_cast<float>(1.0 + beta * ((4.0 / kPi) - 1));
        } else if ((t == (-Tsym / (4.0 * beta))) | (t == (Tsym / (4.0 * beta)))) {
            h[i] = static_cast<float>((beta / std::sqrt(2.0)) * ((1 - 2.0 / kPi) * std::sin(kPi / (4.0 * beta))) +
                                      ((1 - 2.0 / kPi) * std::cos(kPi / (4 * beta))));
        } else {
            const double x4 = 4.0 * beta * t / Tsym;
            const double num = std::sin(kPi * t * (1 - beta) / Tsym) + 4.0 * beta * (t / Tsym) * std::cos(kPi * t * (1 + beta) / Tsym);
            const double den = kPi * t * (1 - x4 * x4) / Tsym;
            h[i] = static_cast<float>(num / den);
        }
    }
    return SDRB_OK;
}

}  // extern "C"
