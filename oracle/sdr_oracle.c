/* TEST INFRASTRUCTURE ONLY (oracle/).  See sdr_oracle.h for scope, pinning and who may use this.
 *
 * Plain-C restatement of the reference receive chain.  It keeps the reference's arithmetic exactly:
 * every float/double promotion, the order of every accumulation, and the quirks SURVEY.md 7.3-3 lists.
 * Build: gcc -O2 -ffp-contract=off (x86-64, no -march: no FMA), same libm as the reference build.
 * Citations are relative to /root/reference.
 */
#include "sdr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* include/dy4.h:14 */

/* =====================================================================================
 * Tap designers
 * ===================================================================================== */

/* src/filter.cpp:13-29.  normalized_cutoff is a float; everything else is evaluated in double and
 * rounded to float when stored; the window multiplies the already-rounded tap. */
void orc_lpf(float Fs, float Fc, int num_taps, float* h) {
    float nc = (float)(Fc / (Fs / 2.0));
    double mid = (num_taps - 1.0) / 2.0;
    for (int i = 0; i < num_taps; i++) {
        if (i == mid) {
            h[i] = nc;
        } else {
            double den = ORC_PI * nc * (i - mid);
            h[i] = (float)(nc * sin(den) / den);
        }
        double w = sin(i * ORC_PI / ((float)num_taps));
        h[i] = (float)(h[i] * w * w);
    }
}

/* src/filter.cpp:33-50.  Same, with the integer gain folded in as a FLOAT product u*nc. */
void orc_lpf_gain(float Fs, float Fc, int num_taps, int u, float* h) {
    float nc = (float)(Fc / (Fs / 2.0));
    double mid = (num_taps - 1.0) / 2.0;
    for (int i = 0; i < num_taps; i++) {
        float unc = u * nc;
        if (i == mid) {
            h[i] = unc;
        } else {
            double den = ORC_PI * nc * (i - mid);
            h[i] = (float)(unc * sin(den) / den);
        }
        double w = sin(i * ORC_PI / ((float)num_taps));
        h[i] = (float)(h[i] * w * w);
    }
}

/* src/filter.cpp:55-71.  centre/pass are floats; (num_taps-1)/2 inside the sinc is INTEGER division
 * (:66); the cosine is indexed by i, not i-mid (:68). */
void orc_bpf(float Fs, float f_lo, float f_hi, int num_taps, float* h) {
    float centre = ((f_hi + f_lo) / 2) / (Fs / 2);
    float pass = (f_hi - f_lo) / (Fs / 2);
    for (int i = 0; i < num_taps; i++) {
        if (i == (num_taps - 1.0) / 2.0) {
            h[i] = pass;
        } else {
            int off = i - (num_taps - 1) / 2;
            double a = ORC_PI * (pass / 2) * off;
            h[i] = (float)(pass * (sin(a) / a));
        }
        h[i] = (float)(h[i] * cos(i * ORC_PI * centre));
        double w = sin(i * ORC_PI / ((float)num_taps));
        h[i] = (float)(h[i] * w * w);
    }
}

/* src/filter.cpp:73-78 */
void orc_apf(float gain, int num_taps, float* h) {
    for (int i = 0; i < num_taps; i++) h[i] = 0.0f;
    h[(int)((num_taps - 1.0) / 2.0)] = gain;
}

/* src/filter.cpp:80-102.  t, beta and T_symbol are floats; (1-beta), (1+beta), t/T_symbol and 4*beta
 * (in the singular branch) are float expressions; the rest is double.  The reference also prints i to
 * stderr in the singular branch (:93); it is never reached for (2375*39, 101) or (2375*20, 101). */
void orc_rrc(float Fs, int num_taps, float* h) {
    float T = (float)(1 / 2375.0);
    float beta = 0.90f;
    for (int i = 0; i < num_taps; i++) {
        float t = (float)((i - (float)num_taps / 2.0) / Fs);
        if (t == 0.0) {
            h[i] = (float)(1.0 + beta * ((4.0 / ORC_PI) - 1));
        } else if ((t == (-T / (4.0 * beta))) | (t == (T / (4.0 * beta)))) {
            h[i] = (float)((beta / sqrt(2.0)) * ((1 - 2.0 / ORC_PI) * (sin(ORC_PI / (4.0 * beta)))) +
                           ((1 - 2.0 / ORC_PI) * (cos(ORC_PI / (4 * beta)))));
        } else {
            double a = 4.0 * beta * t / T;
            double num = sin(ORC_PI * t * (1 - beta) / T) + 4.0 * beta * (t / T) * cos(ORC_PI * t * (1 + beta) / T);
            double den = ORC_PI * t * (1 - a * a) / T;
            h[i] = (float)(num / den);
        }
    }
}

/* =====================================================================================
 * Block FIRs
 * ===================================================================================== */

/* src/filter.cpp:106-121.  y[n/d] accumulates h[k]*x[n-k] for k = 0..K-1 in that order, in float,
 * starting from 0; negative indices read the previous call's tail. */
void orc_fir_decim(float* y, const float* x, int nx, const float* h, int nh, float* state, int nstate, int decim) {
    int ny = nx / decim;
    for (int m = 0; m < ny; m++) y[m] = 0.0f;
    for (int n = 0; n < nx; n += decim) {
        float acc = 0.0f;
        for (int k = 0; k < nh; k++) {
            int j = n - k;
            float xv = (j < 0) ? state[j + nstate] : x[j];
            acc = acc + h[k] * xv;
        }
        y[n / decim] = acc;
    }
    /* :119  state <- last nh-1 inputs (here: last nstate; nx >= nstate at every call site) */
    if (nx >= nstate) {
        memcpy(state, x + nx - nstate, (size_t)nstate * sizeof(float));
    } else {
        memmove(state, state + nx, (size_t)(nstate - nx) * sizeof(float));
        memcpy(state + nstate - nx, x, (size_t)nx * sizeof(float));
    }
}

/* src/filter.cpp:123-147.  |y| = nx*up/down; the polyphase branch restarts from phase 0 at n = 0 of
 * EVERY call (:131); taps k = phase, phase+up, ... < K; input index (n*down-k)/up is an exact division. */
int orc_fir_updown(float* y, const float* x, int nx, const float* h, int nh, float* state, int nstate, int up,
                   int down) {
    int ny = (int)(((long)nx * up) / down);
    for (int n = 0; n < ny; n++) {
        int phase = (n * down) % up;
        float acc = 0.0f;
        for (int k = phase; k < nh; k += up) {
            int xi = (n * down - k) / up;
            float xv = (xi < 0) ? state[nstate + xi] : x[xi];
            acc = acc + h[k] * xv;
        }
        y[n] = acc;
    }
    if (nx >= nstate) {
        memcpy(state, x + nx - nstate, (size_t)nstate * sizeof(float));
    } else {
        memmove(state, state + nx, (size_t)(nstate - nx) * sizeof(float));
        memcpy(state + nstate - nx, x, (size_t)nx * sizeof(float));
    }
    return ny;
}

/* =====================================================================================
 * FM discriminator — src/demod.cpp:3-24.  Numerator in float, denominator pow(.,2.0) in double,
 * double divide, rounded to float on store.
 * ===================================================================================== */
void orc_fmdemod(const float* I, const float* Q, int n, float* prev_I, float* prev_Q, float* out) {
    float pi = *prev_I, pq = *prev_Q;
    for (int i = 0; i < n; i++) {
        if ((I[i] == 0) & (Q[i] == 0)) {
            out[i] = 0;
        } else {
            float num = I[i] * (Q[i] - pq) - Q[i] * (I[i] - pi);
            double den = pow(I[i], 2.0) + pow(Q[i], 2.0);
            out[i] = (float)(num / den);
        }
        pi = I[i];
        pq = Q[i];
    }
    *prev_I = pi;
    *prev_Q = pq;
}

/* =====================================================================================
 * PLL — src/pll.cpp:4-61
 * ===================================================================================== */
void orc_pll_init(orc_pll_state* st) { /* src/stereo.cpp:51-57 */
    st->feedbackI = 1.0f;
    st->feedbackQ = 0.0f;
    st->integrator = 0.0f;
    st->phaseEst = 0.0f;
    st->trigOffset = 0.0;
    st->lastCarrier = 1.0f;
}

void orc_pll(const float* in, int n, float freq, float Fs, float* out, orc_pll_state* st, float ncoScale,
             float phaseAdjust, float normBandwidth) {
    float Cp = 2.666f, Ci = 3.555f;                /* :5-6  */
    float Kp = normBandwidth * Cp;                 /* :8    */
    float Ki = normBandwidth * normBandwidth * Ci; /* :9    */
    out[0] = out[n];                               /* :18   */
    for (int i = 0; i < n; i++) {
        float errorI = in[i] * (st->feedbackI);    /* :36 */
        float errorQ = in[i] * (-st->feedbackQ);   /* :37 */
        float errorD = (float)atan2(errorQ, errorI); /* :39, double atan2 */
        st->integrator = st->integrator + Ki * errorD;               /* :41 */
        st->phaseEst = st->phaseEst + Kp * errorD + st->integrator;  /* :42 */
        st->trigOffset += 1.0;                                        /* :46 */
        /* :47  double expression rounded into a FLOAT trigArg */
        float trigArg = (float)(2 * ORC_PI * (freq / Fs) * (st->trigOffset) + st->phaseEst);
        st->feedbackI = (float)cos(trigArg);       /* :49 */
        st->feedbackQ = (float)sin(trigArg);       /* :50 */
        out[i + 1] = (float)cos(trigArg * ncoScale + phaseAdjust); /* :52, float argument */
    }
    st->lastCarrier = out[n]; /* :58 */
}

/* =====================================================================================
 * RDS symbol / bit utilities
 * ===================================================================================== */

/* src/rds_utilities.cpp:4-21.  abs() there is int abs(int): the float is truncated first. */
int orc_cdr(int sps, const float* signal, int n) {
    int maxi = 0, maxv = 0;
    for (int i = 0; i < sps; i++) {
        int sum = 0;
        for (int k = 0; k < n / sps; k++) sum += abs((int)signal[k * sps + i]);
        if (sum > maxv) {
            maxv = sum;
            maxi = i;
        }
    }
    return maxi;
}

/* src/rds_utilities.cpp:34-68 */
int orc_manchester(int* bits, const int* symbols, int nsym, int block_count, int* half_symbol, int* start) {
    int nb = 0;
    if (*start) bits[nb++] = *half_symbol; /* :38-40 */
    if (block_count == 0) {                /* :42-51, never true in the chain (rds.cpp:135) */
        int score = 0;
        for (int i = 0; i < nsym - 1; i += 2) score += symbols[i] ^ symbols[i + 1];
        for (int j = 1; j < nsym - 1; j += 2) score -= symbols[j] ^ symbols[j + 1];
        *start = score < 0;
    }
    for (int i = *start; i < nsym - 1; i += 2) bits[nb++] = symbols[i]; /* :55-59 */
    if ((((unsigned)nsym - (unsigned)*start) & 0x01) == 1) {          /* :61-66 */
        *half_symbol = symbols[nsym - 1];
        *start = 1;
    } else {
        *start = 0;
    }
    return nb;
}

/* src/rds_utilities.cpp:70-88 */
void orc_differential(int* decoded, const int* bits, int n, int* last_bit, int block_num) {
    if (block_num == 0) decoded[0] = bits[0];
    else decoded[0] = bits[0] ^ *last_bit;
    for (int i = 1; i < n; i++) decoded[i] = bits[i] ^ bits[i - 1];
    *last_bit = bits[n - 1];
}

/* =====================================================================================
 * Frame sync and parser
 * ===================================================================================== */

/* src/rds_utilities.cpp:122-133 (one row per syndrome bit) and :135 (A, B, C, C', D) */
static const char* const k_parity_rows[10] = {
    "10000000001001111101100111", "01000000000100111110110011", "00100000001011100010111110",
    "00010000001100001100111000", "00001000000110000110011100", "00000100001010111110101001",
    "00000010001100100010110011", "00000001001111101100111110", "00000000100111110110011111",
    "00000000010011111011001111"};
static const char* const k_syndromes[5] = {"1111011000", "1111010100", "1001011100", "1111001100", "1001011000"};

/* src/rds_utilities.cpp:137-170 */
static const char* const k_pty[32] = {
    "Undefined", "News", "Information", "Sports", "Talk", "Rock", "Classic Rock", "Adult Hits", "Soft Rock",
    "Top 40", "Country", "Oldies", "Soft", "Nostalgia", "Jazz", "Classical", "Rhythm & Blues",
    "Soft Rhythm & Blues", "Language", "Religious Music", "Religious Talk", "Personality", "Public", "College",
    "Spanish Talk", "Spanish Music", "Hip Hop", "Unassigned", "Unassigned", "Weather", "Emergency Test",
    "Emergency"};

void orc_sync_init(orc_sync_state* st) { /* src/rds.cpp:67-70,89-92 */
    memset(st, 0, sizeof(*st));
    st->first_time = 1;
}

/* check_block's syndrome test, src/rds_utilities.cpp:357-366 */
int orc_block_offset(const int* b) {
    char syn[10];
    for (int col = 0; col < 10; col++) {
        int ones = 0;
        for (int i = 0; i < 26; i++) ones += (b[i] != 0) && (k_parity_rows[col][i] == '1');
        syn[col] = (ones % 2) == 1 ? '1' : '0';
    }
    for (int t = 0; t < 5; t++)
        if (memcmp(syn, k_syndromes[t], 10) == 0) return t;
    return -1;
}

static int text_append(char* text, int cap, int pos, const char* s) {
    if (!text) return pos;
    for (; *s && pos < cap - 1; s++) text[pos++] = *s;
    text[pos] = 0;
    return pos;
}

/* src/rds_utilities.cpp:172-199.  "PI: " is printed with std::hex (lower case, no padding), PTY by
 * name, on every group; type-0 groups update the PS buffer and print it when segment 3 arrives and the
 * buffer changed.  The PS string is printed as a C string, i.e. up to its first NUL (:111-119,196). */
int orc_parse(uint64_t bytes, uint64_t* chars, uint64_t* output, char* text, int text_cap) {
    int pos = text ? (int)strlen(text) : 0;
    int start = pos;
    unsigned group_type = (bytes >> 44) & 0xF;
    unsigned placement = (bytes >> 32) & 0x3;
    unsigned PI = (bytes >> 48) & 0xFFFF;
    unsigned pty = (bytes >> 37) & 0x1F;
    char line[64];
    snprintf(line, sizeof line, "PI: %x\n", PI);
    pos = text_append(text, text_cap, pos, line);
    snprintf(line, sizeof line, "PTY: %s\n", k_pty[pty]);
    pos = text_append(text, text_cap, pos, line);
    if (group_type == 0) {
        uint64_t mask = ~((uint64_t)0xFFFF << (48 - 16 * placement));
        *chars = *chars & mask;
        uint64_t word = bytes & (uint64_t)0xFFFF;
        *chars = *chars | (word << 16 * (3 - placement));
        if ((placement == 3) && (*chars != *output)) {
            *output = *chars;
            char str[9];
            for (int i = 0; i < 8; i++) str[7 - i] = (char)((*chars >> (i << 3)) & 0xFF);
            str[8] = 0;
            pos = text_append(text, text_cap, pos, "Program Service: ");
            pos = text_append(text, text_cap, pos, str);
            pos = text_append(text, text_cap, pos, "\n");
        }
    }
    return pos - start;
}

/* src/rds_utilities.cpp:90-109: remainder of x(z) z^10 modulo g(z) = z^10+z^8+z^7+z^5+z^4+z^3+1 (0x5B9), mlen message bits. */
uint64_t orc_calc_syndrome(uint64_t x, uint64_t mlen) {
    uint64_t reg = 0;
    const uint64_t plen = 10;
    for (int i = (int)mlen; i > 0; i--) {
        reg = (reg << 1) | ((x >> (i - 1)) & 0x01);
        if (reg & (1u << plen)) reg ^= 0x5B9;
    }
    for (int i = (int)plen; i > 0; i--) {
        reg <<= 1;
        if (reg & (1u << plen)) reg ^= 0x5B9;
    }
    return reg & ((1u << plen) - 1);
}

void orc_errdet_init(orc_errdet_state* st) {
    memset(st, 0, sizeof *st);
    st->first_time = 1; /* src/rds.cpp:70 */
}

/* src/rds_utilities.cpp:202-311, statement for statement (the reference never calls it; kept as written, quirks included:
 * `registr` is a local that restarts at 0 for every block, group_good_blocks_cont is never reset, so parse() runs once). */
int orc_error_detection(orc_errdet_state* st, const int* bits, int nbits, orc_errdet_event* events, int max_events,
                        char* text, int text_cap, int debug_lines, long long* n_unsynced) {
    static const uint64_t syndromes[5] = {383, 14, 303, 663, 748};
    static const uint64_t offset_word[5] = {252, 408, 360, 436, 848};
    static const int offset_pos[5] = {0, 1, 2, 3, 2};
    int nev = 0;
    int pos = text ? (int)strlen(text) : 0;
    char line[128];
#define ORC_EVENT(t, aa, bb, vv)                                                    \
    do {                                                                            \
        if (nev < max_events) {                                                     \
            events[nev].type = (t); events[nev].bit = st->rds_bit_cont;             \
            events[nev].a = (aa); events[nev].b = (bb); events[nev].value = (vv);   \
        }                                                                           \
        nev++;                                                                      \
    } while (0)
    for (int i = 0; i < nbits; i++) {
        st->reg = (st->reg << 1) | (uint64_t)bits[i]; /* :210, bits are 0/1 ints */
        if (!st->sync) {
            uint64_t reg_syndrome = orc_calc_syndrome(st->reg, 26);
            if (n_unsynced) (*n_unsynced)++;
            if (debug_lines) {
                snprintf(line, sizeof line, st->hex ? "Reg Syndrome: %llx    Reg: %llx\n" : "Reg Syndrome: %llu    Reg: %llu\n",
                         (unsigned long long)reg_syndrome, (unsigned long long)st->reg);
                pos = text_append(text, text_cap, pos, line);
            }
            for (int j = 0; j < 5; j++) {
                if (reg_syndrome == syndromes[j]) {
                    if (!st->prevsync) {
                        st->lastseen_offset = j;
                        st->lastseen_offset_cont = st->rds_bit_cont;
                        st->prevsync = 1;
                    } else {
                        if (offset_pos[st->lastseen_offset] >= offset_pos[j])
                            st->block_distance = offset_pos[j] + 4 - offset_pos[st->lastseen_offset];
                        else
                            st->block_distance = offset_pos[j] - offset_pos[st->lastseen_offset];
                        if ((st->block_distance * 26) != (st->rds_bit_cont - st->lastseen_offset_cont)) {
                            st->prevsync = 0;
                        } else {
                            pos = text_append(text, text_cap, pos, "Sync State Detected\n");
                            st->wrong_blocks_cont = 0;
                            st->blocks_cont = 0;
                            st->block_bit_cont = 0;
                            st->block_number = (j + 1) & 0x03;
                            st->group_assembly_started = 0;
                            st->sync = 1;
                            ORC_EVENT(1, j, st->block_number, 0);
                        }
                    }
                    break;
                }
            }
        } else {
            if (st->block_bit_cont < 25) {
                st->block_bit_cont++;
            } else {
                int good_block = 0;
                uint64_t dataword = (st->reg >> 10) & 0xffff;
                uint64_t block_calculated_crc = orc_calc_syndrome(dataword, 16);
                uint64_t checkword = st->reg & 0x3ff;
                uint64_t block_recieved_crc;
                if (st->block_number == 2) {
                    block_recieved_crc = checkword ^ offset_word[st->block_number];
                    if (block_recieved_crc == block_calculated_crc) {
                        good_block = 1;
                    } else {
                        block_recieved_crc = checkword ^ offset_word[4];
                        if (block_recieved_crc == block_calculated_crc) {
                            good_block = 1;
                        } else {
                            st->wrong_blocks_cont++;
                            good_block = 0;
                        }
                    }
                } else {
                    block_recieved_crc = checkword ^ offset_word[st->block_number];
                    if (block_recieved_crc == block_calculated_crc) {
                        good_block = 1;
                    } else {
                        st->wrong_blocks_cont++;
                        good_block = 0;
                    }
                }
                uint64_t registr = 0;
                if ((st->block_number == 0) & good_block) {
                    st->group_assembly_started = 1;
                    st->group_good_blocks_cont++;
                }
                if (st->group_assembly_started) {
                    if (!good_block) {
                        st->group_assembly_started = 0;
                    } else {
                        /* :285  registr &= (~(0xFFFF) << (48-block_number*16));  ~(0xFFFF) is an int (-65536), shifted as int,
                         * then widened: on the zero register the AND does nothing either way */
                        registr |= (dataword << (48 - st->block_number * 16));
                        st->group_good_blocks_cont++;
                    }
                    if (st->group_good_blocks_cont == 5) {
                        ORC_EVENT(4, 0, 0, registr);
                        orc_parse(registr, &st->chars, &st->output, text, text_cap); /* appends at strlen(text) == pos */
                        pos = text ? (int)strlen(text) : 0;
                        st->hex = 1; /* parse() leaves std::hex set on std::cerr (:180): every later number prints in hex */
                    }
                }
                st->block_bit_cont = 0;
                st->block_number = (st->block_number + 1) & 0x03;
                st->blocks_cont++;
                if (st->blocks_cont == 50) {
                    if (st->wrong_blocks_cont > 40) {
                        snprintf(line, sizeof line, st->hex ? "Lost Sync (Got %x bad blocks on %x total)\n" : "Lost Sync (Got %d bad blocks on %d total)\n",
                                 st->wrong_blocks_cont, st->blocks_cont);
                        pos = text_append(text, text_cap, pos, line);
                        ORC_EVENT(2, st->wrong_blocks_cont, st->blocks_cont, 0);
                        st->sync = 0;
                        st->prevsync = 0;
                    } else {
                        snprintf(line, sizeof line, st->hex ? "Still Sync-ed (Got %x bad blocks on %x total)\n" : "Still Sync-ed (Got %d bad blocks on %d total)\n",
                                 st->wrong_blocks_cont, st->blocks_cont);
                        pos = text_append(text, text_cap, pos, line);
                        ORC_EVENT(3, st->wrong_blocks_cont, st->blocks_cont, 0);
                    }
                    st->blocks_cont = 0;
                    st->wrong_blocks_cont = 0;
                }
            }
        }
        st->rds_bit_cont++;
    }
#undef ORC_EVENT
    return nev;
}

/* src/rds_utilities.cpp:384-400 with check_block (:352-381), uint_copy (:313-337) and
 * isSequenceABCD (:339-350) folded in.  Note `idx < size-26` (strict): the last full window of a
 * call is never tested, it is carried into the next call instead. */
int orc_frame_sync(orc_sync_state* st, const int* bits, int nbits, uint64_t* groups, int max_groups, char* text,
                   int text_cap) {
    int total = st->ncarry + nbits;
    int* s = (int*)malloc((size_t)(total > 0 ? total : 1) * sizeof(int));
    memcpy(s, st->carry, (size_t)st->ncarry * sizeof(int));
    memcpy(s + st->ncarry, bits, (size_t)nbits * sizeof(int));
    int ngroups = 0;
    unsigned idx = 0;
    /* the reference computes size()-26 in unsigned arithmetic; fewer than 26 bits never happens there
     * (15 blocks x >=36 bits), here it simply carries everything */
    unsigned end_range = total >= 26 ? (unsigned)total - 26u : 0u;
    while (idx < end_range) {
        int t = orc_block_offset(s + idx);
        if (t >= 0) {
            if (t != 3) { /* "Cp" matches but copies nothing (:370) */
                int block_type = (t == 4) ? 3 : t;
                uint64_t mask = ~((uint64_t)0xFFFF << (48 - 16 * block_type));
                st->reg &= mask;
                for (int i = 0; i < 16; i++)
                    st->reg |= (uint64_t)(s[idx + i] != 0) << (15 - i + 48 - 16 * block_type);
            }
            if (st->nwindow == 4) { /* deque of the last four offset names */
                st->window[0] = st->window[1];
                st->window[1] = st->window[2];
                st->window[2] = st->window[3];
                st->nwindow = 3;
            }
            st->window[st->nwindow++] = t;
            if (st->nwindow == 4 && st->window[0] == 0 && st->window[1] == 1 && st->window[2] == 2 &&
                st->window[3] == 4) {
                if (ngroups < max_groups) groups[ngroups] = st->reg;
                ngroups++;
                orc_parse(st->reg, &st->chars, &st->output, text, text_cap);
                st->first_time = 0;
            }
            idx += 26;
        } else {
            idx += 1;
        }
    }
    st->ncarry = total - (int)idx;
    if (st->ncarry < 0) st->ncarry = 0;
    memcpy(st->carry, s + idx, (size_t)st->ncarry * sizeof(int));
    free(s);
    return ngroups;
}

/* =====================================================================================
 * Whole chain
 * ===================================================================================== */
#define ORC_TAPS 101
#define ORC_MAXG 16

typedef struct {
    const char* name;
    float* data;
    int count;
} orc_stage;

struct orc_chain {
    orc_chain_info info;
    int rf_Fs, rf_Fc, rf_decim, if_Fs, audio_Fc, up, down, sps, with_rds, rds_on;
    int audio_taps, rds_lpf_taps;
    /* taps */
    float rf_h[ORC_TAPS], pilot_h[ORC_TAPS], stereo_h[ORC_TAPS], apf_h[ORC_TAPS], rds_h[ORC_TAPS],
        rds_pilot_h[ORC_TAPS], rrc_h[ORC_TAPS];
    float *audio_h, *rds_lpf_h;
    /* RF front-end, src/rffrontend.cpp:29-43 */
    float *I, *Q, *I_ds, *Q_ds, *fm_demod, state_I[100], state_Q[100], prev_I, prev_Q;
    /* mono, src/mono.cpp:25-27 */
    float *audio_filt, state_audio[100];
    /* stereo, src/stereo.cpp:16-57 */
    float *pilot, pilot_state[100], *carrier, *stereo_band, stereo_band_state[100], *stereo_dc, *mono_delay,
        mono_delay_state[100], *mono_filt, mono_state[100], *stereo_filt, stereo_state[100];
    orc_pll_state pll19;
    /* rds, src/rds.cpp:26-93 */
    float *rds_band, rds_band_state[100], *rds_band_sq, *gen_pilot, gen_pilot_state[100], *IPLL, *rds_band_delay,
        rds_band_delay_state[100], *rds_dc, *rds_filt, rds_filt_state[100], *rds_clean, rds_clean_state[100];
    orc_pll_state pll114;
    int block_count, half_symbol, start, last_bit, decoder_cont;
    int symbols[256], nsym, man_bits[256], bits[256], nbits, cdr_offset;
    int stream[4096], nstream;
    orc_sync_state sync;
    uint64_t groups[ORC_MAXG];
    int ngroups;
    char* text;
    int text_cap;
    orc_stage stages[24];
    int nstages;
};

static float* falloc(int n) { return (float*)calloc((size_t)n, sizeof(float)); }

static void reg_stage(orc_chain* c, const char* name, float* p, int n) {
    c->stages[c->nstages].name = name;
    c->stages[c->nstages].data = p;
    c->stages[c->nstages].count = n;
    c->nstages++;
}

orc_chain* orc_chain_create(int mode, int type, int with_rds_dsp) {
    orc_chain* c = (orc_chain*)calloc(1, sizeof(orc_chain));
    /* src/project.cpp:31-44 defaults and :67-108 mode table */
    c->rf_Fs = 2400000; c->rf_Fc = 100000; c->rf_decim = 10; c->down = 5; c->up = 1; c->if_Fs = 240000;
    c->audio_Fc = 16000; c->sps = 39;
    switch (mode) {
        case 0: break;
        case 1: c->rf_Fs = 1440000; c->rf_decim = 4; c->down = 9; c->if_Fs = 360000; break;
        case 2: c->down = 800; c->up = 147; c->sps = 20; break;
        case 3: c->rf_Fs = 1152000; c->rf_decim = 3; c->down = 1280; c->if_Fs = 384000; c->up = 147; c->sps = 20; break;
        default: free(c); return NULL;
    }
    c->rds_on = (type == 'r');
    c->with_rds = with_rds_dsp || c->rds_on;
    c->info.mode = mode;
    c->info.type = type;
    c->info.block_pairs = (1470 * c->rf_decim * c->down) / c->up; /* src/rffrontend.cpp:21 */
    c->info.if_block = (1470 * c->down) / c->up;                  /* src/mono.cpp:19       */
    c->info.audio_block = (c->info.if_block * c->up) / c->down;
    c->info.rds_block = (int)(((long)c->info.if_block * 247) / 640);
    int nb = c->info.block_pairs, ni = c->info.if_block, na = c->info.audio_block, nr = c->info.rds_block;

    orc_lpf((float)c->rf_Fs, (float)c->rf_Fc, ORC_TAPS, c->rf_h); /* src/rffrontend.cpp:24 */
    c->audio_taps = ORC_TAPS * c->up;
    c->audio_h = falloc(c->audio_taps);
    orc_lpf_gain((float)(c->if_Fs * c->up), (float)c->audio_Fc, c->audio_taps, c->up, c->audio_h); /* mono.cpp:22 */
    float if_fs = (float)(c->rf_Fs / c->rf_decim);
    orc_apf(1, ORC_TAPS, c->apf_h);                           /* src/stereo.cpp:63 */
    orc_bpf(if_fs, 18.5e3f, 19.5e3f, ORC_TAPS, c->pilot_h);   /* :65 */
    orc_bpf(if_fs, 22e3f, 54e3f, ORC_TAPS, c->stereo_h);      /* :67 */
    c->rds_lpf_taps = ORC_TAPS * 247;
    c->rds_lpf_h = falloc(c->rds_lpf_taps);
    orc_lpf_gain((float)(c->if_Fs * 247), 3e3f, c->rds_lpf_taps, 247, c->rds_lpf_h);   /* src/rds.cpp:61 */
    orc_bpf((float)c->if_Fs, 54e3f, 60e3f, ORC_TAPS, c->rds_h);                        /* :62 */
    orc_bpf((float)c->if_Fs, 113.5e3f, 114.5e3f, ORC_TAPS, c->rds_pilot_h);            /* :63 */
    orc_rrc((float)(2375 * c->sps), ORC_TAPS, c->rrc_h);                               /* :65 */

    c->I = falloc(nb); c->Q = falloc(nb); c->I_ds = falloc(ni); c->Q_ds = falloc(ni); c->fm_demod = falloc(ni);
    c->audio_filt = falloc(na);
    c->pilot = falloc(ni); c->carrier = falloc(ni + 1); c->stereo_band = falloc(ni); c->stereo_dc = falloc(ni);
    c->mono_delay = falloc(ni); c->mono_filt = falloc(na); c->stereo_filt = falloc(na);
    c->carrier[ni] = 1.0f; /* src/stereo.cpp:45 */
    orc_pll_init(&c->pll19);
    c->rds_band = falloc(ni); c->rds_band_sq = falloc(ni); c->gen_pilot = falloc(ni); c->IPLL = falloc(ni + 1);
    c->rds_band_delay = falloc(ni); c->rds_dc = falloc(ni); c->rds_filt = falloc(nr + 1); c->rds_clean = falloc(nr + 1);
    c->IPLL[ni] = 1.0f; /* src/rds.cpp:38 */
    orc_pll_init(&c->pll114);
    orc_sync_init(&c->sync);
    c->text_cap = 1 << 16;
    c->text = (char*)calloc((size_t)c->text_cap, 1);

    reg_stage(c, "I_ds", c->I_ds, ni); reg_stage(c, "Q_ds", c->Q_ds, ni); reg_stage(c, "fm_demod", c->fm_demod, ni);
    reg_stage(c, "audio_filt", c->audio_filt, na); reg_stage(c, "pilot", c->pilot, ni);
    reg_stage(c, "carrier", c->carrier, ni + 1); reg_stage(c, "stereo_band", c->stereo_band, ni);
    reg_stage(c, "stereo_dc", c->stereo_dc, ni); reg_stage(c, "mono_delay", c->mono_delay, ni);
    reg_stage(c, "mono_filt", c->mono_filt, na); reg_stage(c, "stereo_filt", c->stereo_filt, na);
    reg_stage(c, "rds_band", c->rds_band, ni); reg_stage(c, "gen_pilot", c->gen_pilot, ni);
    reg_stage(c, "IPLL", c->IPLL, ni + 1); reg_stage(c, "rds_band_delay", c->rds_band_delay, ni);
    reg_stage(c, "rds_dc", c->rds_dc, ni); reg_stage(c, "rds_filt", c->rds_filt, nr);
    reg_stage(c, "rds_clean", c->rds_clean, nr);
    return c;
}

void orc_chain_destroy(orc_chain* c) {
    if (!c) return;
    float* ptrs[] = {c->audio_h, c->rds_lpf_h, c->I, c->Q, c->I_ds, c->Q_ds, c->fm_demod, c->audio_filt, c->pilot,
                     c->carrier, c->stereo_band, c->stereo_dc, c->mono_delay, c->mono_filt, c->stereo_filt,
                     c->rds_band, c->rds_band_sq, c->gen_pilot, c->IPLL, c->rds_band_delay, c->rds_dc, c->rds_filt,
                     c->rds_clean};
    for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; i++) free(ptrs[i]);
    free(c->text);
    free(c);
}

void orc_chain_get_info(const orc_chain* c, orc_chain_info* info) { *info = c->info; }

/* float -> short as the x86-64 reference build does it: truncate to int32 (cvttss2si; out-of-range
 * and NaN give INT_MIN), keep the low 16 bits.  src/mono.cpp:41, src/stereo.cpp:101-102. */
static int16_t to_pcm(float v) {
    int32_t i;
    if (!(v > -2147483904.0f && v < 2147483648.0f)) i = INT32_MIN;
    else i = (int32_t)v;
    return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}

int orc_chain_block(orc_chain* c, const uint8_t* iq, int16_t* pcm) {
    int nb = c->info.block_pairs, ni = c->info.if_block;
    /* src/rffrontend.cpp:58-63 */
    for (int n = 0; n < nb; n++) {
        c->I[n] = (float)((iq[2 * n] - 128.0) / 128.0);
        c->Q[n] = (float)((iq[2 * n + 1] - 128.0) / 128.0);
    }
    orc_fir_decim(c->I_ds, c->I, nb, c->rf_h, ORC_TAPS, c->state_I, 100, c->rf_decim);   /* :67 */
    orc_fir_decim(c->Q_ds, c->Q, nb, c->rf_h, ORC_TAPS, c->state_Q, 100, c->rf_decim);   /* :68 */
    orc_fmdemod(c->I_ds, c->Q_ds, ni, &c->prev_I, &c->prev_Q, c->fm_demod);              /* :71 */

    if (c->info.type == 'm') {
        /* src/mono.cpp:34-42 */
        int na = orc_fir_updown(c->audio_filt, c->fm_demod, ni, c->audio_h, c->audio_taps, c->state_audio, 100, c->up,
                                c->down);
        for (int i = 0; i < na; i++) pcm[i] = to_pcm(16384 * c->audio_filt[i]);
    } else {
        /* src/stereo.cpp:74-107 */
        orc_fir_decim(c->pilot, c->fm_demod, ni, c->pilot_h, ORC_TAPS, c->pilot_state, 100, 1);
        orc_pll(c->pilot, ni, 19e3f, (float)(c->rf_Fs / c->rf_decim), c->carrier, &c->pll19, 2.0f, 0.0f, 0.01f);
        orc_fir_decim(c->stereo_band, c->fm_demod, ni, c->stereo_h, ORC_TAPS, c->stereo_band_state, 100, 1);
        for (int i = 0; i < ni; i++) c->stereo_dc[i] = (float)(2.0 * c->stereo_band[i] * c->carrier[i]); /* :84 */
        orc_fir_decim(c->mono_delay, c->fm_demod, ni, c->apf_h, ORC_TAPS, c->mono_delay_state, 100, 1);
        int na = orc_fir_updown(c->mono_filt, c->mono_delay, ni, c->audio_h, c->audio_taps, c->mono_state, 100, c->up,
                                c->down);
        orc_fir_updown(c->stereo_filt, c->stereo_dc, ni, c->audio_h, c->audio_taps, c->stereo_state, 100, c->up,
                       c->down);
        for (int i = 0; i < na; i++) {
            pcm[2 * i + 1] = to_pcm(16384 * (c->mono_filt[i] - c->stereo_filt[i])); /* right, odd index  */
            pcm[2 * i] = to_pcm(16384 * (c->mono_filt[i] + c->stereo_filt[i]));     /* left, even index  */
        }
    }

    c->ngroups = 0;
    c->cdr_offset = -1;
    c->nsym = 0;
    c->nbits = 0;
    if (c->with_rds) {
        /* src/rds.cpp:105-133 */
        orc_fir_decim(c->rds_band, c->fm_demod, ni, c->rds_h, ORC_TAPS, c->rds_band_state, 100, 1);
        for (int i = 0; i < ni; i++) c->rds_band_sq[i] = c->rds_band[i] * c->rds_band[i];
        orc_fir_decim(c->gen_pilot, c->rds_band_sq, ni, c->rds_pilot_h, ORC_TAPS, c->gen_pilot_state, 100, 1);
        orc_pll(c->gen_pilot, ni, 114e3f, (float)c->if_Fs, c->IPLL, &c->pll114, 0.5f, 0.0f, 0.001f);
        orc_fir_decim(c->rds_band_delay, c->rds_band, ni, c->apf_h, ORC_TAPS, c->rds_band_delay_state, 100, 1);
        for (int i = 0; i < ni; i++) c->rds_dc[i] = 2 * c->rds_band_delay[i] * c->IPLL[i]; /* :126, float */
        int nr = orc_fir_updown(c->rds_filt, c->rds_dc, ni, c->rds_lpf_h, c->rds_lpf_taps, c->rds_filt_state, 100, 247,
                                640);
        orc_fir_decim(c->rds_clean, c->rds_filt, nr, c->rrc_h, ORC_TAPS, c->rds_clean_state, 100, 1);

        if (c->block_count > 5 && c->rds_on) { /* :135 */
            c->cdr_offset = orc_cdr(c->sps, c->rds_clean, nr);
            c->nsym = 0;
            for (int i = 0; c->cdr_offset + i * c->sps < nr; i++) /* :157-161 */
                c->symbols[c->nsym++] = c->rds_clean[c->cdr_offset + i * c->sps] > 0;
            int nbm = orc_manchester(c->man_bits, c->symbols, c->nsym, c->block_count, &c->half_symbol, &c->start);
            orc_differential(c->bits, c->man_bits, nbm, &c->last_bit, c->block_count);
            c->nbits = nbm;
            c->decoder_cont++;
            memcpy(c->stream + c->nstream, c->bits, (size_t)nbm * sizeof(int));
            c->nstream += nbm;
            if (c->decoder_cont == 15) { /* :184-189 */
                c->ngroups = orc_frame_sync(&c->sync, c->stream, c->nstream, c->groups, ORC_MAXG, c->text, c->text_cap);
                c->decoder_cont = 0;
                c->nstream = 0;
            }
        }
        c->block_count++;
    }
    return 0;
}

const float* orc_chain_stage(const orc_chain* c, const char* name, int* count) {
    for (int i = 0; i < c->nstages; i++)
        if (strcmp(c->stages[i].name, name) == 0) {
            if (count) *count = c->stages[i].count;
            return c->stages[i].data;
        }
    return NULL;
}

int orc_chain_rds_block(const orc_chain* c, int* cdr_offset, const int** symbols, int* nsym, const int** bits,
                        int* nbits) {
    if (cdr_offset) *cdr_offset = c->cdr_offset;
    if (symbols) *symbols = c->symbols;
    if (nsym) *nsym = c->nsym;
    if (bits) *bits = c->bits;
    if (nbits) *nbits = c->nbits;
    return 0;
}

int orc_chain_groups(const orc_chain* c, const uint64_t** groups) {
    if (groups) *groups = c->groups;
    return c->ngroups;
}

const char* orc_chain_text(const orc_chain* c) { return c->text; }

static void orc_pll_jump(orc_pll_state* st, float freq, float Fs, double n0) {
    st->trigOffset = n0;
    float trigArg = (float)(2 * ORC_PI * (freq / Fs) * (st->trigOffset) + st->phaseEst); /* src/pll.cpp:47 */
    st->feedbackI = (float)cos(trigArg);
    st->feedbackQ = (float)sin(trigArg);
}

void orc_chain_set_pll_sample_count(orc_chain* c, double n0) {
    orc_pll_jump(&c->pll19, 19e3f, (float)(c->rf_Fs / c->rf_decim), n0);
    orc_pll_jump(&c->pll114, 114e3f, (float)c->if_Fs, n0);
}

/* ---- multi-threaded batch runner (CPU baseline only) ---- */
typedef struct {
    int mode, type, nstreams, nblocks, tid, nthreads;
    const uint8_t* iq;
    int16_t* pcm;
    int* groups_out;
} orc_job;

static void* orc_worker(void* arg) {
    orc_job* j = (orc_job*)arg;
    for (int s = j->tid; s < j->nstreams; s += j->nthreads) {
        orc_chain* c = orc_chain_create(j->mode, j->type, 1);
        int per = (j->type == 'm' ? 1 : 2) * c->info.audio_block;
        size_t blk_bytes = (size_t)2 * c->info.block_pairs;
        int16_t* scratch = (int16_t*)malloc((size_t)per * sizeof(int16_t));
        int ng = 0;
        for (int b = 0; b < j->nblocks; b++) {
            int16_t* dst = j->pcm ? j->pcm + ((size_t)s * j->nblocks + b) * per : scratch;
            orc_chain_block(c, j->iq + ((size_t)s * j->nblocks + b) * blk_bytes, dst);
            ng += c->ngroups;
        }
        if (j->groups_out) j->groups_out[s] = ng;
        free(scratch);
        orc_chain_destroy(c);
    }
    return NULL;
}

int orc_run_batch(int mode, int type, int nstreams, int nblocks, const uint8_t* iq, int16_t* pcm, int* groups_out,
                  int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nstreams) nthreads = nstreams;
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    orc_job* jobs = (orc_job*)malloc((size_t)nthreads * sizeof(orc_job));
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (orc_job){mode, type, nstreams, nblocks, t, nthreads, iq, pcm, groups_out};
        pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
