/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement of the reference receive chain.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (real-time-sdr_b200/) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no test of this path, so the
 * restatement is pinned against the reference itself: oracle/_ref/ref_harness is the unmodified
 * reference sources compiled in this container, tests/test_oracle_vs_ref.py compares every function
 * and every stage of the chain bit-for-bit, and tests/golden/ holds fixtures produced by that
 * harness (tests/golden/make_golden.py) for machines where /root/reference is absent.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- tap designers (src/filter.cpp:13-102) ---- */
void orc_lpf(float Fs, float Fc, int num_taps, float* h);                 /* :13-29  */
void orc_lpf_gain(float Fs, float Fc, int num_taps, int u, float* h);     /* :33-50  */
void orc_bpf(float Fs, float f_lo, float f_hi, int num_taps, float* h);   /* :55-71  */
void orc_apf(float gain, int num_taps, float* h);                         /* :73-78  */
void orc_rrc(float Fs, int num_taps, float* h);                           /* :80-102 */

/* ---- block FIRs with carried state (src/filter.cpp:106-147) ----
 * `state` holds the last `nstate` input samples of the previous call (zeros before the first call)
 * and is updated in place.  nstate = nh-1 for the decimator; the resampler only ever reads
 * floor((nh-1)/up) entries back (filter.cpp:135), which is 100 at every call site. */
void orc_fir_decim(float* y, const float* x, int nx, const float* h, int nh, float* state, int nstate, int decim);
int orc_fir_updown(float* y, const float* x, int nx, const float* h, int nh, float* state, int nstate, int up,
                   int down); /* returns the number of outputs, nx*up/down */

/* ---- FM discriminator (src/demod.cpp:3-24) ---- */
void orc_fmdemod(const float* I, const float* Q, int n, float* prev_I, float* prev_Q, float* out);

/* ---- PLL + NCO (include/pll.h:10-17, src/pll.cpp:4-61) ---- */
typedef struct {
    float feedbackI, feedbackQ, integrator, phaseEst;
    double trigOffset;
    float lastCarrier;
} orc_pll_state;
void orc_pll_init(orc_pll_state* st);
/* out has n+1 entries; out[0] <- out[n] of the previous call (callers seed out[n] = 1) */
void orc_pll(const float* in, int n, float freq, float Fs, float* out, orc_pll_state* st, float ncoScale,
             float phaseAdjust, float normBandwidth);

/* ---- RDS symbol / bit utilities (src/rds_utilities.cpp:4-88) ---- */
int orc_cdr(int sps, const float* signal, int n);
int orc_manchester(int* bits, const int* symbols, int nsym, int block_count, int* half_symbol, int* start);
void orc_differential(int* decoded, const int* bits, int n, int* last_bit, int block_num);

/* ---- frame sync and group parser (src/rds_utilities.cpp:111-199, 313-400) ---- */
typedef struct {
    int carry[64];
    int ncarry;
    uint64_t reg, chars, output;
    int first_time;
    int window[4]; /* last offsets seen: 0 A, 1 B, 2 C, 3 C', 4 D */
    int nwindow;
} orc_sync_state;
void orc_sync_init(orc_sync_state* st);
/* 0 A, 1 B, 2 C, 3 C', 4 D, -1 none (check_block, :352-381, without its side effects) */
int orc_block_offset(const int* bits26);
/* One start_frame_sync call.  groups[] receives `reg` at every A,B,C,D completion (max_groups entries),
 * text (may be NULL) receives what the reference prints on stderr; returns the number of groups. */
int orc_frame_sync(orc_sync_state* st, const int* bits, int nbits, uint64_t* groups, int max_groups, char* text,
                   int text_cap);
/* parse (:172-199) on one group register; appends to text; returns bytes written */
int orc_parse(uint64_t reg, uint64_t* chars, uint64_t* output, char* text, int text_cap);

/* ---- error_detection, the sync-state-machine decoder the reference declares (include/rds_utilities.h:14) and defines
 * (src/rds_utilities.cpp:202-311, calc_syndrome :90-109) but never calls (src/rds.cpp:177-179 is commented out) ---- */
typedef struct {
    uint64_t reg, chars, output;
    int first_time;
    int hex; /* std::hex is sticky on std::cerr once parse() ran (src/rds_utilities.cpp:180) */
    int sync, prevsync, lastseen_offset, rds_bit_cont, lastseen_offset_cont, block_distance, block_number, block_bit_cont,
        blocks_cont, wrong_blocks_cont, group_assembly_started, group_good_blocks_cont;
} orc_errdet_state;
typedef struct {
    int type;       /* 1 "Sync State Detected" (a = matched offset 0 A 1 B 2 C 3 D 4 C', b = next block number),
                       2 "Lost Sync" (a wrong, b total), 3 "Still Sync-ed" (a wrong, b total), 4 parse(registr) (value) */
    int bit;        /* rds_bit_cont when it happened */
    int a, b;
    uint64_t value;
} orc_errdet_event;
void orc_errdet_init(orc_errdet_state* st);  /* the initial values of src/rds.cpp:67-86 */
uint64_t orc_calc_syndrome(uint64_t x, uint64_t mlen);
/* One error_detection call over nbits bits.  events (max_events entries) receives what happened; text (may be NULL)
 * what the reference prints on stderr, the per-bit "Reg Syndrome" debug lines included when debug_lines != 0;
 * *n_unsynced (may be NULL) is incremented by the number of bits processed out of sync (= debug lines).
 * Returns the number of events. */
int orc_error_detection(orc_errdet_state* st, const int* bits, int nbits, orc_errdet_event* events, int max_events,
                        char* text, int text_cap, int debug_lines, long long* n_unsynced);

/* ---- whole receive chain, one stream (src/rffrontend.cpp, mono.cpp, stereo.cpp, rds.cpp) ---- */
typedef struct orc_chain orc_chain;

typedef struct {
    int mode, type;            /* type: 'm', 's' or 'r' */
    int block_pairs;           /* IQ pairs per block */
    int if_block;              /* samples at the IF rate per block */
    int audio_block;           /* audio frames per block (1470) */
    int rds_block;             /* samples per block at the RDS rate (if_block*247/640) */
} orc_chain_info;

orc_chain* orc_chain_create(int mode, int type, int with_rds_dsp);
void orc_chain_destroy(orc_chain* c);
void orc_chain_get_info(const orc_chain* c, orc_chain_info* info);
/* Process one block of 2*block_pairs bytes.  pcm receives audio_block (mono) or 2*audio_block (stereo)
 * int16 samples.  Returns 0. */
int orc_chain_block(orc_chain* c, const uint8_t* iq, int16_t* pcm);
/* Pointers to the intermediates of the last block (valid until the next call); NULL if unknown.
 * Names: I_ds Q_ds fm_demod audio_filt pilot carrier stereo_band stereo_dc mono_delay mono_filt stereo_filt
 *        rds_band gen_pilot IPLL rds_band_delay rds_dc rds_filt rds_clean */
const float* orc_chain_stage(const orc_chain* c, const char* name, int* count);
/* RDS decode results of the last block: cdr offset (-1 if gated), symbols, decoded bits */
int orc_chain_rds_block(const orc_chain* c, int* cdr_offset, const int** symbols, int* nsym, const int** bits,
                        int* nbits);
/* groups completed during the last block (only on frame-sync blocks) */
int orc_chain_groups(const orc_chain* c, const uint64_t** groups);
/* all stderr text so far */
const char* orc_chain_text(const orc_chain* c);

/* Test hook: puts both PLLs of the chain at sample count n0 (trigOffset = n0, feedbackI/Q = cos/sin of the phase that
 * src/pll.cpp:47 gives for the carried phaseEst), as if n0 samples had already gone through.  Lets a short test reach the
 * regime hours into a run where the float NCO phase has an ulp of radians. */
void orc_chain_set_pll_sample_count(orc_chain* c, double n0);

/* Run `nstreams` independent chains over `nblocks` blocks each on `nthreads` host threads (CPU baseline).
 * iq: [nstreams][nblocks*2*block_pairs]; pcm (may be NULL): [nstreams][nblocks*pcm_per_block];
 * groups_out (may be NULL): [nstreams] group counts.  Returns 0. */
int orc_run_batch(int mode, int type, int nstreams, int nblocks, const uint8_t* iq, int16_t* pcm, int* groups_out,
                  int nthreads);

#ifdef __cplusplus
}
#endif
#endif
