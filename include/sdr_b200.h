/* sdr_b200 — C ABI of the B200 receive-chain library (libsdr_b200.so).
 *
 * Drop-in boundary for the block-based FM receive chain of TheZxc07/real-time-SDR.  The reference has
 * no FFI layer: its boundary is a set of C++ free functions (include/filter.h, demod.h, pll.h,
 * rds_utilities.h) and four thread bodies taking `args*` (include/rffrontend.h, mono.h, stereo.h, rds.h).
 * real-time-sdr_b200/host/dy4_api.h re-declares exactly those C++ signatures and implements them on top
 * of the entry points below; INTEGRATION.md shows the binding.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions: plain pointers and sizes only; every function returns an sdrb_status (0 = ok) and never
 * terminates the process (the reference calls exit(1) on EOF, src/rffrontend.cpp:50-52).  Pointers named
 * d_* are device pointers, h_* are host pointers.  `stream` arguments are CUDA streams passed as void*
 * (NULL = the default stream).  All batched arrays are stream-major: row s starts at base + s*pitch
 * (pitch in ELEMENTS of the array's type).
 */
#ifndef SDR_B200_H
#define SDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDRB_TAPS 101 /* rf_taps, src/project.cpp:34 */

typedef enum {
    SDRB_OK = 0,
    SDRB_ERR_INVALID = 1,   /* bad argument / unsupported configuration */
    SDRB_ERR_CUDA = 2,      /* a CUDA call failed; sdrb_last_error() has the text */
    SDRB_ERR_NO_DEVICE = 3, /* no usable GPU: there is no CPU fallback */
    SDRB_ERR_STATE = 4      /* call order violated (e.g. reading results before a block was processed) */
} sdrb_status;

const char* sdrb_last_error(void);
int sdrb_version(void);

/* ------------------------------------------------------------------------------------------------
 * 1. Tap designers (host, init time).  Bit-exact replacements for
 *    impulseResponseLPF/BPF/APF/RRC, include/filter.h:18-23, src/filter.cpp:13-102.
 * ------------------------------------------------------------------------------------------------ */
int sdrb_design_lpf(float Fs, float Fc, int num_taps, float* h);              /* filter.h:18 */
int sdrb_design_lpf_gain(float Fs, float Fc, int num_taps, int u, float* h);  /* filter.h:19 */
int sdrb_design_bpf(float Fs, float f_lo, float f_hi, int num_taps, float* h);/* filter.h:21 */
int sdrb_design_apf(float gain, int num_taps, float* h);                      /* filter.h:22 */
int sdrb_design_rrc(float Fs, int num_taps, float* h);                        /* filter.h:23 */

/* ------------------------------------------------------------------------------------------------
 * 2. Batched stage primitives on device memory.  Each processes ONE block for n_streams independent
 *    streams and carries the same state the reference function carries, but as explicit arrays.
 * ------------------------------------------------------------------------------------------------ */

/* convolveFIR(y, x, h, state, decim) — include/filter.h:20, src/filter.cpp:106-121.
 * d_x [n_streams][x_pitch] (nx samples used), d_state [n_streams][nh-1] in/out, d_y [n_streams][y_pitch]
 * receives nx/decim samples.  nh <= 128. */
int sdrb_fir_decim(const float* d_x, size_t x_pitch, int nx, const float* h_taps, int nh, float* d_state,
                   float* d_y, size_t y_pitch, int decim, int n_streams, void* stream);

/* convolveFIR(y, x, h, state, up, down) — include/filter.h:24, src/filter.cpp:123-147.
 * h_taps holds the full prototype (nh = 101*up at every reference call site); d_state [n_streams][nstate]
 * with nstate = (nh-1)/up; d_y receives nx*up/down samples. */
int sdrb_fir_updown(const float* d_x, size_t x_pitch, int nx, const float* h_taps, int nh, float* d_state,
                    int nstate, float* d_y, size_t y_pitch, int up, int down, int n_streams, void* stream);

/* fmDemodNoArctan(I, Q, prev_I, prev_Q, out) — include/demod.h:5-6, src/demod.cpp:3-24.
 * d_prev [n_streams][2] = {prev_I, prev_Q} in/out. */
int sdrb_fm_demod(const float* d_I, const float* d_Q, size_t iq_pitch, int n, float* d_prev, float* d_out,
                  size_t out_pitch, int n_streams, void* stream);

/* fmpll(in, freq, Fs, out, state, ncoScale, phaseAdjust, normBandwidth) — include/pll.h:10-20,
 * src/pll.cpp:4-61.  State mirrors pllblock_args; `last_out` is the element the reference keeps in
 * pllOut[N] between calls (src/pll.cpp:18).  d_out [n_streams][out_pitch] receives N+1 samples. */
typedef struct {
    float feedbackI, feedbackQ, integrator, phaseEst;
    double trigOffset;
    float lastCarrier;
    float last_out;
} sdrb_pll_state;
int sdrb_pll(const float* d_in, size_t in_pitch, int n, float freq, float Fs, float ncoScale, float phaseAdjust,
             float normBandwidth, sdrb_pll_state* d_state, float* d_out, size_t out_pitch, int n_streams,
             void* stream);

/* cdr(sps, signal) — include/rds_utilities.h:6, src/rds_utilities.cpp:4-21.  d_offset [n_streams]. */
int sdrb_cdr(const float* d_x, size_t x_pitch, int n, int sps, int* d_offset, int n_streams, void* stream);

/* manchester_decode(bits, symbols, block_count, half_symbol, start) — include/rds_utilities.h:8-9,
 * src/rds_utilities.cpp:34-68.  d_symbols [n_streams][sym_pitch] (0/1), d_nsym [n_streams] symbols used per stream;
 * d_state [n_streams] carries half_symbol/start in and out; d_bits [n_streams][bits_pitch] receives d_nbits[s]
 * bits (at most 1 + nsym/2).  block_count == 0 re-estimates the pairing phase first (:42-51), as the reference does. */
typedef struct {
    int32_t half_symbol, start;
} sdrb_manchester_state;
int sdrb_manchester_decode(const int32_t* d_symbols, size_t sym_pitch, const int32_t* d_nsym, int block_count,
                           sdrb_manchester_state* d_state, int32_t* d_bits, size_t bits_pitch, int32_t* d_nbits, int n_streams,
                           void* stream);

/* differential_decode(decoded, bits, last_bit, block_num) — include/rds_utilities.h:11-12,
 * src/rds_utilities.cpp:70-88.  d_last_bit [n_streams] in/out; d_decoded receives d_nbits[s] bits.  A stream with
 * d_nbits[s] == 0 (undefined behaviour in the reference) is left untouched. */
int sdrb_differential_decode(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int block_num, int32_t* d_last_bit,
                             int32_t* d_decoded, size_t dec_pitch, int n_streams, void* stream);

/* start_frame_sync(idx = 0, stream, carry, reg, ...) with check_block — include/rds_utilities.h:14-19,
 * src/rds_utilities.cpp:352-400.  The d_nbits[s] (<= max_nbits <= 8128) new bits of each stream are appended to the
 * carried tail in d_state; d_groups [n_streams][groups_pitch] receives the group register at every A,B,C,D completion
 * (the value the reference hands to parse(), see sdrb_rds_parse below), d_ngroups [n_streams] their number (which may
 * exceed max_groups: only the first max_groups are stored).  A zero-initialised state is the reference's initial state. */
typedef struct {
    uint64_t reg;      /* src/rds.cpp:67 */
    int32_t window[4]; /* last matched offsets, 0 A, 1 B, 2 C, 3 C', 4 D (src/rds.cpp:70) */
    int32_t nwindow;
    int32_t ncarry;
    uint8_t carry[64]; /* unread tail of the previous call (src/rds_utilities.cpp:398-399) */
} sdrb_framesync_state;
int sdrb_frame_sync(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int max_nbits, sdrb_framesync_state* d_state,
                    uint64_t* d_groups, size_t groups_pitch, int32_t* d_ngroups, int max_groups, int n_streams, void* stream);

/* error_detection(reg, chars, output, first_time, sync, prevsync, ..., decoded_bits) - include/rds_utilities.h:14,
 * src/rds_utilities.cpp:202-311 (with calc_syndrome :90-109): the sync-state-machine decoder the reference declares and
 * defines but never calls (src/rds.cpp:177-179 is commented out).  The d_nbits[s] (<= max_nbits <= 8160) bits of each stream
 * go through it; the state mirrors the function's reference parameters (a zero-initialised state is src/rds.cpp:67-84).
 * d_events [n_streams][events_pitch] receives what the reference prints, as records: type 1 "Sync State Detected"
 * (a = matched offset 0 A 1 B 2 C 3 D 4 C', b = number of the next block), 2 "Lost Sync (Got a bad blocks on b total)",
 * 3 "Still Sync-ed (Got a ... b ...)", 4 the function's call of parse(value) (it makes exactly one, with a register that
 * holds a single block: `registr` restarts at zero for every block, :280), all with bit = rds_bit_cont at that moment.
 * EXTENSION beyond the reference: type 5 carries value = A|B|C|D of every group whose four blocks passed the check in
 * order (what `registr` was meant to collect); feed it to sdrb_rds_parse.  d_nevents [n_streams] may exceed max_events
 * (only the first max_events are stored).  d_syndromes (may be NULL) [n_streams][bits_pitch] receives calc_syndrome(reg, 26)
 * after every bit (the value of the reference's per-bit debug line while it searches for sync). */
typedef struct {
    uint64_t reg;
    int32_t sync, prevsync, lastseen_offset, rds_bit_cont, lastseen_offset_cont, block_distance, block_number, block_bit_cont,
        blocks_cont, wrong_blocks_cont, group_assembly_started, group_good_blocks_cont;
    uint64_t ext_reg; /* extension state: group register, blocks of the current group that passed */
    int32_t ext_good;
    int32_t reserved;
} sdrb_rds_sync_state;
typedef struct {
    int32_t type, bit, a, b;
    uint64_t value;
} sdrb_rds_sync_event;
int sdrb_rds_sync(const int32_t* d_bits, size_t bits_pitch, const int32_t* d_nbits, int max_nbits, sdrb_rds_sync_state* d_state,
                  sdrb_rds_sync_event* d_events, size_t events_pitch, int32_t* d_nevents, int max_events, uint16_t* d_syndromes,
                  int n_streams, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3. The fused receive chain: RF_frontend + mono|stereo + rds for n_streams stations
 *    (include/rffrontend.h:5, mono.h:5, stereo.h:4, rds.h:4; struct args, include/args.h:6-19).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    /* mirrors struct args, include/args.h:6-19 */
    int rf_Fs, rf_Fc, rf_taps, rf_decim;
    int audio_decim, audio_upsample;
    int if_Fs, audio_Fc, audio_Fs, symbol_Fs;
    int rds_on;
    /* which audio thread body runs: 'm' = mono(), 's'/'r' = stereo() (src/project.cpp:111-132) */
    int type;
    /* batch */
    int n_streams;
    int device;         /* CUDA device ordinal */
    int keep_stages;    /* 1: keep every intermediate addressable through sdrb_chain_stage (parity tests) */
} sdrb_config;

/* Fills cfg from the reference's mode table (src/project.cpp:31-44,67-108) and type switch (:111-132). */
int sdrb_config_for_mode(int mode, int type, int n_streams, sdrb_config* cfg);

typedef struct sdrb_chain sdrb_chain;

typedef struct {
    int block_pairs;   /* IQ pairs per block per stream (src/rffrontend.cpp:21) */
    int block_bytes;   /* 2*block_pairs */
    int if_block;      /* IF-rate samples per block (src/mono.cpp:19) */
    int audio_block;   /* audio frames per block */
    int pcm_per_block; /* int16 samples per block per stream: audio_block (mono) or 2*audio_block */
    int rds_block;     /* RDS-rate samples per block (if_block*247/640) */
    int max_bits;      /* capacity of the per-block bit record */
    int max_groups;    /* capacity of the per-block group record */
} sdrb_chain_info;

/* Tuning knob read here: the environment variable SDRB_PLL_MAX_CTAS (1..148) overrides the number of SMs the PLL
 * kernel may occupy (default: 32; results do not depend on it), SDRB_SM_PARTITION=0|1 switches the green-context SM
 * partition of the overlap-mode streams off | on (default: on up to 1024 stereo+RDS stations, see sdrb_chain_sm_partition). */
int sdrb_chain_create(const sdrb_config* cfg, sdrb_chain** out);
int sdrb_chain_destroy(sdrb_chain* c);
int sdrb_chain_get_info(const sdrb_chain* c, sdrb_chain_info* info);

/* One block for every stream, input already in device memory: d_iq [n_streams][iq_pitch] bytes,
 * interleaved I,Q,I,Q (src/rffrontend.cpp:48,58-63).  Asynchronous: work is enqueued on the chain's own
 * CUDA streams; results are valid after sdrb_chain_sync() or when read through the calls below. */
int sdrb_chain_process_device(sdrb_chain* c, const uint8_t* d_iq, size_t iq_pitch);

/* Same, from host memory (pinned memory recommended).  The H2D copy is part of the call and is
 * overlapped with the previous block's kernels. */
int sdrb_chain_process_host(sdrb_chain* c, const uint8_t* h_iq, size_t iq_pitch);

int sdrb_chain_sync(sdrb_chain* c);

/* Overlap mode (default 0).
 * 0: every kernel of a block is issued on the chain's stream, in order.
 * 1: the front end of block b+1 (RF front end, band filters) runs concurrently with the PLL and back end of block b
 *    on internal CUDA streams chained by events (the rings are three slots deep for this), and process_host copies on
 *    a copy-engine stream.  The input buffer of block b is then read asynchronously with respect to the chain's
 *    stream, and issuing further blocks never waits for it on the host.  It may be changed or freed only after one of:
 *    sdrb_chain_sync(); sdrb_chain_join() followed by a synchronise of the chain's stream; a lag-0 read of block b
 *    (sdrb_chain_read_* right after issuing it); a lag-1 read issued after block b+1 (it waits for the back end of b,
 *    which is ordered after everything that read b's input); or sdrb_chain_input_consumed(c, 0) returning 1. */
int sdrb_chain_set_overlap(sdrb_chain* c, int on);

/* 1 if the input buffer handed to the block issued `lag` (0 or 1) calls ago has been consumed completely (the H2D copy
 * for process_host, the RF front end for process_device), 0 if it may still be read, < 0 on error (-sdrb_status).
 * Never blocks. */
int sdrb_chain_input_consumed(sdrb_chain* c, int lag);

/* Makes the chain's stream wait (on the device, without blocking the host) for everything issued so far. */
int sdrb_chain_join(sdrb_chain* c);

/* Issue this chain's work on the caller's CUDA stream (e.g. the application's or torch's current stream)
 * instead of the private stream created by sdrb_chain_create.  Call between blocks only. */
int sdrb_chain_set_stream(sdrb_chain* c, void* stream);

/* Page-locked host memory for process_host / read_* buffers (cudaHostAlloc / cudaFreeHost). */
int sdrb_pinned_alloc(size_t bytes, void** h_ptr);
int sdrb_pinned_free(void* h_ptr);

/* Results of the most recent block.
 * pcm: [n_streams][pcm_pitch] int16, what mono()/stereo() fwrite to stdout (src/mono.cpp:40-45,
 *      src/stereo.cpp:100-111): L on even, R on odd indices for stereo. */
int sdrb_chain_read_pcm(sdrb_chain* c, int16_t* h_pcm, size_t pcm_pitch);
int sdrb_chain_pcm_device(sdrb_chain* c, const int16_t** d_pcm, size_t* pcm_pitch);

/* Per-block RDS record (src/rds.cpp:135-189). */
typedef struct {
    int32_t cdr_offset;  /* -1 while the decoder is gated (block_count <= 5 or !rds_on) */
    int32_t n_symbols;
    int32_t n_bits;
    int32_t n_groups;    /* groups completed by this block's frame sync (0 except every 15th decode block) */
    uint8_t bits[48];    /* differentially decoded bits of this block */
    uint64_t groups[8];  /* 64-bit group registers A|B|C|D as handed to parse() (src/rds_utilities.cpp:172) */
} sdrb_rds_record;
int sdrb_chain_read_rds(sdrb_chain* c, sdrb_rds_record* h_records /* [n_streams] */);

/* PCM and/or RDS records (either pointer may be NULL) of block (most recent - lag), lag = 0 or 1.  Outputs are double
 * buffered by block parity, so with lag = 1 the read of block b-1 overlaps the processing of block b (overlap mode). */
int sdrb_chain_read_results(sdrb_chain* c, int lag, int16_t* h_pcm, size_t pcm_pitch, sdrb_rds_record* h_records);

/* parse() — src/rds_utilities.cpp:172-199: the text the reference prints on stderr for one group.
 * chars/output are the caller-held PS assembly state (src/rds.cpp:68-69).  Returns bytes written. */
int sdrb_rds_parse(uint64_t group, uint64_t* chars, uint64_t* output, char* text, int text_cap);

/* Intermediates of the most recent block (keep_stages = 1 only).  Names: fm_demod pilot carrier stereo_band
 * stereo_dc mono_filt stereo_filt audio_filt rds_band gen_pilot IPLL rds_band_delay rds_dc rds_filt rds_clean I_ds Q_ds.  Copies `count` floats per
 * stream into h_out [n_streams][count]; *count is set to the per-stream length. */
int sdrb_chain_stage(sdrb_chain* c, const char* name, float* h_out, int cap_per_stream, int* count);

/* Carried state (checkpoint/resume): opaque blob, size via sdrb_chain_state_bytes. */
size_t sdrb_chain_state_bytes(const sdrb_chain* c);
int sdrb_chain_state_save(sdrb_chain* c, void* h_blob);
int sdrb_chain_state_load(sdrb_chain* c, const void* h_blob);
/* Byte offset, inside the blob sdrb_chain_state_save writes for the current block index, of a named item: "pll19" /
 * "pll114" ([n_streams] x {float feedbackI, feedbackQ, integrator, phaseEst; double trigOffset}), "iq_halo",
 * "rds_decoder", "rds_filt_state"; -1 if the chain has no such item.  For tools that edit a checkpoint. */
long long sdrb_chain_state_item_offset(const sdrb_chain* c, const char* name);
/* Same with the size of the caller's buffer: a blob shorter than the size recorded in its header (a truncated file)
 * is rejected with SDRB_ERR_INVALID instead of being read past its end.  After a load no results exist until the
 * next block was processed (the read calls return SDRB_ERR_STATE). */
int sdrb_chain_state_load_n(sdrb_chain* c, const void* h_blob, size_t blob_bytes);

/* Mean device time per launch of each kernel family, in milliseconds, over the blocks processed since profiling was
 * switched on (CUDA event pairs recorded on the stream each kernel runs on; at most the last 128 blocks).
 * names/ms arrays of capacity cap; returns the count in *n. */
int sdrb_chain_kernel_times(sdrb_chain* c, const char** names, float* ms, int cap, int* n);
int sdrb_chain_set_profiling(sdrb_chain* c, int on);
/* Debugging aid (compute-sanitizer substitute): a chain created while the environment variable SDRB_GUARD=1 is set puts a
 * 512-byte canary zone before and after every device allocation it makes.  This call verifies all of them: SDRB_OK and
 * the number of allocations checked, or SDRB_ERR_STATE naming the allocation a kernel wrote outside of. */
int sdrb_chain_check_guards(sdrb_chain* c, int* n_checked);
/* Health counter of the PLL kernel: how many times a station's 4-sample chunk left the speculative fast path and was
 * recomputed by the careful path (counts[0] the 19 kHz loop, counts[1] the 114 kHz loop), since the chain was created.
 * Results never depend on it; it is what explains a slow k_pll (normally ~1e-5 per sample). */
int sdrb_chain_pll_redos(sdrb_chain* c, unsigned long long counts[2]);
/* The same two totals followed by per-test counters (counts[2*(1+test)+loop]); the per-test part is only filled by a
 * library built with -DSDRB_PLL_DIAG (tools/pll_drift.py --diag), zero otherwise. */
int sdrb_chain_pll_redo_detail(sdrb_chain* c, unsigned long long counts[20]);
/* Capacity events of the RDS back end since the chain was created: counts[0] blocks whose bit count exceeded
 * sdrb_chain_info.max_bits, counts[1] blocks whose bits did not fit the frame-sync buffer, counts[2] blocks that completed
 * more than max_groups groups.  All three are unreachable with the reference's rates (37 bits per block, 15 blocks per
 * frame-sync call); a non-zero count means data was cut off and says where. */
int sdrb_chain_rds_overflows(sdrb_chain* c, unsigned int counts[3]);
/* SM partition of the overlap-mode streams: sms[0] SMs owned by the PLL stream, sms[1] by the front- and back-end streams
 * (CUDA green contexts, created with the chain while the PLL bounds the step, i.e. up to 1024 stereo+RDS stations; see
 * DESIGN.md 5).  Both 0 for larger batches, when the driver refused, or with the environment variable SDRB_SM_PARTITION=0
 * (1 forces the partition on): the chain then runs on ordinary priority streams, with identical results. */
int sdrb_chain_sm_partition(const sdrb_chain* c, int sms[2]);
/* Number of kernels launched by this chain so far. */
long long sdrb_chain_launch_count(const sdrb_chain* c);

#ifdef __cplusplus
}
#endif
#endif /* SDR_B200_H */
