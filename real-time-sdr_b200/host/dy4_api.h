// dy4_api.h — the reference's C++ free-function API, implemented on libsdr_b200 (the C ABI of include/sdr_b200.h).
//
// Every declaration below has the name, argument order, argument meaning and ownership conventions of the
// reference function it replaces (cited per function; paths relative to the reference repository):
// outputs are caller-owned vectors that the callee clears and resizes, carried state is passed by non-const
// reference and replaced by the callee, fmpll's output doubles as carried state (element N of the previous call
// becomes element 0).  Code written against the reference's include/{filter,demod,pll,rds_utilities}.h compiles
// unchanged against host/compat/ (one-line forwarding headers of the same names) and links with -ldy4_b200.
//
// Differences, all deliberate:
//   * the DSP runs on the GPU (batch = 1 through the sdrb_* stage primitives); wrong sizes that are undefined
//     behaviour in the reference (e.g. x shorter than the carried state, src/filter.cpp:119,145) throw
//     std::invalid_argument instead; GPU failures throw std::runtime_error with sdrb_last_error();
//   * nothing here terminates the process.
// The thread bodies RF_frontend/mono/stereo/rds (infinite loops around a queue in the reference) are implemented in
// dy4_threads.cpp on dy4::ReceiveChain below with their reference signatures (declared in compat/{rffrontend,mono,
// stereo,rds}.h next to compat/{args,threadsafequeue}.h), so the reference's own src/project.cpp builds unchanged
// against compat/ (host/project_dropin).  ReceiveChain itself is the batched form: one call per block for n_streams
// stations; sdr_project is the single-process command line on top of it.
#pragma once

#include <cstdint>
#include <deque>
#include <string>
#include <vector>

// include/pll.h:10-17
struct pllblock_args {
    float feedbackI;
    float feedbackQ;
    float integrator;
    float phaseEst;
    double trigOffset;
    float lastCarrier;
};

// include/filter.h:18-24, src/filter.cpp:13-147
void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float>& h);
void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float>& h, int u);
void impulseResponseBPF(float Fs, float* Fb, unsigned short int num_taps, std::vector<float>& h);
void impulseResponseAPF(float gain, unsigned short int num_taps, std::vector<float>& h);
void impulseResponseRRC(float Fs, unsigned short int num_taps, std::vector<float>& h);
void convolveFIR(std::vector<float>& y, const std::vector<float>& x, const std::vector<float>& h, std::vector<float>& state, int decim);
void convolveFIR(std::vector<float>& y, const std::vector<float>& x, const std::vector<float>& h, std::vector<float>& state, int up, int down);

// include/demod.h:5-6, src/demod.cpp:3-24
void fmDemodNoArctan(const std::vector<float>& I, const std::vector<float>& Q, float& prev_I, float& prev_Q, std::vector<float>& fm_demod);

// include/pll.h:20, src/pll.cpp:4-61
void fmpll(const std::vector<float>& pllIn, float freq, float Fs, std::vector<float>& pllOut, pllblock_args& block,
           float ncoScale = 1.0, float phaseAdjust = 0.0, float normBandwidth = 0.01);

// include/rds_utilities.h:6-19, src/rds_utilities.cpp
int cdr(int sps, const std::vector<float>& signal);                                                       // :4-21
void manchester_decode(std::vector<int>& bits, const std::vector<int>& symbols, int& block_count, int& half_symbol, int& start);  // :34-68
void differential_decode(std::vector<int>& decoded, const std::vector<int>& bits, int& last_bit, int& block_num);                // :70-88
void check_block(std::string& offset_type, std::vector<int>::iterator bitstream_start, std::vector<int>::iterator bitstream_end,
                 uint64_t& reg, uint64_t& chars, uint64_t& output, bool& first_time, std::deque<std::string>& window);          // :352-381
void start_frame_sync(unsigned int& idx, std::vector<int>& stream, std::vector<int>& synch_state_bits, uint64_t& reg,
                      uint64_t& chars, uint64_t& output, bool& first_time, std::deque<std::string>& window);                    // :384-400
void parse(uint64_t bytes, uint64_t& chars, uint64_t& output, bool& first_time);                                                // :172-199
// the sync-state-machine decoder the reference declares and defines but never calls (src/rds.cpp:177-179)                      // :202-311
void error_detection(uint64_t& reg, uint64_t& chars, uint64_t& output, bool& first_time, int& sync, int& prevsync, int& lastseen_offset,
                     int& rds_bit_cont, int& lastseen_offset_cont, int& block_distance, int& block_number, int& block_bit_cont, int& blocks_cont,
                     int& wrong_blocks_cont, int& group_assembly_started, int& group_good_blocks_cont, const std::vector<int>& decoded_bits);

struct sdrb_chain;

namespace dy4 {

// Batched replacement of the three thread bodies (src/rffrontend.cpp, mono.cpp|stereo.cpp, rds.cpp): mode and type
// as on the reference's command line (src/project.cpp:67-132), n_streams independent stations on one GPU.
// The numeric members of the reference's `struct args` (include/args.h:6-19), in its order.
struct ChainParams {
    int rf_Fs, rf_Fc, rf_taps, rf_decim;
    int audio_decim, audio_upsample;
    int if_Fs, audio_Fc, audio_Fs, symbol_Fs;
    bool rds_on;
};

class ReceiveChain {
public:
    ReceiveChain(int mode, char type, int n_streams = 1, int device = 0);
    // From the parameter block the reference's main() fills (src/project.cpp:31-108); type 'm' = mono() is the audio
    // body, 's' = stereo() (rds_on decides whether the RDS decoder prints, src/project.cpp:111-132).
    ReceiveChain(const ChainParams& p, char type, int n_streams = 1, int device = 0);
    ~ReceiveChain();
    ReceiveChain(const ReceiveChain&) = delete;
    ReceiveChain& operator=(const ReceiveChain&) = delete;

    int block_bytes() const;      // bytes of interleaved uint8 IQ per station per block (src/rffrontend.cpp:21,48)
    int pcm_per_block() const;    // int16 samples per station per block (L,R interleaved for stereo)
    int if_block() const;         // demodulated FM samples per station per block (what RF_frontend pushes, src/rffrontend.cpp:55)
    int n_streams() const { return n_; }

    // iq: n_streams rows of block_bytes() bytes, row pitch in bytes.  Results of this block are then available below.
    void process(const uint8_t* iq, size_t pitch);
    // what mono()/stereo() write to stdout for this block (src/mono.cpp:40-45, src/stereo.cpp:100-111)
    void read_pcm(int16_t* pcm, size_t pitch_samples);
    // what rds() writes to stderr for this block, per station (src/rds_utilities.cpp:179-197); empty most of the time
    const std::vector<std::string>& rds_text();
    // the block's demodulated FM signal, n_streams rows of if_block() floats (the payload of the reference's queue)
    void read_fm_demod(float* out, size_t pitch_samples);

private:
    sdrb_chain* c_ = nullptr;
    int n_ = 0;
    bool rds_ = false;
    std::vector<uint64_t> chars_, output_;
    std::vector<std::string> text_;
};

}  // namespace dy4
