// Implementation of dy4_api.h on the C ABI of libsdr_b200 (see the header for the contract).
#include "dy4_api.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "../../include/sdr_b200.h"

namespace {

void ok(int rc) {
    if (rc != SDRB_OK) throw std::runtime_error(std::string("libsdr_b200: ") + sdrb_last_error());
}
void cu(cudaError_t e) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
}

// device buffer holding a copy of a host vector (batch = 1 staging for the single-stream API)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    explicit DevBuf(size_t count) : n(count) { cu(cudaMalloc((void**)&p, sizeof(T) * (count ? count : 1))); }
    DevBuf(const T* src, size_t count) : DevBuf(count) { if (count) cu(cudaMemcpy(p, src, sizeof(T) * count, cudaMemcpyHostToDevice)); }
    ~DevBuf() { cudaFree(p); }
    void to_host(T* dst, size_t count) const { if (count) cu(cudaMemcpy(dst, p, sizeof(T) * count, cudaMemcpyDeviceToHost)); }
};

// RDS block check (src/rds_utilities.cpp:122-135): parity rows and syndromes as 26-/10-bit masks, first received
// bit = bit 0.  Same constants as the device kernel (sdr_kernels.cuh: rds_block_type).
const uint32_t kParityRows[10] = {0x39BE401u, 0x337C802u, 0x1F47404u, 0x0730C08u, 0x0E61810u,
                                  0x257D420u, 0x3344C40u, 0x1F37C80u, 0x3E6F900u, 0x3CDF200u};
const uint32_t kSyndromes[5] = {0x06Fu, 0x0AFu, 0x0E9u, 0x0CFu, 0x069u};
const char* const kOffsetNames[5] = {"A", "B", "C", "Cp", "D"};

}  // namespace

// ---- tap designers: host side of the library --------------------------------------------------------------
void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float>& h) {
    h.clear(); h.resize(num_taps, 0.0f);
    ok(sdrb_design_lpf(Fs, Fc, num_taps, h.data()));
}
void impulseResponseLPF(float Fs, float Fc, unsigned short int num_taps, std::vector<float>& h, int u) {
    h.clear(); h.resize(num_taps, 0.0f);
    ok(sdrb_design_lpf_gain(Fs, Fc, num_taps, u, h.data()));
}
void impulseResponseBPF(float Fs, float* Fb, unsigned short int num_taps, std::vector<float>& h) {
    h.clear(); h.resize(num_taps, 0.0f);
    ok(sdrb_design_bpf(Fs, Fb[0], Fb[1], num_taps, h.data()));
}
void impulseResponseAPF(float gain, unsigned short int num_taps, std::vector<float>& h) {
    h.clear(); h.resize(num_taps, 0.0f);
    ok(sdrb_design_apf(gain, num_taps, h.data()));
}
void impulseResponseRRC(float Fs, unsigned short int num_taps, std::vector<float>& h) {
    h.clear(); h.resize(num_taps, 0.0f);
    ok(sdrb_design_rrc(Fs, num_taps, h.data()));
}

// ---- block FIRs ------------------------------------------------------------------------------------------
void convolveFIR(std::vector<float>& y, const std::vector<float>& x, const std::vector<float>& h, std::vector<float>& state, int decim) {
    if (h.empty() || decim < 1) throw std::invalid_argument("convolveFIR: empty taps or decim < 1");
    if (state.size() != h.size() - 1) throw std::invalid_argument("convolveFIR: state must hold h.size()-1 samples");
    if (x.size() < state.size()) throw std::invalid_argument("convolveFIR: x shorter than the carried state (undefined in the reference)");
    y.clear(); y.resize(x.size() / decim, 0.0f);
    DevBuf<float> dx(x.data(), x.size()), ds(state.data(), state.size()), dy(y.size());
    ok(sdrb_fir_decim(dx.p, x.size(), (int)x.size(), h.data(), (int)h.size(), ds.p, dy.p, y.size() ? y.size() : 1, decim, 1, nullptr));
    cu(cudaDeviceSynchronize());
    dy.to_host(y.data(), y.size());
    ds.to_host(state.data(), state.size());
}

void convolveFIR(std::vector<float>& y, const std::vector<float>& x, const std::vector<float>& h, std::vector<float>& state, int up, int down) {
    if (h.empty() || up < 1 || down < 1) throw std::invalid_argument("convolveFIR: bad resampling ratio");
    // The reference sizes this state like the decimator's (h.size()-1) although only the last (h.size()-1)/up entries are
    // ever read (src/filter.cpp:135) and copies "the last h.size()-1 samples of x" even when x is shorter (:145, an
    // out-of-bounds read).  Here: the live part is carried exactly, the never-read part is zero.
    const size_t live = (h.size() - 1) / (size_t)up;
    if (state.size() < live) throw std::invalid_argument("convolveFIR: state shorter than (h.size()-1)/up");
    if (x.size() < live) throw std::invalid_argument("convolveFIR: x shorter than the live state");
    const size_t ny = (x.size() * (size_t)up) / (size_t)down;
    y.clear(); y.resize(ny, 0.0f);
    std::vector<float> tail(state.end() - live, state.end());
    DevBuf<float> dx(x.data(), x.size()), ds(tail.data(), live), dy(ny);
    ok(sdrb_fir_updown(dx.p, x.size(), (int)x.size(), h.data(), (int)h.size(), ds.p, (int)live, dy.p, ny ? ny : 1, up, down, 1, nullptr));
    cu(cudaDeviceSynchronize());
    dy.to_host(y.data(), ny);
    ds.to_host(tail.data(), live);
    std::fill(state.begin(), state.end(), 0.0f);
    std::copy(tail.begin(), tail.end(), state.end() - live);
    if (x.size() >= state.size()) std::copy(x.end() - state.size(), x.end(), state.begin());  // what the reference holds when in bounds
}

// ---- discriminator ------------------------------------------------------------------------------------------
void fmDemodNoArctan(const std::vector<float>& I, const std::vector<float>& Q, float& prev_I, float& prev_Q, std::vector<float>& fm_demod) {
    if (I.size() != Q.size() || I.empty()) throw std::invalid_argument("fmDemodNoArctan: I and Q must be non-empty and of equal size");
    fm_demod.clear(); fm_demod.resize(I.size());
    float prev[2] = {prev_I, prev_Q};
    DevBuf<float> di(I.data(), I.size()), dq(Q.data(), Q.size()), dp(prev, 2), dout(I.size());
    ok(sdrb_fm_demod(di.p, dq.p, I.size(), (int)I.size(), dp.p, dout.p, I.size(), 1, nullptr));
    cu(cudaDeviceSynchronize());
    dout.to_host(fm_demod.data(), I.size());
    dp.to_host(prev, 2);
    prev_I = prev[0];
    prev_Q = prev[1];
}

// ---- PLL ----------------------------------------------------------------------------------------------------
void fmpll(const std::vector<float>& pllIn, float freq, float Fs, std::vector<float>& pllOut, pllblock_args& block, float ncoScale,
           float phaseAdjust, float normBandwidth) {
    const size_t n = pllIn.size();
    if (pllOut.size() < n + 1) throw std::invalid_argument("fmpll: pllOut must be pre-sized to pllIn.size()+1 (it carries the last NCO sample)");
    sdrb_pll_state st{block.feedbackI, block.feedbackQ, block.integrator, block.phaseEst, block.trigOffset, block.lastCarrier,
                      pllOut[pllOut.size() - 1]};  // src/pll.cpp:18
    DevBuf<float> din(pllIn.data(), n), dout(n + 1);
    DevBuf<sdrb_pll_state> dst(&st, 1);
    ok(sdrb_pll(din.p, n ? n : 1, (int)n, freq, Fs, ncoScale, phaseAdjust, normBandwidth, dst.p, dout.p, n + 1, 1, nullptr));
    cu(cudaDeviceSynchronize());
    dout.to_host(pllOut.data(), n + 1);
    dst.to_host(&st, 1);
    block.feedbackI = st.feedbackI; block.feedbackQ = st.feedbackQ; block.integrator = st.integrator;
    block.phaseEst = st.phaseEst; block.trigOffset = st.trigOffset; block.lastCarrier = st.lastCarrier;
}

// ---- RDS symbol and bit utilities ---------------------------------------------------------------------------
int cdr(int sps, const std::vector<float>& signal) {
    if (sps < 1) throw std::invalid_argument("cdr: sps < 1");
    DevBuf<float> dx(signal.data(), signal.size());
    DevBuf<int> doff(1);
    ok(sdrb_cdr(dx.p, signal.size() ? signal.size() : 1, (int)signal.size(), sps, doff.p, 1, nullptr));
    cu(cudaDeviceSynchronize());
    int off = 0;
    doff.to_host(&off, 1);
    return off;
}

// The bit-level stages run on the GPU as well: batch = 1 calls of sdrb_manchester_decode, sdrb_differential_decode and
// sdrb_frame_sync (one warp per stream; the fused chain does the same work inside k_rds_backend).
void manchester_decode(std::vector<int>& bits, const std::vector<int>& symbols, int& block_count, int& half_symbol, int& start) {
    const size_t n = symbols.size();
    const int32_t nsym = (int32_t)n;
    sdrb_manchester_state st{half_symbol, start};
    DevBuf<int32_t> dsym(symbols.data(), n), dn(&nsym, 1), dbits(n / 2 + 2), dnb(1);
    DevBuf<sdrb_manchester_state> dst(&st, 1);
    ok(sdrb_manchester_decode(dsym.p, n ? n : 1, dn.p, block_count, dst.p, dbits.p, n / 2 + 2, dnb.p, 1, nullptr));
    cu(cudaDeviceSynchronize());
    int32_t nb = 0;
    dnb.to_host(&nb, 1);
    bits.assign((size_t)nb, 0);
    dbits.to_host(bits.data(), (size_t)nb);
    dst.to_host(&st, 1);
    half_symbol = st.half_symbol;
    start = st.start;
}

void differential_decode(std::vector<int>& decoded, const std::vector<int>& bits, int& last_bit, int& block_num) {
    if (bits.empty()) throw std::invalid_argument("differential_decode: empty input (undefined in the reference)");
    const int32_t nb = (int32_t)bits.size();
    int32_t last = last_bit;
    DevBuf<int32_t> dbits(bits.data(), bits.size()), dn(&nb, 1), dlast(&last, 1), ddec(bits.size());
    ok(sdrb_differential_decode(dbits.p, bits.size(), dn.p, block_num, dlast.p, ddec.p, bits.size(), 1, nullptr));
    cu(cudaDeviceSynchronize());
    decoded.assign(bits.size(), 0);
    ddec.to_host(decoded.data(), bits.size());
    dlast.to_host(&last, 1);
    last_bit = last;
}

void parse(uint64_t bytes, uint64_t& chars, uint64_t& output, bool& first_time) {
    char text[256];
    sdrb_rds_parse(bytes, &chars, &output, text, sizeof text);
    std::cerr << text;
    std::cerr << std::hex;  // the reference's parse() leaves std::hex set on std::cerr (src/rds_utilities.cpp:180): numbers printed later are hex
    (void)first_time;  // the reference writes it after the first group and never reads it again
}

// include/rds_utilities.h:14, src/rds_utilities.cpp:202-311.  The state machine runs on the GPU (sdrb_rds_sync, one warp
// per stream); what the reference prints on std::cerr is replayed here from the events and per-bit syndromes it returns.
void error_detection(uint64_t& reg, uint64_t& chars, uint64_t& output, bool& first_time, int& sync, int& prevsync, int& lastseen_offset,
                     int& rds_bit_cont, int& lastseen_offset_cont, int& block_distance, int& block_number, int& block_bit_cont, int& blocks_cont,
                     int& wrong_blocks_cont, int& group_assembly_started, int& group_good_blocks_cont, const std::vector<int>& decoded_bits) {
    size_t done = 0;
    while (done < decoded_bits.size()) {  // the device entry point takes at most 8160 bits per call
        const size_t n = std::min<size_t>(decoded_bits.size() - done, 8160);
        sdrb_rds_sync_state st{};
        st.reg = reg; st.sync = sync; st.prevsync = prevsync; st.lastseen_offset = lastseen_offset; st.rds_bit_cont = rds_bit_cont;
        st.lastseen_offset_cont = lastseen_offset_cont; st.block_distance = block_distance; st.block_number = block_number;
        st.block_bit_cont = block_bit_cont; st.blocks_cont = blocks_cont; st.wrong_blocks_cont = wrong_blocks_cont;
        st.group_assembly_started = group_assembly_started; st.group_good_blocks_cont = group_good_blocks_cont;
        const std::vector<int32_t> chunk(decoded_bits.begin() + done, decoded_bits.begin() + done + n);
        const int32_t nb = (int32_t)n;
        const int max_events = (int)n / 26 + 8;
        DevBuf<int32_t> dbits(chunk.data(), n), dn(&nb, 1), dnev(1);
        DevBuf<sdrb_rds_sync_state> dst(&st, 1);
        DevBuf<sdrb_rds_sync_event> dev((size_t)max_events);
        DevBuf<uint16_t> dsyn(n);
        ok(sdrb_rds_sync(dbits.p, n, dn.p, nb, dst.p, dev.p, (size_t)max_events, dnev.p, max_events, dsyn.p, 1, nullptr));
        cu(cudaDeviceSynchronize());
        int32_t nev = 0;
        dnev.to_host(&nev, 1);
        std::vector<sdrb_rds_sync_event> ev((size_t)std::min(nev, max_events));
        dev.to_host(ev.data(), ev.size());
        std::vector<uint16_t> syn(n);
        dsyn.to_host(syn.data(), n);
        // replay of the reference's output, bit by bit: the debug line while searching, then whatever happened at that bit
        bool in_sync = sync != 0;
        uint64_t r = reg;
        size_t e = 0;
        for (size_t i = 0; i < n; i++) {
            r = (r << 1) | (uint64_t)chunk[i];
            if (!in_sync) std::cerr << "Reg Syndrome: " << (uint64_t)syn[i] << "    Reg: " << r << std::endl;
            for (; e < ev.size() && ev[e].bit == rds_bit_cont + (int)i; e++) {
                switch (ev[e].type) {
                    case 1: std::cerr << "Sync State Detected" << std::endl; in_sync = true; break;
                    case 2: std::cerr << "Lost Sync (Got " << ev[e].a << " bad blocks on " << ev[e].b << " total)" << std::endl; in_sync = false; break;
                    case 3: std::cerr << "Still Sync-ed (Got " << ev[e].a << " bad blocks on " << ev[e].b << " total)" << std::endl; break;
                    case 4: parse(ev[e].value, chars, output, first_time); break;
                    default: break;  // type 5 (complete groups) is this library's extension: the reference prints nothing for it
                }
            }
        }
        dst.to_host(&st, 1);
        reg = st.reg; sync = st.sync; prevsync = st.prevsync; lastseen_offset = st.lastseen_offset; rds_bit_cont = st.rds_bit_cont;
        lastseen_offset_cont = st.lastseen_offset_cont; block_distance = st.block_distance; block_number = st.block_number;
        block_bit_cont = st.block_bit_cont; blocks_cont = st.blocks_cont; wrong_blocks_cont = st.wrong_blocks_cont;
        group_assembly_started = st.group_assembly_started; group_good_blocks_cont = st.group_good_blocks_cont;
        done += n;
    }
}

void check_block(std::string& offset_type, std::vector<int>::iterator b, std::vector<int>::iterator e, uint64_t& reg, uint64_t& chars,
                 uint64_t& output, bool& first_time, std::deque<std::string>& window) {
    uint32_t w = 0;
    int n = 0;
    for (auto it = b; it != e && n < 26; ++it, ++n) w |= (uint32_t)(*it != 0) << n;
    uint32_t syn = 0;
    for (int c = 0; c < 10; c++) syn |= (uint32_t)(__builtin_popcount(w & kParityRows[c]) & 1) << c;
    for (int t = 0; t < 5; t++) {
        if (syn != kSyndromes[t]) continue;
        offset_type = kOffsetNames[t];
        if (t != 3) {  // C' matches but carries no copy in the reference (:370)
            const int slot = t == 4 ? 3 : t;
            uint64_t word = 0;
            for (int i = 0; i < 16; i++) word |= (uint64_t)((w >> i) & 1u) << (15 - i);
            const int sh = 48 - 16 * slot;
            reg = (reg & ~((uint64_t)0xFFFF << sh)) | (word << sh);
        }
        window.push_back(offset_type);
        if (window.size() > 4) window.pop_front();
        if (window.size() == 4 && window[0] == "A" && window[1] == "B" && window[2] == "C" && window[3] == "D") {
            parse(reg, chars, output, first_time);
            first_time = false;
        }
        return;
    }
    offset_type = "None";
}

void start_frame_sync(unsigned int& idx, std::vector<int>& stream, std::vector<int>& carry, uint64_t& reg, uint64_t& chars, uint64_t& output,
                      bool& first_time, std::deque<std::string>& window) {
    stream.insert(stream.begin(), carry.begin(), carry.end());  // the reference leaves the joined stream in its argument
    if (stream.size() < 26 || idx >= stream.size() - 26) {      // nothing to scan (the reference's unsigned size-26 wraps for < 26 bits)
        carry.assign(stream.begin() + std::min<size_t>(idx, stream.size()), stream.end());
        return;
    }
    // the scan from `idx` of the joined stream = a scan from 0 of its tail, with no carried bits in the device state
    const std::vector<int32_t> tail(stream.begin() + idx, stream.end());
    if (tail.size() > 8128) throw std::invalid_argument("start_frame_sync: more than 8128 bits in one call");
    sdrb_framesync_state st{};
    st.reg = reg;
    for (size_t i = 0; i < window.size() && i < 4; i++)
        for (int t = 0; t < 5; t++)
            if (window[i] == kOffsetNames[t]) st.window[i] = t;
    st.nwindow = (int32_t)std::min<size_t>(window.size(), 4);
    const int32_t nb = (int32_t)tail.size();
    const int max_groups = nb / 104 + 2;  // a group is 4 x 26 bits
    DevBuf<int32_t> dbits(tail.data(), tail.size()), dn(&nb, 1), dng(1);
    DevBuf<sdrb_framesync_state> dst(&st, 1);
    DevBuf<uint64_t> dgroups((size_t)max_groups);
    ok(sdrb_frame_sync(dbits.p, tail.size(), dn.p, nb, dst.p, dgroups.p, (size_t)max_groups, dng.p, max_groups, 1, nullptr));
    cu(cudaDeviceSynchronize());
    int32_t ng = 0;
    dng.to_host(&ng, 1);
    std::vector<uint64_t> groups((size_t)std::min(ng, max_groups));
    dgroups.to_host(groups.data(), groups.size());
    dst.to_host(&st, 1);
    for (uint64_t g : groups) {  // what check_block does at every A,B,C,D completion (src/rds_utilities.cpp:376-379)
        parse(g, chars, output, first_time);
        first_time = false;
    }
    reg = st.reg;
    window.clear();
    for (int i = 0; i < st.nwindow; i++) window.push_back(kOffsetNames[st.window[i]]);
    idx = (unsigned int)(stream.size() - (size_t)st.ncarry);
    carry.assign(stream.end() - st.ncarry, stream.end());
}

// ---- batched chain ------------------------------------------------------------------------------------------
namespace dy4 {

ReceiveChain::ReceiveChain(int mode, char type, int n_streams, int device) : n_(n_streams), rds_(type == 'r') {
    sdrb_config cfg;
    ok(sdrb_config_for_mode(mode, type, n_streams, &cfg));
    cfg.device = device;
    ok(sdrb_chain_create(&cfg, &c_));
    chars_.assign(n_, 0);
    output_.assign(n_, 0);
    text_.assign(n_, std::string());
}
ReceiveChain::ReceiveChain(const ChainParams& p, char type, int n_streams, int device) : n_(n_streams), rds_(type != 'm' && p.rds_on) {
    if (type != 'm' && type != 's' && type != 'r') throw std::invalid_argument("ReceiveChain: type must be 'm', 's' or 'r'");
    sdrb_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.rf_Fs = p.rf_Fs; cfg.rf_Fc = p.rf_Fc; cfg.rf_taps = p.rf_taps; cfg.rf_decim = p.rf_decim;
    cfg.audio_decim = p.audio_decim; cfg.audio_upsample = p.audio_upsample;
    cfg.if_Fs = p.if_Fs; cfg.audio_Fc = p.audio_Fc; cfg.audio_Fs = p.audio_Fs; cfg.symbol_Fs = p.symbol_Fs;
    cfg.rds_on = rds_ ? 1 : 0;
    cfg.type = type == 'm' ? 'm' : (rds_ ? 'r' : 's');
    cfg.n_streams = n_streams;
    cfg.device = device;
    ok(sdrb_chain_create(&cfg, &c_));
    sdrb_chain_info inf;
    ok(sdrb_chain_get_info(c_, &inf));
    rds_ = rds_ && inf.rds_block > 0;  // modes without an RDS back end (see sdrb_chain_create) produce no records
    chars_.assign(n_, 0);
    output_.assign(n_, 0);
    text_.assign(n_, std::string());
}
ReceiveChain::~ReceiveChain() { sdrb_chain_destroy(c_); }
int ReceiveChain::if_block() const { sdrb_chain_info i; sdrb_chain_get_info(c_, &i); return i.if_block; }
void ReceiveChain::read_fm_demod(float* out, size_t pitch_samples) {
    int count = 0;
    ok(sdrb_chain_stage(c_, "fm_demod", out, (int)pitch_samples, &count));
}
int ReceiveChain::block_bytes() const { sdrb_chain_info i; sdrb_chain_get_info(c_, &i); return i.block_bytes; }
int ReceiveChain::pcm_per_block() const { sdrb_chain_info i; sdrb_chain_get_info(c_, &i); return i.pcm_per_block; }
void ReceiveChain::process(const uint8_t* iq, size_t pitch) { ok(sdrb_chain_process_host(c_, iq, pitch)); }
void ReceiveChain::read_pcm(int16_t* pcm, size_t pitch_samples) { ok(sdrb_chain_read_pcm(c_, pcm, pitch_samples)); }
const std::vector<std::string>& ReceiveChain::rds_text() {
    for (auto& t : text_) t.clear();
    if (!rds_) return text_;
    std::vector<sdrb_rds_record> rec(n_);
    ok(sdrb_chain_read_rds(c_, rec.data()));
    char buf[256];
    for (int s = 0; s < n_; s++)
        for (int g = 0; g < rec[s].n_groups; g++) {
            sdrb_rds_parse(rec[s].groups[g], &chars_[s], &output_[s], buf, sizeof buf);
            text_[s] += buf;
        }
    return text_;
}

}  // namespace dy4
