// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Function-level harness around the UNMODIFIED reference sources.  It is compiled by
// oracle/Makefile together with /root/reference/src/{filter,demod,pll,rds_utilities}.cpp
// (from where they lie; nothing is copied into this repo) plus oracle/padnew.cpp, and the
// binary lands in oracle/_ref/ref_harness.
//
// The reference exposes its DSP as free functions (include/filter.h, demod.h, pll.h,
// rds_utilities.h) but its block loops live inside three never-returning thread bodies
// (src/rffrontend.cpp:45-76, src/mono.cpp:29-49, src/stereo.cpp:69-114, src/rds.cpp:95-192).
// This harness calls the reference functions in the same order, with the same arguments and
// the same carried state as those loops, single-threaded, and dumps every intermediate so the
// CUDA path and the C restatement (oracle/sdr_oracle.c) can be compared stage by stage.
//
// Commands (files are oracle/recfile.h containers):
//   ref_harness taps  out.rec
//   ref_harness chain <mode 0-3> <m|s|r> in.raw out.rec [max_blocks] [stage,stage,...|all|out]
//   ref_harness op    in.rec out.rec
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "filter.h"
#include "demod.h"
#include "pll.h"
#include "rds_utilities.h"
#include "recfile.h"

namespace {

// The mode table of src/project.cpp:31-44,67-108.
struct ModeCfg {
    int rf_Fs, rf_Fc, rf_decim, if_Fs, audio_Fc, audio_Fs, sps;
    unsigned short rf_taps;
    float audio_decim, audio_upsample;
};

ModeCfg mode_cfg(int mode) {
    ModeCfg c{2400000, 100000, 10, 240000, 16000, 48000, 39, 101, 5.0f, 1.0f};
    switch (mode) {
        case 0: c.rf_Fs = 2.4e6; c.rf_decim = 10; c.audio_decim = 5; c.if_Fs = 240e3; break;
        case 1: c.rf_Fs = 1.44e6; c.rf_decim = 4; c.audio_decim = 9; c.if_Fs = 360e3; break;
        case 2: c.rf_Fs = 2.4e6; c.rf_decim = 10; c.audio_decim = 800; c.if_Fs = 240e3;
                c.audio_upsample = 147; c.sps = 20; break;
        case 3: c.rf_Fs = 1.152e6; c.rf_decim = 3; c.audio_decim = 1280; c.if_Fs = 384e3;
                c.audio_upsample = 147; c.sps = 20; break;
        default: throw std::runtime_error("mode must be 0..3");
    }
    return c;
}

template <typename T> void append(std::vector<T>& dst, const std::vector<T>& src) {
    dst.insert(dst.end(), src.begin(), src.end());
}
void append_i(std::vector<int32_t>& dst, const std::vector<int>& src) {
    dst.insert(dst.end(), src.begin(), src.end());
}

// Captures everything the reference prints on std::cerr (src/rds_utilities.cpp:179-197).
struct CerrCapture {
    std::ostringstream os;
    std::streambuf* old;
    CerrCapture() : old(std::cerr.rdbuf(os.rdbuf())) {}
    ~CerrCapture() { std::cerr.rdbuf(old); }
};

// start_frame_sync (src/rds_utilities.cpp:384-400) inlines the only place where a group is
// recognised (check_block -> isSequenceABCD -> parse).  To observe the group registers we run
// the same sliding scan over the reference's own check_block and record `reg` whenever the
// offset window reads A,B,C,D after a hit; `verify` re-runs the real start_frame_sync on a
// copy of the state and checks both end in the same state.
struct FrameSyncState {
    std::vector<int> carry;
    uint64_t reg = 0, chars = 0, output = 0;
    bool first_time = true;
    std::deque<std::string> window;
};

void frame_sync_observed(std::vector<int> stream, FrameSyncState& st, std::vector<uint64_t>& groups,
                         std::vector<int32_t>& offsets_seen) {
    // reference call on a copy (ground truth for the carried state); its stderr output is kept
    FrameSyncState ref = st;
    std::vector<int> ref_stream = stream;
    unsigned int ref_idx = 0;
    start_frame_sync(ref_idx, ref_stream, ref.carry, ref.reg, ref.chars, ref.output, ref.first_time, ref.window);

    // observed scan (text of this second pass is discarded)
    {
        std::ostringstream sink;
        std::streambuf* keep = std::cerr.rdbuf(sink.rdbuf());
        stream.insert(stream.begin(), st.carry.begin(), st.carry.end());
        unsigned int idx = 0;
        unsigned int end_range = stream.size() - 26;
        std::string kind;
        while (idx < end_range) {
            check_block(kind, stream.begin() + idx, stream.begin() + idx + 26, st.reg, st.chars, st.output,
                        st.first_time, st.window);
            if (kind != "None") {
                int code = kind == "A" ? 0 : kind == "B" ? 1 : kind == "C" ? 2 : kind == "Cp" ? 3 : 4;
                offsets_seen.push_back(static_cast<int32_t>(idx));
                offsets_seen.push_back(code);
                if (st.window.size() == 4 && st.window[0] == "A" && st.window[1] == "B" && st.window[2] == "C" &&
                    st.window[3] == "D")
                    groups.push_back(st.reg);
                idx += 26;
            } else {
                idx += 1;
            }
        }
        st.carry.assign(stream.begin() + idx, stream.end());
        std::cerr.rdbuf(keep);
    }
    if (st.carry != ref.carry || st.reg != ref.reg || st.chars != ref.chars || st.output != ref.output ||
        st.window != ref.window || st.first_time != ref.first_time)
        throw std::runtime_error("ref_harness: observed frame-sync scan diverged from start_frame_sync");
}

// ---------------------------------------------------------------------------------------------
int cmd_taps(const std::string& out_path) {
    rec::Writer w(out_path);
    for (int mode = 0; mode < 4; mode++) {
        ModeCfg c = mode_cfg(mode);
        std::string sfx = "_m" + std::to_string(mode);
        std::vector<float> h;
        impulseResponseLPF(c.rf_Fs, c.rf_Fc, c.rf_taps, h);  // src/rffrontend.cpp:24
        w.put("rf_h" + sfx, h);
        int up = c.audio_upsample;
        impulseResponseLPF(c.if_Fs * up, c.audio_Fc, c.rf_taps * up, h, up);  // src/mono.cpp:22
        w.put("audio_h" + sfx, h);
        float fb_pilot[] = {18.5e3, 19.5e3}, fb_carrier[] = {37.5e3, 38.5e3}, fb_stereo[] = {22e3, 54e3};
        impulseResponseBPF(c.rf_Fs / c.rf_decim, fb_pilot, c.rf_taps, h);  // src/stereo.cpp:65
        w.put("pilot_h" + sfx, h);
        impulseResponseBPF(c.rf_Fs / c.rf_decim, fb_carrier, c.rf_taps, h);  // src/stereo.cpp:66 (unused there)
        w.put("carrier_h" + sfx, h);
        impulseResponseBPF(c.rf_Fs / c.rf_decim, fb_stereo, c.rf_taps, h);  // src/stereo.cpp:67
        w.put("stereo_h" + sfx, h);
        float fb_rds[] = {54e3, 60e3}, fb_rds_sq[] = {113.5e3, 114.5e3};
        impulseResponseLPF(c.if_Fs * 247, 3e3, c.rf_taps * 247, h, 247);  // src/rds.cpp:61
        w.put("rds_lpf_h" + sfx, h);
        impulseResponseBPF(c.if_Fs, fb_rds, c.rf_taps, h);  // src/rds.cpp:62
        w.put("rds_h" + sfx, h);
        impulseResponseBPF(c.if_Fs, fb_rds_sq, c.rf_taps, h);  // src/rds.cpp:63
        w.put("rds_pilot_h" + sfx, h);
        impulseResponseRRC(2375 * c.sps, c.rf_taps, h);  // src/rds.cpp:65
        w.put("rrc_h" + sfx, h);
    }
    std::vector<float> h;
    impulseResponseAPF(1, 101, h);  // src/stereo.cpp:63, src/rds.cpp:64
    w.put("apf_h", h);
    return 0;
}

// ---------------------------------------------------------------------------------------------
struct StageSink {
    std::set<std::string> want;
    bool all = false;
    std::map<std::string, std::vector<float>> f;
    std::map<std::string, std::vector<int32_t>> i;
    bool on(const std::string& n) const { return all || want.count(n); }
    void add(const std::string& n, const std::vector<float>& v) { if (on(n)) append(f[n], v); }
    void add_n(const std::string& n, const std::vector<float>& v, size_t count) {
        if (on(n)) f[n].insert(f[n].end(), v.begin(), v.begin() + count);
    }
    void add_i(const std::string& n, const std::vector<int>& v) { if (on(n)) append_i(i[n], v); }
};

int cmd_chain(int mode, char type, const std::string& in_path, const std::string& out_path, long max_blocks,
              const std::string& stages) {
    ModeCfg c = mode_cfg(mode);
    const bool want_stereo = (type == 's' || type == 'r');
    const bool rds_on = (type == 'r');

    StageSink sink;
    {
        std::stringstream ss(stages);
        std::string tok;
        while (std::getline(ss, tok, ',')) {
            if (tok == "all") sink.all = true;
            else if (tok == "out") { /* only final products, always written */ }
            else if (!tok.empty()) sink.want.insert(tok);
        }
    }

    std::ifstream in(in_path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + in_path);

    const int rf_decim = c.rf_decim;
    const int audio_decim = c.audio_decim, audio_upsample = c.audio_upsample;
    const int block_size = (1470 * rf_decim * audio_decim) / audio_upsample;  // src/rffrontend.cpp:21
    const int if_block = (1470 * audio_decim) / audio_upsample;               // src/mono.cpp:19

    // --- RF front-end state (src/rffrontend.cpp:24-43)
    std::vector<float> rf_h;
    impulseResponseLPF(c.rf_Fs, c.rf_Fc, c.rf_taps, rf_h);
    std::vector<uint8_t> iq(2 * block_size);
    std::vector<float> I(block_size), Q(block_size), I_ds, Q_ds, fm_demod;
    std::vector<float> state_I(rf_h.size() - 1, 0.0f), state_Q(rf_h.size() - 1, 0.0f);
    float prev_I = 0, prev_Q = 0;

    // --- audio filters (src/mono.cpp:22-27, src/stereo.cpp:14-67)
    std::vector<float> audio_h, mono_delay_h, pilot_h, stereo_h;
    impulseResponseLPF(c.if_Fs * audio_upsample, c.audio_Fc, c.rf_taps * audio_upsample, audio_h, audio_upsample);
    float fb_pilot[] = {18.5e3, 19.5e3}, fb_stereo[] = {22e3, 54e3};
    impulseResponseAPF(1, c.rf_taps, mono_delay_h);
    impulseResponseBPF(c.rf_Fs / c.rf_decim, fb_pilot, c.rf_taps, pilot_h);
    impulseResponseBPF(c.rf_Fs / c.rf_decim, fb_stereo, c.rf_taps, stereo_h);

    // mono driver state (src/mono.cpp:25-27)
    std::vector<float> audio_filt, state_audio(audio_h.size() - 1);
    // stereo driver state (src/stereo.cpp:16-57)
    std::vector<float> carrier(if_block + 1, 0.0f), extracted_pilot, extracted_pilot_state(c.rf_taps - 1, 0.0f),
        stereo_band, stereo_band_state(c.rf_taps - 1, 0.0f), stereo_dc(if_block, 0.0f), mono_state(c.rf_taps - 1, 0.0f),
        mono_filt, mono_delay, mono_delay_state(c.rf_taps - 1, 0.0f), stereo_filt, stereo_state(c.rf_taps - 1, 0.0f);
    carrier[carrier.size() - 1] = 1.0f;
    pllblock_args pll19{1.0f, 0.0f, 0.0f, 0.0f, 0.0, 1.0f};

    // --- RDS driver state (src/rds.cpp:13-93)
    std::vector<float> rds_h, rds_delay_h, rds_lpf_h, rds_pilot_h, rrc_h;
    float fb_rds[] = {54e3, 60e3}, fb_rds_sq[] = {113.5e3, 114.5e3};
    impulseResponseLPF(c.if_Fs * 247, 3e3, c.rf_taps * 247, rds_lpf_h, 247);
    impulseResponseBPF(c.if_Fs, fb_rds, c.rf_taps, rds_h);
    impulseResponseBPF(c.if_Fs, fb_rds_sq, c.rf_taps, rds_pilot_h);
    impulseResponseAPF(1, c.rf_taps, rds_delay_h);
    impulseResponseRRC(2375 * c.sps, c.rf_taps, rrc_h);
    std::vector<float> rds_band, rds_band_sq(if_block, 0.0f), rds_band_state(c.rf_taps - 1, 0.0f), gen_pilot,
        gen_pilot_state(c.rf_taps - 1, 0.0f), IPLL(if_block + 1, 0.0f), rds_band_delay,
        rds_band_delay_state(c.rf_taps - 1, 0.0f), rds_dc(if_block, 0.0f), rds_filt, rds_filt_state(c.rf_taps - 1, 0.0f),
        rds_clean, rds_clean_state(c.rf_taps - 1, 0.0f);
    IPLL[IPLL.size() - 1] = 1.0f;
    pllblock_args pll114{1.0f, 0.0f, 0.0f, 0.0f, 0.0, 0.0f};
    std::vector<int> symbols, bits, decoded_bits, decoded_stream;
    int rds_block_count = 0, half_symbol = 0, start = 0, last_bit = 0, decoder_cont = 0;
    FrameSyncState fs;

    // --- outputs
    std::vector<int16_t> pcm;
    std::vector<int32_t> cdr_offsets, n_symbols, n_bits, all_bits, group_block, sync_hits;
    std::vector<uint64_t> groups;
    CerrCapture text;

    long blocks = 0;
    while (max_blocks < 0 || blocks < max_blocks) {
        in.read(reinterpret_cast<char*>(iq.data()), 2 * block_size);  // src/rffrontend.cpp:48
        if (in.gcount() != 2 * block_size) break;                     // :50 (the binary exits here)

        // src/rffrontend.cpp:58-63
        for (int n = 0; n < 2 * block_size; n++) {
            float v = float(((unsigned char)iq[n] - 128.0) / 128.0);
            if (n & 1) Q[n >> 1] = v; else I[n >> 1] = v;
        }
        convolveFIR(I_ds, I, rf_h, state_I, rf_decim);            // :67
        convolveFIR(Q_ds, Q, rf_h, state_Q, rf_decim);            // :68
        fmDemodNoArctan(I_ds, Q_ds, prev_I, prev_Q, fm_demod);    // :71
        sink.add("I_ds", I_ds);
        sink.add("Q_ds", Q_ds);
        sink.add("fm_demod", fm_demod);

        if (!want_stereo) {
            // src/mono.cpp:34-42
            convolveFIR(audio_filt, fm_demod, audio_h, state_audio, audio_upsample, audio_decim);
            sink.add("audio_filt", audio_filt);
            for (size_t i = 0; i < audio_filt.size(); i++) pcm.push_back(static_cast<short int>(16384 * audio_filt[i]));
        } else {
            // src/stereo.cpp:74-107
            convolveFIR(extracted_pilot, fm_demod, pilot_h, extracted_pilot_state, 1);
            fmpll(extracted_pilot, 19e3, c.rf_Fs / c.rf_decim, carrier, pll19, 2.0, 0, 0.01);
            convolveFIR(stereo_band, fm_demod, stereo_h, stereo_band_state, 1);
            for (unsigned int i = 0; i < stereo_band.size(); i++) stereo_dc[i] = 2.0 * stereo_band[i] * carrier[i];
            convolveFIR(mono_delay, fm_demod, mono_delay_h, mono_delay_state, 1);
            convolveFIR(mono_filt, mono_delay, audio_h, mono_state, audio_upsample, audio_decim);
            convolveFIR(stereo_filt, stereo_dc, audio_h, stereo_state, audio_upsample, audio_decim);
            sink.add("pilot", extracted_pilot);
            sink.add("carrier", carrier);  // if_block+1 values per block, [0] = previous block's last
            sink.add("stereo_band", stereo_band);
            sink.add("stereo_dc", stereo_dc);
            sink.add("mono_delay", mono_delay);
            sink.add("mono_filt", mono_filt);
            sink.add("stereo_filt", stereo_filt);
            for (size_t i = 0; i < mono_filt.size(); i++) {
                short right = static_cast<short int>(16384 * (mono_filt[i] - stereo_filt[i]));
                short left = static_cast<short int>(16384 * (mono_filt[i] + stereo_filt[i]));
                pcm.push_back(left);   // even output index = left  (src/stereo.cpp:103)
                pcm.push_back(right);  // odd  output index = right
            }
        }

        if (type == 'r' || sink.on("rds_clean") || sink.all) {
            // src/rds.cpp:105-133 (this DSP runs in every mode of the binary; only the decode is gated)
            convolveFIR(rds_band, fm_demod, rds_h, rds_band_state, 1);
            for (int i = 0; i < if_block; i++) rds_band_sq[i] = rds_band[i] * rds_band[i];
            convolveFIR(gen_pilot, rds_band_sq, rds_pilot_h, gen_pilot_state, 1);
            fmpll(gen_pilot, 114e3, c.if_Fs, IPLL, pll114, 0.5, 0, 0.001);
            convolveFIR(rds_band_delay, rds_band, rds_delay_h, rds_band_delay_state, 1);
            for (int i = 0; i < if_block; i++) rds_dc[i] = 2 * rds_band_delay[i] * IPLL[i];
            convolveFIR(rds_filt, rds_dc, rds_lpf_h, rds_filt_state, 247, 640);
            convolveFIR(rds_clean, rds_filt, rrc_h, rds_clean_state, 1);
            sink.add("rds_band", rds_band);
            sink.add("gen_pilot", gen_pilot);
            sink.add("IPLL", IPLL);
            sink.add("rds_band_delay", rds_band_delay);
            sink.add("rds_dc", rds_dc);
            sink.add("rds_filt", rds_filt);
            sink.add("rds_clean", rds_clean);

            if (rds_block_count > 5 && rds_on) {  // src/rds.cpp:135
                int sample_offset = cdr(c.sps, rds_clean);  // :137
                symbols.clear();
                for (int i = 0; sample_offset + i * c.sps < (int)rds_clean.size(); i++)  // :157-161
                    symbols.push_back(rds_clean[sample_offset + i * c.sps] > 0);
                manchester_decode(bits, symbols, rds_block_count, half_symbol, start);  // :164
                differential_decode(decoded_bits, bits, last_bit, rds_block_count);     // :167
                cdr_offsets.push_back(sample_offset);
                n_symbols.push_back(symbols.size());
                n_bits.push_back(decoded_bits.size());
                append_i(all_bits, decoded_bits);
                sink.add_i("symbols", symbols);
                sink.add_i("manchester_bits", bits);
                decoder_cont++;
                decoded_stream.insert(decoded_stream.end(), decoded_bits.begin(), decoded_bits.end());  // :182
                if (decoder_cont == 15) {  // :184-189
                    size_t before = groups.size();
                    frame_sync_observed(decoded_stream, fs, groups, sync_hits);
                    for (size_t g = before; g < groups.size(); g++) group_block.push_back(blocks);
                    decoder_cont = 0;
                    decoded_stream.clear();
                }
            } else if (rds_on) {
                cdr_offsets.push_back(-1);
                n_symbols.push_back(0);
                n_bits.push_back(0);
            }
            rds_block_count++;
        }
        blocks++;
    }

    rec::Writer w(out_path);
    std::vector<int32_t> meta = {mode, (int32_t)type, (int32_t)blocks, block_size, if_block};
    w.put("meta", meta);
    w.put("pcm", pcm);
    for (auto& kv : sink.f) w.put(kv.first, kv.second);
    for (auto& kv : sink.i) w.put(kv.first, kv.second);
    if (rds_on) {
        w.put("cdr_offset", cdr_offsets);
        w.put("n_symbols", n_symbols);
        w.put("n_bits", n_bits);
        w.put("rds_bits", all_bits);
        w.put("groups", groups);
        w.put("group_block", group_block);
        w.put("sync_hits", sync_hits);
        w.put("text", text.os.str());
        std::vector<uint64_t> st = {fs.reg, fs.chars, fs.output, (uint64_t)fs.carry.size()};
        w.put("framesync_state", st);
    }
    // carried state at end of run (for checkpoint parity)
    std::vector<float> pst = {pll19.feedbackI, pll19.feedbackQ, pll19.integrator, pll19.phaseEst, (float)pll19.trigOffset,
                              pll114.feedbackI, pll114.feedbackQ, pll114.integrator, pll114.phaseEst, (float)pll114.trigOffset};
    w.put("pll_state", pst);
    return 0;
}

// ---------------------------------------------------------------------------------------------
float f32_param(const std::map<std::string, rec::Array>& m, const std::string& k, size_t idx = 0) {
    return m.at(k).as<float>()[idx];
}
int i32_param(const std::map<std::string, rec::Array>& m, const std::string& k, size_t idx = 0) {
    return m.at(k).as<int32_t>()[idx];
}

int cmd_op(const std::string& in_path, const std::string& out_path) {
    auto m = rec::read_all(in_path);
    std::string op(m.at("op").as<char>(), m.at("op").bytes.size());
    rec::Writer w(out_path);

    if (op == "design") {
        // kind: 0 LPF, 1 LPF with gain u, 2 BPF, 3 APF, 4 RRC;  p = {Fs, a, b}; n = taps; u = gain
        int kind = i32_param(m, "kind"), n = i32_param(m, "n");
        std::vector<float> h;
        float Fs = f32_param(m, "p", 0), a = f32_param(m, "p", 1), b = f32_param(m, "p", 2);
        if (kind == 0) impulseResponseLPF(Fs, a, (unsigned short)n, h);
        else if (kind == 1) impulseResponseLPF(Fs, a, (unsigned short)n, h, i32_param(m, "u"));
        else if (kind == 2) { float fb[2] = {a, b}; impulseResponseBPF(Fs, fb, (unsigned short)n, h); }
        else if (kind == 3) impulseResponseAPF(a, (unsigned short)n, h);
        else if (kind == 4) impulseResponseRRC(Fs, (unsigned short)n, h);
        else throw std::runtime_error("design: bad kind");
        w.put("h", h);
        return 0;
    }

    const int nblocks = m.count("nblocks") ? i32_param(m, "nblocks") : 1;

    if (op == "fir_decim" || op == "fir_updown") {
        std::vector<float> x = m.at("x").vec<float>(), h = m.at("h").vec<float>();
        const bool updown = (op == "fir_updown");
        int decim = updown ? 0 : i32_param(m, "decim");
        int up = updown ? i32_param(m, "up") : 0, down = updown ? i32_param(m, "down") : 0;
        // every call site starts from rf_taps-1 zeros (src/stereo.cpp:24-47, src/rds.cpp:35-47);
        // callers may override with "state0"
        std::vector<float> state = m.count("state0") ? m.at("state0").vec<float>() : std::vector<float>(100, 0.0f);
        size_t blk = x.size() / nblocks;
        std::vector<float> y_all, y;
        for (int b = 0; b < nblocks; b++) {
            std::vector<float> xb(x.begin() + b * blk, x.begin() + (b + 1) * blk);
            if (updown) convolveFIR(y, xb, h, state, up, down);
            else convolveFIR(y, xb, h, state, decim);
            append(y_all, y);
        }
        w.put("y", y_all);
        return 0;
    }
    if (op == "fmdemod") {
        std::vector<float> I = m.at("I").vec<float>(), Q = m.at("Q").vec<float>();
        size_t blk = I.size() / nblocks;
        float pi = 0, pq = 0;
        std::vector<float> y_all, y;
        for (int b = 0; b < nblocks; b++) {
            std::vector<float> ib(I.begin() + b * blk, I.begin() + (b + 1) * blk);
            std::vector<float> qb(Q.begin() + b * blk, Q.begin() + (b + 1) * blk);
            fmDemodNoArctan(ib, qb, pi, pq, y);
            append(y_all, y);
        }
        w.put("y", y_all);
        std::vector<float> prev = {pi, pq};
        w.put("prev", prev);
        return 0;
    }
    if (op == "pll") {
        std::vector<float> x = m.at("x").vec<float>();
        float freq = f32_param(m, "p", 0), Fs = f32_param(m, "p", 1), scale = f32_param(m, "p", 2),
              adjust = f32_param(m, "p", 3), bw = f32_param(m, "p", 4);
        size_t blk = x.size() / nblocks;
        std::vector<float> out(blk + 1, 0.0f);
        out[blk] = 1.0f;  // src/stereo.cpp:45, src/rds.cpp:38
        pllblock_args st{1.0f, 0.0f, 0.0f, 0.0f, 0.0, 1.0f};
        std::vector<float> y_all, state_trace;
        for (int b = 0; b < nblocks; b++) {
            std::vector<float> xb(x.begin() + b * blk, x.begin() + (b + 1) * blk);
            fmpll(xb, freq, Fs, out, st, scale, adjust, bw);
            append(y_all, out);  // blk+1 per block
            state_trace.push_back(st.feedbackI);
            state_trace.push_back(st.feedbackQ);
            state_trace.push_back(st.integrator);
            state_trace.push_back(st.phaseEst);
        }
        w.put("y", y_all);
        w.put("state", state_trace);
        std::vector<double> off = {st.trigOffset};
        w.put("trig_offset", off);
        return 0;
    }
    if (op == "cdr") {
        std::vector<float> x = m.at("x").vec<float>();
        int sps = i32_param(m, "sps");
        size_t blk = x.size() / nblocks;
        std::vector<int32_t> offs;
        for (int b = 0; b < nblocks; b++) {
            std::vector<float> xb(x.begin() + b * blk, x.begin() + (b + 1) * blk);
            offs.push_back(cdr(sps, xb));
        }
        w.put("offset", offs);
        return 0;
    }
    if (op == "bits") {
        // symbols: concatenated per-block symbol vectors; lens: their lengths; block0: first block_count value
        std::vector<int32_t> sym = m.at("symbols").vec<int32_t>(), lens = m.at("lens").vec<int32_t>();
        int block_count = i32_param(m, "block0");
        int half_symbol = 0, start = 0, last_bit = 0;
        std::vector<int32_t> man_all, dec_all, out_lens;
        size_t pos = 0;
        for (size_t b = 0; b < lens.size(); b++) {
            std::vector<int> s(sym.begin() + pos, sym.begin() + pos + lens[b]);
            pos += lens[b];
            std::vector<int> bits, dec;
            manchester_decode(bits, s, block_count, half_symbol, start);
            differential_decode(dec, bits, last_bit, block_count);
            append_i(man_all, bits);
            append_i(dec_all, dec);
            out_lens.push_back(dec.size());
            block_count++;
        }
        w.put("manchester", man_all);
        w.put("decoded", dec_all);
        w.put("lens", out_lens);
        std::vector<int32_t> st = {half_symbol, start, last_bit};
        w.put("state", st);
        return 0;
    }
    if (op == "framesync") {
        // bits: concatenated chunks handed to start_frame_sync one by one; lens: chunk lengths
        std::vector<int32_t> bitv = m.at("bits").vec<int32_t>(), lens = m.at("lens").vec<int32_t>();
        FrameSyncState fs;
        std::vector<uint64_t> groups;
        std::vector<int32_t> hits, groups_per_call;
        CerrCapture text;
        size_t pos = 0;
        for (size_t b = 0; b < lens.size(); b++) {
            std::vector<int> chunk(bitv.begin() + pos, bitv.begin() + pos + lens[b]);
            pos += lens[b];
            size_t before = groups.size();
            frame_sync_observed(chunk, fs, groups, hits);
            groups_per_call.push_back(groups.size() - before);
        }
        w.put("groups", groups);
        w.put("groups_per_call", groups_per_call);
        w.put("sync_hits", hits);
        w.put("text", text.os.str());
        std::vector<uint64_t> st = {fs.reg, fs.chars, fs.output, (uint64_t)fs.carry.size()};
        w.put("state", st);
        std::vector<int32_t> carry(fs.carry.begin(), fs.carry.end());
        w.put("carry", carry);
        return 0;
    }
    if (op == "errdet") {
        // the reference's sync-state-machine decoder (declared in include/rds_utilities.h:14, defined in
        // src/rds_utilities.cpp:202-311, never called by the reference itself): bits handed over chunk by chunk
        std::vector<int32_t> bitv = m.at("bits").vec<int32_t>(), lens = m.at("lens").vec<int32_t>();
        uint64_t reg = 0, chars = 0, output = 0;
        bool first_time = true;
        int sync = 0, prevsync = 0, lastseen_offset = 0, rds_bit_cont = 0, lastseen_offset_cont = 0, block_distance = 0,
            block_number = 0, block_bit_cont = 0, block_cont = 0, wrong_blocks_cont = 0, group_assembly_started = 0,
            group_good_blocks_conts = 0;  // src/rds.cpp:67-84
        CerrCapture text;
        size_t pos = 0;
        for (size_t b = 0; b < lens.size(); b++) {
            std::vector<int> chunk(bitv.begin() + pos, bitv.begin() + pos + lens[b]);
            pos += lens[b];
            error_detection(reg, chars, output, first_time, sync, prevsync, lastseen_offset, rds_bit_cont, lastseen_offset_cont,
                            block_distance, block_number, block_bit_cont, block_cont, wrong_blocks_cont, group_assembly_started,
                            group_good_blocks_conts, chunk);
        }
        w.put("text", text.os.str());
        std::vector<uint64_t> st = {reg, chars, output};
        w.put("state64", st);
        std::vector<int32_t> sti = {sync, prevsync, lastseen_offset, rds_bit_cont, lastseen_offset_cont, block_distance, block_number,
                                    block_bit_cont, block_cont, wrong_blocks_cont, group_assembly_started, group_good_blocks_conts};
        w.put("state", sti);
        return 0;
    }
    throw std::runtime_error("unknown op " + op);
}

}  // namespace

int main(int argc, char** argv) {
    try {
        std::string cmd = argc > 1 ? argv[1] : "";
        if (cmd == "taps" && argc == 3) return cmd_taps(argv[2]);
        if (cmd == "chain" && argc >= 6)
            return cmd_chain(std::atoi(argv[2]), argv[3][0], argv[4], argv[5], argc > 6 ? std::atol(argv[6]) : -1,
                             argc > 7 ? argv[7] : "out");
        if (cmd == "op" && argc == 4) return cmd_op(argv[2], argv[3]);
        std::cerr << "usage: ref_harness taps out.rec | chain <mode> <m|s|r> in.raw out.rec [max_blocks] [stages] | "
                     "op in.rec out.rec\n";
        return 2;
    } catch (const std::exception& e) {
        std::cerr << "ref_harness: " << e.what() << "\n";
        return 1;
    }
}
