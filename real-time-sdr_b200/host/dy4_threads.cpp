// The reference's four thread bodies with their own signatures, on the B200 chain:
//   void RF_frontend(args*)   include/rffrontend.h:5, src/rffrontend.cpp:9-77
//   void mono(args*)          include/mono.h:5,       src/mono.cpp:8-50
//   void stereo(args*)        include/stereo.h:4,     src/stereo.cpp:10-115
//   void rds(args*)           include/rds.h:4,        src/rds.cpp:11-193
// so that the reference's main() (src/project.cpp:26-147) compiles unchanged against compat/ and links with
// libdy4_b200.so (host/project_dropin; `make project_dropin`).
//
// Division of work.  In the reference the RF thread demodulates a block and hands it over the queue; the audio and RDS
// threads each run their half of the DSP on it.  Here one fused GPU chain (dy4::ReceiveChain -> sdrb_chain_process_host)
// does the DSP of all three bodies for a block in one submission, issued by the RF thread, which owns the chain and the
// block loop.  The queue keeps its role and its payload: RF_frontend pushes the block's demodulated FM signal
// (std::vector<float>*, src/rffrontend.cpp:55,73); mono()/stereo() pop it (consumer 0), take the block's PCM and write
// it to stdout; rds() pops it (consumer 1), takes the block's RDS text and writes it to stderr.  prepare() is called
// where the reference calls it: as soon as a consumer no longer needs the payload.
//
// Which audio body main() started decides the chain type ('m' = mono, 's' = stereo; with args::rds_on the RDS decoder
// prints, src/project.cpp:111-132): mono()/stereo() announce themselves to a per-queue session and RF_frontend waits for
// that announcement before it creates the chain (the reference's RF thread would block in its second push() until both
// consumers run, so no caller can observe the difference).
//
// End of input: the reference calls exit(1) from the RF thread the moment std::cin hits EOF (src/rffrontend.cpp:50-52),
// killing its consumers wherever they are: the last one or two blocks are lost at random.  Here every complete block
// that was read is written out first (the consumers acknowledge each finished block), then the process ends with the
// reference's status 1.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <string>
#include <vector>

#include "compat/args.h"
#include "compat/mono.h"
#include "compat/rds.h"
#include "compat/rffrontend.h"
#include "compat/stereo.h"
#include "dy4_api.h"

namespace {

// What the consumers need of one block besides the queue's payload, keyed by that payload's address.
struct BlockOut {
    std::vector<int16_t> pcm;
    std::string text;
    int takers = 2;  // audio + rds
};

struct Session {
    std::mutex m;
    std::condition_variable cv;
    char audio_type = 0;  // 'm' or 's' once the audio body runs
    std::map<const void*, std::shared_ptr<BlockOut>> blocks;
    long long pushed = 0, done[2] = {0, 0};

    void announce_audio(char t) {
        std::lock_guard<std::mutex> lk(m);
        audio_type = t;
        cv.notify_all();
    }
    char wait_audio() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return audio_type != 0; });
        return audio_type;
    }
    void publish(const void* key, std::shared_ptr<BlockOut> b) {
        std::lock_guard<std::mutex> lk(m);
        blocks[key] = std::move(b);
        pushed++;
    }
    std::shared_ptr<BlockOut> take(const void* key) {
        std::lock_guard<std::mutex> lk(m);
        auto it = blocks.find(key);
        if (it == blocks.end()) return nullptr;
        std::shared_ptr<BlockOut> b = it->second;
        if (--b->takers == 0) blocks.erase(it);
        return b;
    }
    void finished(int who) {
        std::lock_guard<std::mutex> lk(m);
        done[who]++;
        cv.notify_all();
    }
    void drain() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done[0] == pushed && done[1] == pushed; });
    }
};

// one session per queue object (args::queue is a reference: its address identifies the pipeline)
Session& session_of(args* p) {
    static std::mutex m;
    static std::map<const void*, std::unique_ptr<Session>> all;
    std::lock_guard<std::mutex> lk(m);
    std::unique_ptr<Session>& s = all[static_cast<const void*>(&p->queue)];
    if (!s) s.reset(new Session());
    return *s;
}

[[noreturn]] void die(const char* who, const std::exception& e) {
    std::fprintf(stderr, "%s: %s\n", who, e.what());
    std::exit(1);
}

void audio_body(args* p, char type) {
    Session& s = session_of(p);
    s.announce_audio(type);
    std::vector<float>* fm_demod = nullptr;
    while (true) {
        p->queue.wait_and_pop(fm_demod, 0);
        std::shared_ptr<BlockOut> b = s.take(fm_demod);
        p->queue.prepare(0);  // the payload is not needed any more (src/mono.cpp:37, src/stereo.cpp:91)
        if (b && !b->pcm.empty()) std::fwrite(b->pcm.data(), sizeof(int16_t), b->pcm.size(), stdout);  // src/mono.cpp:45, src/stereo.cpp:110
        s.finished(0);
    }
}

}  // namespace

void mono(args* p) { audio_body(p, 'm'); }
void stereo(args* p) { audio_body(p, 's'); }

void rds(args* p) {
    Session& s = session_of(p);
    std::vector<float>* fm_demod = nullptr;
    while (true) {
        p->queue.wait_and_pop(fm_demod, 1);
        std::shared_ptr<BlockOut> b = s.take(fm_demod);
        p->queue.prepare(1);  // src/rds.cpp:108
        if (b && !b->text.empty()) std::cerr << b->text;  // parse(), src/rds_utilities.cpp:179-197
        s.finished(1);
    }
}

void RF_frontend(args* p) {
    Session& s = session_of(p);
    try {
        const char type = s.wait_audio();
        const dy4::ChainParams cp{p->rf_Fs, p->rf_Fc, p->rf_taps, p->rf_decim, (int)p->audio_decim, (int)p->audio_upsample,
                                  p->if_Fs, p->audio_Fc, p->audio_Fs, p->symbol_Fs, p->rds_on};
        dy4::ReceiveChain chain(cp, type, 1, 0);
        std::vector<uint8_t> iq((size_t)chain.block_bytes());
        const int n_if = chain.if_block(), n_pcm = chain.pcm_per_block();
        while (true) {
            std::cin.read(reinterpret_cast<char*>(iq.data()), (std::streamsize)iq.size());  // src/rffrontend.cpp:48
            if (std::cin.eof() || std::cin.gcount() != (std::streamsize)iq.size()) {
                s.drain();
                std::fflush(stdout);
                std::_Exit(1);  // src/rffrontend.cpp:50-52 (exit(1)); the consumer threads are parked in wait_and_pop
            }
            chain.process(iq.data(), iq.size());
            std::vector<float>* fm_demod = new std::vector<float>((size_t)n_if);  // owned by the queue from push() on
            chain.read_fm_demod(fm_demod->data(), (size_t)n_if);
            std::shared_ptr<BlockOut> b(new BlockOut());
            b->pcm.resize((size_t)n_pcm);
            chain.read_pcm(b->pcm.data(), (size_t)n_pcm);
            b->text = chain.rds_text()[0];
            s.publish(fm_demod, b);
            p->queue.push(fm_demod);
        }
    } catch (const std::exception& e) {
        die("RF_frontend", e);
    }
}
