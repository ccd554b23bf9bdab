// sdr_project <mode 0-3> <m|s|r> — the reference's command line (src/project.cpp:26-147) on the B200 chain:
//   rtl_sdr ... - | sdr_project 0 r | aplay -f S16_LE -c 2 -r 48000
// stdin: interleaved uint8 IQ; stdout: int16 PCM (mono, or L,R interleaved); stderr: RDS text ("PI:", "PTY:",
// "Program Service:").  Ends with status 0 at EOF (the reference calls exit(1), src/rffrontend.cpp:50-52, and may lose
// the last one or two blocks to its consumer threads; every complete block read here is written out).
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <vector>

#include "dy4_api.h"

int main(int argc, char** argv) {
    int mode = argc > 1 ? std::atoi(argv[1]) : 0;
    char type = argc > 2 ? argv[2][0] : 'm';
    if (mode < 0 || mode > 3 || (type != 'm' && type != 's' && type != 'r')) {
        std::fprintf(stderr, "usage: %s <mode 0-3> <m|s|r>\n", argv[0]);
        return 2;
    }
    try {
        dy4::ReceiveChain chain(mode, type, 1, 0);
        std::vector<uint8_t> iq(chain.block_bytes());
        std::vector<int16_t> pcm(chain.pcm_per_block());
        while (std::fread(iq.data(), 1, iq.size(), stdin) == iq.size()) {
            chain.process(iq.data(), iq.size());
            chain.read_pcm(pcm.data(), pcm.size());
            std::fwrite(pcm.data(), sizeof(int16_t), pcm.size(), stdout);
            const std::string& t = chain.rds_text()[0];
            if (!t.empty()) std::cerr << t;
        }
        std::fflush(stdout);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "sdr_project: %s\n", e.what());
        return 1;
    }
    return 0;
}
