// Same name and layout as the reference's include/args.h:6-19: the parameter block main() hands to the four thread
// bodies (src/project.cpp:31-44 initialises it positionally, so member order and types are part of the interface).
#pragma once

#include <vector>

#include "threadsafequeue.h"

struct args {
    ThreadSafeQueue<std::vector<float>*>& queue;  // RF front end -> audio / RDS hand-off
    int rf_Fs;
    int rf_Fc;
    unsigned short int rf_taps;
    int rf_decim;
    float audio_decim;
    float audio_upsample;
    int if_Fs;
    int audio_Fc;
    int audio_Fs;
    int symbol_Fs;
    bool rds_on;
};
