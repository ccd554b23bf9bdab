// Same name as the reference header include/utilities.h:1-11: the band-type enum and shutdown(), which prints the usage
// text on stderr and ends the process with status 1 (called by main() on a bad mode or type, src/project.cpp:104,129).
#pragma once
#include <cstdlib>
#include <iostream>

typedef enum { M = 0, S = 4, R = 8 } bandtype;

inline void shutdown() {
    std::cerr << "Sorry, the parameters you provided were not valid.\n\tValid modes are:\n\t\t0 - mode 0\n\t\t1 - mode 1\n"
                 "\t\t2 - mode 2\n\t\t3 - mode 3\n\tValid types are:\n\t\tm - mono\n\t\ts - stereo\n\t\tr - rds\n";
    std::exit(1);
}
