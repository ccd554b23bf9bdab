// Same name as the reference header include/stereo.h: void stereo(args*) (include/stereo.h:4, body src/stereo.cpp:10-115), implemented on the B200
// chain in host/dy4_threads.cpp.
#pragma once
#include <iostream>
#include <vector>

#include "args.h"

void stereo(args*);
