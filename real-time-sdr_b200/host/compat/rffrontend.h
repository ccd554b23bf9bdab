// Same name as the reference header include/rffrontend.h: void RF_frontend(args*) (include/rffrontend.h:5, body src/rffrontend.cpp:9-77), implemented on the B200
// chain in host/dy4_threads.cpp.
#pragma once
#include <iostream>
#include <vector>

#include "args.h"

void RF_frontend(args*);
