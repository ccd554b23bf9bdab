// Same name as the reference header include/mono.h: void mono(args*) (include/mono.h:5, body src/mono.cpp:8-50), implemented on the B200
// chain in host/dy4_threads.cpp.
#pragma once
#include <iostream>
#include <vector>

#include "args.h"

void mono(args*);
