// Same name as the reference header include/rds.h: void rds(args*) (include/rds.h:4, body src/rds.cpp:11-193), implemented on the B200
// chain in host/dy4_threads.cpp.
#pragma once
#include <iostream>
#include <vector>

#include "args.h"

void rds(args*);
