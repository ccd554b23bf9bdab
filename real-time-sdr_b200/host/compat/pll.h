// Same name as the reference header include/pll.h: code written against the reference compiles unchanged.
#pragma once
#include "../dy4_api.h"
