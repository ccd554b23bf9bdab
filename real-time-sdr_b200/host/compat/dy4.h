// Same name as the reference header include/dy4.h (its two project-wide constants, :13,17).
#pragma once
#ifndef PI
#define PI 3.14159265358979323846
#endif
#ifndef NFFT
#define NFFT 512
#endif
