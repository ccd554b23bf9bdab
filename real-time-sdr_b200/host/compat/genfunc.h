// Same name as the reference header include/genfunc.h so that translation units which merely include it (src/project.cpp:11-15)
// compile unchanged.  The helpers it declares in the reference (Fourier transforms, signal generators, file and log I/O)
// are outside the receive path (out of scope): nothing is declared here, so code that calls them
// fails at compile time instead of silently linking to nothing.
#pragma once
#include <complex>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <vector>
