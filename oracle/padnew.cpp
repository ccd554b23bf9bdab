// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Front-padded global operator new/delete for the reference build under oracle/_ref.
//
// Why: the unmodified reference over-reads the heap in its rational resampler
// (/root/reference/src/filter.cpp:145, and :119) whenever x.size() < h.size()-1
// (RDS path: 7350 < 24946, /root/reference/src/rds.cpp:130; mode-2 audio: 8000 < 14846,
// /root/reference/src/mono.cpp:34).  The values read there are never used afterwards (only
// the last floor((K-1)/up) = 100 state entries are ever indexed, filter.cpp:135), but the
// read itself can fault.  Padding every allocation in front with zeroed, owned memory makes
// the stray read land in mapped memory without touching a single reference source line.
#include <cstdlib>
#include <new>

static const std::size_t kFrontPad = 128 * 1024;  // > 99 784 B, the largest over-read

static void* padded_alloc(std::size_t n) {
    char* p = static_cast<char*>(std::calloc(1, n + kFrontPad));
    if (!p) throw std::bad_alloc();
    return p + kFrontPad;
}
static void padded_free(void* q) noexcept {
    if (q) std::free(static_cast<char*>(q) - kFrontPad);
}

void* operator new(std::size_t n) { return padded_alloc(n); }
void* operator new[](std::size_t n) { return padded_alloc(n); }
void operator delete(void* p) noexcept { padded_free(p); }
void operator delete[](void* p) noexcept { padded_free(p); }
void operator delete(void* p, std::size_t) noexcept { padded_free(p); }
void operator delete[](void* p, std::size_t) noexcept { padded_free(p); }
