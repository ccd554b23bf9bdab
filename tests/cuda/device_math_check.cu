// Device-side check of csrc/pllmath.cuh against this host's glibc: cos_lean_f, sincos_f, atan2_f on the GPU vs
// (float)cos((double)x) etc. on the CPU.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3
//   usage: device_math_check   (prints one JSON line; exit 0 iff no mismatch)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../real-time-sdr_b200/csrc/pllmath.cuh"

using namespace sdrb::cr;

__global__ void k_eval(const float* t, int n, float* c_lean, float* s, float* c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    c_lean[i] = cos_lean_f(t[i]);
    sincos_f(t[i], s[i], c[i]);
}

int main() {
    std::vector<float> t;
    for (double v = 1.0e7; v < 1.2e7; v += 1.0) t.push_back((float)v);          // integer-valued floats (NCO phase / 2 above 2^24)
    for (double v = 2.0e7; v < 2.2e7; v += 2.0) t.push_back((float)v);
    uint64_t s = 12345;
    for (int i = 0; i < 2000000; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        t.push_back((float)((s >> 11) * (1.0 / 9007199254740992.0) * 6.0e7));
    }
    int n = (int)t.size();
    float *dt, *dl, *ds, *dc;
    cudaMalloc(&dt, 4 * n); cudaMalloc(&dl, 4 * n); cudaMalloc(&ds, 4 * n); cudaMalloc(&dc, 4 * n);
    cudaMemcpy(dt, t.data(), 4 * n, cudaMemcpyHostToDevice);
    k_eval<<<(n + 255) / 256, 256>>>(dt, n, dl, ds, dc);
    std::vector<float> l(n), sv(n), cv(n);
    if (cudaMemcpy(l.data(), dl, 4 * n, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 2; }
    cudaMemcpy(sv.data(), ds, 4 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(cv.data(), dc, 4 * n, cudaMemcpyDeviceToHost);
    long bad_l = 0, bad_s = 0, bad_c = 0; float first = 0;
    for (int i = 0; i < n; i++) {
        float wc = (float)cos((double)t[i]), ws = (float)sin((double)t[i]);
        if (memcmp(&wc, &l[i], 4)) { if (!bad_l) first = t[i]; bad_l++; }
        if (memcmp(&ws, &sv[i], 4)) bad_s++;
        if (memcmp(&wc, &cv[i], 4)) bad_c++;
    }
    printf("{\"n\": %d, \"cos_lean_mismatch\": %ld, \"sin_mismatch\": %ld, \"cos_mismatch\": %ld, \"first_bad_arg\": %.1f}\n", n, bad_l, bad_s, bad_c, first);
    return (bad_l || bad_s || bad_c) ? 1 : 0;
}
