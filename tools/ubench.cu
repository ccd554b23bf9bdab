// Microbenchmarks that size the receive-chain kernels on B200 (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu && build/ubench
// 1. dependent-issue latency of the FP32/FP64 instructions the PLL recurrence is built from
// 2. issue throughput of the bit-exact MAC (FMUL+FADD, no contraction) in scalar and packed f32x2 form
// Results are printed as JSON lines; profiles/ubench_rNN.txt keeps a copy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kIters = 4096;

template <int OP>
__global__ void lat_kernel(double seed, long long* cycles, double* sink) {
    double d = seed, e = seed * 0.5 + 1.0;
    float f = (float)seed, g = 1.0f + (float)seed * 0.25f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < kIters; i++) {
        if (OP == 0) f = __fmaf_rn(f, g, g);
        if (OP == 1) f = __fadd_rn(f, g);
        if (OP == 2) f = __fmul_rn(f, g);
        if (OP == 3) d = __fma_rn(d, e, e);
        if (OP == 4) d = __dadd_rn(d, e);
        if (OP == 5) d = __dmul_rn(d, e);
        if (OP == 6) d = __ddiv_rn(e, d) + 1.0;
        if (OP == 7) { f = (float)d; d = (double)f + e; }            // F2F.F32.F64 + F2F.F64.F32 + DADD
        if (OP == 8) d = rint(d * e);                                 // DMUL + FRND.F64
        if (OP == 9) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(d)); d += e; }  // MUFU.RCP64H + DADD
        if (OP == 10) f = __fdividef(g, f) + g;                       // MUFU.RCP + FMUL + FADD
        if (OP == 11) { long long b = __double_as_longlong(d); b ^= (long long)i; d = __longlong_as_double(b) + e; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = d + f;
}

// throughput: NACC independent accumulators per thread, MODE 0: FMUL+FADD scalar, 1: FFMA scalar,
// 2: mul.f32x2 + add.f32x2, 3: fma.f32x2
template <int MODE>
__global__ void tput_kernel(float seed, long long* cycles, float* sink, unsigned long long nz) {
    constexpr int NACC = 16;
    float a[NACC];
    float h = 1.0f + seed * 1e-6f;
#pragma unroll
    for (int j = 0; j < NACC; j++) a[j] = seed + j;
    float x = seed * 0.5f;
    long long t0 = clock64();
    for (int i = 0; i < 512; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x = __fmul_rn(x, 1.0001f);  // loop-variant operand: keeps the multiplies inside the loop
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < NACC; j++) a[j] = __fadd_rn(a[j], __fmul_rn(h + (float)j, x));
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < NACC; j++) a[j] = __fmaf_rn(h, a[j], x);
            } else if (MODE == 2) {
                // unfused packed MAC: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (even with -fmad=false), so
                // the product is formed as fma(h, x, nz) with nz = (-0, -0) passed at run time: RN(h*x) exactly, not foldable
#pragma unroll
                for (int j = 0; j < NACC; j += 2) {
                    unsigned long long acc, hh, xx, p;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[j]), "f"(a[j + 1]));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(a[(j + 3) % NACC]));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(hh), "l"(xx), "l"(nz));
                    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(acc), "l"(p));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[j]), "=f"(a[j + 1]) : "l"(acc));
                }
            } else {
#pragma unroll
                for (int j = 0; j < NACC; j += 2) {
                    unsigned long long acc, hh, xx;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[j]), "f"(a[j + 1]));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(hh) : "f"(h));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc) : "l"(hh), "l"(acc), "l"(xx));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a[j]), "=f"(a[j + 1]) : "l"(acc));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; j++) s += a[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// FP64 issue rate of ONE warp with `lanes` active lanes: 8 independent DFMA chains per thread.  (Does a half-empty warp
// pass through the 16-lane FP64 pipe in one cycle instead of two?  It decides whether k_pll should run 16 stations per warp.)
__global__ void dp_tput_kernel(int lanes, double seed, long long* cycles, double* sink) {
    if ((int)threadIdx.x >= lanes) return;
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed + j;
    const double e = seed * 0.5 + 1.0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = __fma_rn(a[j], e, e);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    sink[threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
int run_dp_tput(int lanes) {
    long long* cyc; double* sink;
    CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 8 * 32));
    dp_tput_kernel<<<1, 32>>>(lanes, 1.000001, cyc, sink);
    dp_tput_kernel<<<1, 32>>>(lanes, 1.000001, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"ubench\": \"fp64 issue, one warp\", \"active_lanes\": %d, \"cycles_per_dfma\": %.2f}\n", lanes, (double)h / (1024 * 8));
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

template <int OP>
int run_lat(const char* name, int instr_per_iter) {
    long long* cyc; double* sink;
    CK(cudaMalloc(&cyc, 8 * 148)); CK(cudaMalloc(&sink, 8 * 148 * 32));
    lat_kernel<OP><<<1, 32>>>(1.000001, cyc, sink);
    lat_kernel<OP><<<1, 32>>>(1.000001, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"ubench\": \"latency\", \"op\": \"%s\", \"cycles_per_iter\": %.2f, \"instr_per_iter\": %d}\n", name,
           (double)h / kIters, instr_per_iter);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

template <int MODE>
int run_tput(const char* name, int warps_per_sm) {
    long long* cyc; float* sink;
    int threads = 32 * warps_per_sm;
    CK(cudaMalloc(&cyc, 8 * 148)); CK(cudaMalloc(&sink, 4 * 148 * threads));
    tput_kernel<MODE><<<148, threads>>>(1.5f, cyc, sink, 0x8000000080000000ull);
    tput_kernel<MODE><<<148, threads>>>(1.5f, cyc, sink, 0x8000000080000000ull);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < 148; i++) mean += h[i]; mean /= 148;
    double macs = 512.0 * 4 * 16 * threads;  // per SM
    printf("{\"ubench\": \"throughput\", \"op\": \"%s\", \"warps_per_sm\": %d, \"mac_lanes_per_clk_per_sm\": %.1f}\n", name,
           warps_per_sm, macs / mean);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", p.name, p.multiProcessorCount, p.major, p.minor);
    run_lat<0>("ffma", 1); run_lat<1>("fadd", 1); run_lat<2>("fmul", 1);
    run_lat<3>("dfma", 1); run_lat<4>("dadd", 1); run_lat<5>("dmul", 1);
    run_lat<6>("ddiv_rn+dadd", 2); run_lat<7>("f2f_d2f+f2f_f2d+dadd", 3); run_lat<8>("dmul+drint", 2);
    run_lat<9>("rcp64h+dadd", 2); run_lat<10>("fast_fdiv+fadd", 3); run_lat<11>("lop64+dadd", 2);
    for (int l : {32, 16, 8}) run_dp_tput(l);
    for (int w : {4, 8, 16, 32}) {
        run_tput<0>("fmul+fadd", w); run_tput<1>("ffma", w); run_tput<2>("ffma2(h,x,-0)+fadd2 (unfused packed MAC)", w); run_tput<3>("fma.f32x2", w);
    }
    return 0;
}
