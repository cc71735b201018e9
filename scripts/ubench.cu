// Instruction-throughput microbenchmarks for sm_100a (B200): what bounds the fp32 pair kernels?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench.cu && /tmp/ubench
// Every pattern is a loop of 8 (or 16) independent instructions per thread pinned by `asm volatile`;
// the result is cycles per warp-instruction per SM sub-partition (SMSP), assuming the SM clock printed.
// Findings are summarised in profiles/r02_ubench.txt and DESIGN.md section 6.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;

#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define MUL2(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define ADD2(d, a, b) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MUL1(d, a, b) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define DFMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define RSQ(d, a) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a))

__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const float *in, float *out, int iters) {
    const int t = threadIdx.x;
    u64 x[8], y[8], a[8];
    float xf[16], yf[16], af[16];
    double xd[8], yd[8], ad[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // disjoint index ranges: no two operands may be provably equal (the compiler would merge them)
        x[i] = pk(in[t + i], in[t + i + 16]); y[i] = pk(in[300 + t + i], in[320 + t + i]); a[i] = pk(in[600 + t + i], in[620 + t + i]);
        xd[i] = in[1800 + t + i]; yd[i] = in[2100 + t + i]; ad[i] = in[2400 + t + i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { xf[i] = in[900 + t + i]; yf[i] = in[1200 + t + i]; af[i] = in[1500 + t + i]; }
    const u64 A = pk(in[3000], in[3001]), B = pk(in[3002], in[3003]);
    const float As = in[3004], Bs = in[3005];
    const double Ad = in[3006], Bd = in[3007];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) FMA2(a[i], a[i], A, B);                    // chain, two loop-invariant operands
                if (MODE == 1) FMA2(a[i], x[i], y[i], a[i]);              // three distinct 64-bit operands
                if (MODE == 2) MUL2(a[i], a[i], y[i]);                    // two distinct operands
                if (MODE == 3) ADD2(a[i], a[i], y[i]);
                if (MODE == 4) { FMA1(af[2 * i], xf[2 * i], yf[2 * i], af[2 * i]); FMA1(af[2 * i + 1], xf[2 * i + 1], yf[2 * i + 1], af[2 * i + 1]); }
                if (MODE == 5) { FMA1(af[2 * i], af[2 * i], As, Bs); FMA1(af[2 * i + 1], af[2 * i + 1], As, Bs); }
                if (MODE == 6) DFMA(ad[i], xd[i], yd[i], ad[i]);
                if (MODE == 7) DFMA(ad[i], ad[i], Ad, Bd);
                if (MODE == 8) RSQ(af[i], af[i]);
                if (MODE == 10) FMA2(a[i], x[i], a[i], a[i]);             // two distinct 64-bit operands (one used twice)
                if (MODE == 11) FMA2(a[i], x[i], A, a[i]);                // one loop-invariant operand
                if (MODE == 12) { MUL1(af[2 * i], af[2 * i], yf[2 * i]); MUL1(af[2 * i + 1], af[2 * i + 1], yf[2 * i + 1]); }
                if (MODE == 13) { FMA2(a[i], x[i], y[i], a[i]); RSQ(af[i], af[i]); }          // FFMA2 + MUFU co-issue
                if (MODE == 15) { FMA2(a[i], x[i], y[i], a[i]); DFMA(ad[i], xd[i], yd[i], ad[i]); }   // FFMA2 + DFMA co-issue
                if (MODE == 16) { FMA2(a[i], a[i], A, B); DFMA(ad[i], ad[i], Ad, Bd); }
                if (MODE == 17) { FMA1(af[2 * i], xf[2 * i], yf[2 * i], af[2 * i]); FMA1(af[2 * i + 1], xf[2 * i + 1], yf[2 * i + 1], af[2 * i + 1]); DFMA(ad[i], xd[i], yd[i], ad[i]); }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(a[i] ^ (a[i] >> 32))) + (float)ad[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += af[i];
    if (s == -12345.678f) out[0] = s;
}

template <int MODE>
void run(const char *name, double instr_per_inner, const float *in, float *out, double mhz) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 256, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(in, out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    // warp-instructions per SMSP: 8 warps per block * 4 blocks per SM / 4 SMSPs = 8 warps per SMSP
    const double winstr = 8.0 * iters * 32.0 * instr_per_inner;
    const double cycles = best * 1e-3 * mhz * 1e6;
    printf("%-58s %8.3f ms  %6.3f cycles per warp-instruction per SMSP\n", name, best, cycles / winstr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    const double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
    float *in, *out;
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMalloc(&out, 16);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 1.0f + 1e-6f * (i % 7);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("assumed SM clock %.0f MHz; 8 warps per SMSP, 8 independent instructions per thread\n", mhz);
    run<0>("FFMA2  a = a*A+B         (2 loop-invariant operands)", 1, in, out, mhz);
    run<1>("FFMA2  a = x*y+a         (3 distinct 64-bit operands)", 1, in, out, mhz);
    run<10>("FFMA2  a = x*a+a         (2 distinct 64-bit operands)", 1, in, out, mhz);
    run<11>("FFMA2  a = x*A+a         (1 loop-invariant operand)", 1, in, out, mhz);
    run<2>("FMUL2  a = a*y", 1, in, out, mhz);
    run<3>("FADD2  a = a+y", 1, in, out, mhz);
    run<4>("FFMA   a = x*y+a         (3 distinct, scalar)", 2, in, out, mhz);
    run<5>("FFMA   a = a*A+B         (scalar chain)", 2, in, out, mhz);
    run<12>("FMUL   a = a*y           (scalar)", 2, in, out, mhz);
    run<6>("DFMA   a = x*y+a         (3 distinct)", 1, in, out, mhz);
    run<7>("DFMA   a = a*A+B         (chain)", 1, in, out, mhz);
    run<8>("MUFU.RSQ", 1, in, out, mhz);
    run<13>("FFMA2 (3 distinct) + MUFU.RSQ     (per pair)", 1, in, out, mhz);
    run<15>("FFMA2 (3 distinct) + DFMA (3 distinct) (per pair)", 1, in, out, mhz);
    run<16>("FFMA2 chain + DFMA chain          (per pair)", 1, in, out, mhz);
    run<17>("2 FFMA (3 distinct) + DFMA (3 distinct) (per triple)", 1, in, out, mhz);
    return 0;
}
