// ubench4.cu -- does the FP64 pipe run next to the FMA-heavy pipe (IMAD / IMAD.WIDE) on sm_100a?  (developer tool)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench4 tools/ubench4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 2048
#define CH 8

// bit 0: dfma chain, bit 1: mul.wide chain, bit 2: mad.lo chain, bit 3: add chain, bit 4: I2F (u32 -> f64) chain,
// bit 5: F2I (f64 -> u64) chain, bit 6: mad.wide with a 64-bit addend
template <int OP>
__global__ void __launch_bounds__(256) k(const u64 *in, u64 *out) {
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u32 a[CH], b[CH], e[CH]; u64 w[CH], y[CH]; double d[CH], g[CH]; u32 ci[CH]; u64 fo[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { u64 x = in[(t + c * 977) & 1023]; a[c] = (u32)x; b[c] = (u32)(x >> 32); e[c] = a[c] ^ b[c]; w[c] = x; y[c] = x ^ 77; d[c] = (double)(x >> 40); g[c] = (double)(x >> 30); ci[c] = (u32)x >> 3; fo[c] = 0; }
    u32 m = (u32)in[t & 1023] | 1;
    double dm = 1.0000001;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP & 1) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(dm)); }
            if (OP & 2) { u32 lo = (u32)w[c], hi = (u32)(w[c] >> 32); asm volatile("mul.wide.u32 %0,%1,%2;" : "=l"(w[c]) : "r"(lo), "r"(hi | 1)); }
            if (OP & 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m), "r"(a[c])); }
            if (OP & 8) { asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(b[c])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP & 16) { double r; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(r) : "r"(ci[c])); ci[c] = (u32)__double2hiint(r) + (u32)__double2loint(r); }
            if (OP & 32) { u64 r; asm volatile("cvt.rni.u64.f64 %0, %1;" : "=l"(r) : "d"(g[c])); g[c] = __hiloint2double(0x41f00000 | ((u32)r & 0xfffff), (int)(u32)(r >> 7)); }
            if (OP & 64) { u32 lo = (u32)y[c]; asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(y[c]) : "r"(lo), "r"(m)); }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= w[c] ^ y[c] ^ a[c] ^ ((u64)b[c] << 32) ^ (u64)d[c] ^ e[c] ^ ci[c] ^ (u64)g[c];
    out[t] = acc;
}

template <int OP>
void run(const char *name, const u64 *in, u64 *out, int sms) {
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(in, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<OP><<<blocks, 256>>>(in, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double groups = blocks * 8.0 / sms / 4 * ITERS * CH;
    printf("%-44s %7.3f ms  %6.2f cyc/warp-group/SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / groups);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64 *in, *out; cudaMalloc(&in, 1024 * 8); cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    u64 h[1024]; u64 s = 12345;
    for (int i = 0; i < 1024; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; h[i] = s; }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<1>("dfma", in, out, sms);
    run<2>("mul.wide", in, out, sms);
    run<64>("mad.wide (64-bit addend)", in, out, sms);
    run<4>("mad.lo", in, out, sms);
    run<8>("add + xor", in, out, sms);
    run<16>("i2f.f64.u32 (+2 alu)", in, out, sms);
    run<32>("f2i.u64.f64 (+alu)", in, out, sms);
    run<3>("dfma + mul.wide", in, out, sms);
    run<65>("dfma + mad.wide", in, out, sms);
    run<5>("dfma + mad.lo", in, out, sms);
    run<9>("dfma + add + xor", in, out, sms);
    run<6>("mul.wide + mad.lo", in, out, sms);
    run<10>("mul.wide + add + xor", in, out, sms);
    run<12>("mad.lo + add + xor", in, out, sms);
    run<13>("dfma + mad.lo + add + xor", in, out, sms);
    run<11>("dfma + mul.wide + add + xor", in, out, sms);
    run<15>("dfma + mul.wide + mad.lo + add + xor", in, out, sms);
    run<17>("dfma + i2f", in, out, sms);
    run<33>("dfma + f2i", in, out, sms);
    run<18>("mul.wide + i2f", in, out, sms);
    run<48>("i2f + f2i", in, out, sms);
    return 0;
}
