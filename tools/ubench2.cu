// ubench2.cu -- which integer instructions overlap on sm_100a (developer tool).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 2048
#define CH 8

template <int OP>
__global__ void __launch_bounds__(256) k(const u64 *in, u64 *out) {
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u32 a[CH], b[CH]; u64 w[CH]; double d[CH]; float f[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) { u64 x = in[(t + c * 977) & 1023]; a[c] = (u32)x; b[c] = (u32)(x >> 32); w[c] = x; d[c] = (double)(x >> 40); f[c] = (float)(x >> 50);}
    u32 m = (u32)in[t & 1023] | 1;
    double dm = 1.0000001; float fm = 1.0001f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (OP == 0) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m), "r"(b[c])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP == 1) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m), "r"(m)); }
            if (OP == 2) { asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP == 3) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m)); }
            if (OP == 4) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[c]) : "r"(a[c]), "r"(m)); a[c] ^= (u32)w[c]; }
            if (OP == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(m)); }
            if (OP == 7) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(m)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));}
            if (OP == 8) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(dm)); }
            if (OP == 9) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(fm)); }
            if (OP == 10) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(fm)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP == 11) { asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a[c]) : "r"(b[c])); }
            if (OP == 12) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(m), "r"(b[c])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[c]) : "r"(m));}
            if (OP == 13) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(m)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[c]) : "r"(m), "r"(m)); }
            if (OP == 14) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(dm)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(m)); }
            if (OP == 15) { asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(a[c]), "+r"(b[c]) : "r"(m)); }
            if (OP == 16) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[c]), "+r"(b[c]) : "r"(m), "r"((u32)w[c])); }
            if (OP == 17) { asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(a[c]) : "h"((unsigned short)b[c]), "h"((unsigned short)m)); b[c] += a[c]; }
            if (OP == 18) { u32 p; asm volatile("{.reg .pred q; setp.lt.u32 q, %1, %2; selp.u32 %0, %1, %2, q;}" : "=r"(p) : "r"(a[c]), "r"(b[c])); a[c] = p + m; }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= w[c] ^ a[c] ^ ((u64)b[c] << 32) ^ (u64)d[c] ^ (u64)f[c];
    out[t] = acc;
}

template <int OP>
void run(const char *name, double n_instr, const u64 *in, u64 *out, int sms) {
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(in, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<OP><<<blocks, 256>>>(in, out); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warps_per_smsp = blocks * 8.0 / sms / 4;
    double groups = warps_per_smsp * ITERS * CH;  // warp-level groups per SMSP
    // cycles per group per SMSP at 1.965 GHz nominal
    double cyc = ms * 1e-3 * 1.965e9 / groups;
    printf("%-44s %7.3f ms  %6.2f cyc/warp-group/SMSP (%g instr in group)\n", name, ms, cyc, n_instr);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64 *in, *out; cudaMalloc(&in, 1024 * 8); cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    u64 h[1024]; u64 s = 12345;
    for (int i = 0; i < 1024; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; h[i] = s; }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<1>("mad.lo", 1, in, out, sms);
    run<2>("add32", 1, in, out, sms);
    run<0>("mad.lo + add32", 2, in, out, sms);
    run<12>("mad.lo + add32 + xor", 3, in, out, sms);
    run<3>("mad.wide (acc)", 1, in, out, sms);
    run<4>("mul.wide + xor", 2, in, out, sms);
    run<5>("mad.wide + 1 add32", 2, in, out, sms);
    run<6>("mad.wide + 2 add32", 3, in, out, sms);
    run<7>("mad.wide + 2 add32 + xor", 4, in, out, sms);
    run<13>("mad.wide + mad.lo", 2, in, out, sms);
    run<8>("dfma", 1, in, out, sms);
    run<14>("dfma + add32", 2, in, out, sms);
    run<9>("ffma", 1, in, out, sms);
    run<10>("ffma + add32", 2, in, out, sms);
    run<11>("prmt", 1, in, out, sms);
    run<15>("add.cc + addc", 2, in, out, sms);
    run<16>("mad.lo.cc + madc.hi", 2, in, out, sms);
    run<17>("mul.wide.u16 + add", 2, in, out, sms);
    run<18>("setp+selp+add", 3, in, out, sms);
    return 0;
}
