// ubench3.cu -- cost of the 64x64 product / fold variants on sm_100a (developer tool).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../twenty-first_b200/csrc/field.cuh"

#define ITERS 2048
#define CH 8

// 4 x mul.wide (no addend) + 6 adds with carry
__device__ __forceinline__ void mul128_nw(u64 a, u64 b, u32 &r0, u32 &r1, u32 &r2, u32 &r3) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u64 p00,p01,p10,p11; .reg .u32 l1,h1,l2,h2;\n\t"
        "mul.wide.u32 p00,%4,%6;\n\t"
        "mul.wide.u32 p01,%4,%7;\n\t"
        "mul.wide.u32 p10,%5,%6;\n\t"
        "mul.wide.u32 p11,%5,%7;\n\t"
        "mov.b64 {%0,%1},p00;\n\t"
        "mov.b64 {%2,%3},p11;\n\t"
        "mov.b64 {l1,h1},p01;\n\t"
        "mov.b64 {l2,h2},p10;\n\t"
        "add.cc.u32 %1,%1,l1;\n\t"
        "addc.cc.u32 %2,%2,h1;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "add.cc.u32 %1,%1,l2;\n\t"
        "addc.cc.u32 %2,%2,h2;\n\t"
        "addc.u32 %3,%3,0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

template <int OP>
__global__ void __launch_bounds__(256) k(const u64 *in, u64 *out) {
    u64 v[CH];
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int c = 0; c < CH; c++) v[c] = in[(t + c * 977) & 1023];
    u64 m = in[t & 1023] | 1;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            u64 x = v[c];
            u32 r0, r1, r2, r3;
            if (OP == 0) { x = gl_mul(x, m); }
            if (OP == 1) { gl_mul128(x, m, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3); }
            if (OP == 2) { mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3); }
            if (OP == 3) { mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3); }
            if (OP == 4) { gl_mul128w(x, m, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3); }
            if (OP == 5) { gl_mul128(x, m, r0, r1, r2, r3); x = gl_pack(r0 ^ r2, r1 ^ r3); }   // product only
            if (OP == 6) { mul128_nw(x, m, r0, r1, r2, r3); x = gl_pack(r0 ^ r2, r1 ^ r3); }   // product only
            if (OP == 7) { u32 a = (u32)x, b = (u32)(x >> 32); u64 p; asm volatile("mul.wide.u32 %0,%1,%2;" : "=l"(p) : "r"(a), "r"(b)); x = p | 1; }
            if (OP == 8) { u32 a = (u32)x; u64 p = x; asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(p) : "r"(a), "r"((u32)m)); x = p; }
            if (OP == 9) { u32 a = (u32)x, b = (u32)(x >> 32); u32 h; asm volatile("mul.hi.u32 %0,%1,%2;" : "=r"(h) : "r"(a), "r"(b)); x = gl_pack(h, b); }
            if (OP == 10) { x = gl_mul(x, x); x = gl_mul(x, x); x = gl_mul(x, m); x = gl_mul(x, m); }  // x^7-like chain
            if (OP == 11) { x = gl_mul_alu(x, x); x = gl_mul_alu(x, x); x = gl_mul_alu(x, m); x = gl_mul_alu(x, m); }
            if (OP == 12) {
                mul128_nw(x, x, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3);
                mul128_nw(x, x, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3);
                mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3);
                mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128a(r0, r1, r2, r3);
            }
            if (OP == 13) {
                mul128_nw(x, x, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3);
                mul128_nw(x, x, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3);
                mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3);
                mul128_nw(x, m, r0, r1, r2, r3); x = gl_reduce128p(r0, r1, r2, r3);
            }
            v[c] = x;
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= v[c];
    out[t] = acc;
}

template <int OP>
void run(const char *name, double ops, const u64 *in, u64 *out, int sms) {
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(in, out);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<blocks, 256>>>(in, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double groups = (double)blocks * 8.0 / sms / 4 * ITERS * CH;  // warp-level ops per SMSP
    printf("%-46s %8.3f ms  %7.2f cyc per warp-op per SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / groups / ops);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64 *in, *out;
    cudaMalloc(&in, 1024 * 8);
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    u64 h[1024];
    u64 s = 12345;
    for (int i = 0; i < 1024; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = s;
    }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<7>("mul.wide.u32 (no addend)", 1, in, out, sms);
    run<8>("mad.wide.u32 (64-bit addend)", 1, in, out, sms);
    run<9>("mul.hi.u32", 1, in, out, sms);
    run<5>("product: mul + 3 mad.wide (current)", 1, in, out, sms);
    run<6>("product: 4 mul.wide + 6 add", 1, in, out, sms);
    run<0>("gl_mul current (mad.wide + IMAD fold)", 1, in, out, sms);
    run<4>("gl_mul low product as mul.wide", 1, in, out, sms);
    run<1>("gl_mul mad.wide + ALU fold", 1, in, out, sms);
    run<2>("gl_mul 4 mul.wide + IMAD fold", 1, in, out, sms);
    run<3>("gl_mul 4 mul.wide + ALU fold", 1, in, out, sms);
    run<10>("x^7-like chain current", 4, in, out, sms);
    run<11>("x^7-like chain gl_mul_alu", 4, in, out, sms);
    run<12>("x^7-like chain 4 mul.wide + ALU fold", 4, in, out, sms);
    run<13>("x^7-like chain 4 mul.wide + IMAD fold", 4, in, out, sms);
    return 0;
}
