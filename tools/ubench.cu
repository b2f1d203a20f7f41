// ubench.cu -- integer pipe throughput on sm_100a (developer tool; informs the kernel design).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../twenty-first_b200/csrc/field.cuh"

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k(const u64 *in, u64 *out) {
    u64 v[CHAINS];
    u32 t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = in[(t + c * 977) & 1023];
    u64 m = in[t & 1023] | 1;
    u32 m32 = (u32)m;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            u64 x = v[c];
            if (OP == 0) {  // mad.wide.u32
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"((u32)x), "r"(m32));
            } else if (OP == 1) {  // mad.lo.u32 x2 (two 32-bit halves)
                u32 a = (u32)x, b = (u32)(x >> 32);
                asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(a) : "r"(b), "r"(m32));
                asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(m32));
                x = gl_pack(a, b);
            } else if (OP == 2) {  // mad.hi.u32 x2
                u32 a = (u32)x, b = (u32)(x >> 32);
                asm volatile("mad.hi.u32 %0, %0, %2, %1;" : "+r"(a) : "r"(b), "r"(m32));
                asm volatile("mad.hi.u32 %0, %0, %2, %1;" : "+r"(b) : "r"(a), "r"(m32));
                x = gl_pack(a, b);
            } else if (OP == 3) {  // 64-bit add = IADD3 + IADD3.X
                asm volatile("add.u64 %0, %0, %1;" : "+l"(x) : "l"(m));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(x) : "l"(m));
            } else if (OP == 4) {  // lop3 x2
                u32 a = (u32)x, b = (u32)(x >> 32);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(m32));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b) : "r"(a), "r"(m32));
                x = gl_pack(a, b);
            } else if (OP == 5) {  // shf x2
                u32 a = (u32)x, b = (u32)(x >> 32);
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 9;" : "+r"(b) : "r"(a));
                x = gl_pack(a, b);
            } else if (OP == 6) {  // 1 mad.wide + 2 adds (mixed pipes)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"((u32)x), "r"(m32));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(x) : "l"(m));
            } else if (OP == 7) {
                x = gl_mul(x, m);
            } else if (OP == 8) {
                x = gl_add_weak(x, m >> 1);
            } else if (OP == 9) {
                x = gl_sub(x, m >> 1);
            } else if (OP == 10) {
                x = gl_canon(x);
                x ^= m;
            } else if (OP == 11) {
                x = gl_mul_pow2(x, 39);
            } else if (OP == 12) {
                x = gl_mul_pow2(x, 78 - 64 + 0) ;
            } else if (OP == 13) {  // mul.lo.u64 + mul.hi.u64 (compiler 128-bit product)
                u64 lo = x * m, hi = __umul64hi(x, m);
                x = lo ^ hi;
            } else if (OP == 14) {  // 2 mad.wide + 2 iadd3 + 2 lop
                u32 a = (u32)x, b = (u32)(x >> 32);
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(a), "r"(m32));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(a), "r"(m32));
                x += gl_pack(a, b);
            }
            v[c] = x;
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= v[c];
    out[t] = acc;
}

template <int OP>
void run(const char *name, double ops_per_iter, const u64 *in, u64 *out, int sms) {
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
    double total = (double)blocks * 256 * ITERS * CHAINS * ops_per_iter;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_clk_sm = total / (ms * 1e-3) / (clk * 1e3) / sms;
    printf("%-28s %8.3f ms  %8.2f Gop/s  %7.2f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms,
           total / ms / 1e6, per_clk_sm, clk / 1000);
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
    run<0>("mad.wide.u32", 1, in, out, sms);
    run<1>("mad.lo.u32", 2, in, out, sms);
    run<2>("mad.hi.u32", 2, in, out, sms);
    run<3>("add.u64 (2 instr each)", 2, in, out, sms);
    run<4>("lop3", 2, in, out, sms);
    run<5>("shf", 2, in, out, sms);
    run<6>("mad.wide + add.u64", 2, in, out, sms);
    run<14>("mad.wide+add32+lop3+add64", 4, in, out, sms);
    run<13>("mul.lo.u64+mul.hi.u64", 1, in, out, sms);
    run<7>("gl_mul", 1, in, out, sms);
    run<8>("gl_add_weak", 1, in, out, sms);
    run<9>("gl_sub", 1, in, out, sms);
    run<10>("gl_canon", 1, in, out, sms);
    run<11>("gl_mul_pow2<39>", 1, in, out, sms);
    run<12>("gl_mul_pow2<14>", 1, in, out, sms);
    return 0;
}
