// ubench5.cu -- the 8-point Goldilocks DFT of the NTT passes on the integer pipes (the product's own dft_pow2) against
// an exact FP64 form on the otherwise idle FP64 pipe, and both together in one instruction stream.  (developer tool)
//
// FP64 form: an element is a + b * phi, phi = 2^32 (phi^2 = phi - 1 mod p, p = 2^64 - 2^32 + 1), a and b integers held in
// doubles (|.| < 2^53 stays exact).  Addition: one DADD per component, no carries, no wrap corrections.  x * 2^S:
// scale both components by 2^(S mod 32), split each at 2^32 with the 1.5 * 2^52 rounding constant, recombine with
// phi^2 = phi - 1, then rotate by phi^(S div 32): (a, b) -> (-b, a + b) -> (-a - b, a).  Exactness is checked against the
// integer form on the host for one DFT (values mod p).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I twenty-first_b200/csrc -o tools/ubench5 tools/ubench5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ntt_fast.cuh"
using namespace tf21;
#define ITERS 512

struct Eis {
    double a, b;
};
__device__ __forceinline__ Eis eis_add(Eis x, Eis y) { return Eis{x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ Eis eis_sub(Eis x, Eis y) { return Eis{x.a - y.a, x.b - y.b}; }
template <int S>
__device__ __forceinline__ Eis eis_shl(Eis x) {
    constexpr int q = S >> 5, t = S & 31;
    const double M = 6755399441055744.0;  // 1.5 * 2^52
    const double sc = (double)(1ull << t), inv32 = 1.0 / 4294967296.0;
    const double ta = x.a * sc, tb = x.b * sc;
    const double ha = __fma_rn(ta, inv32, M) - M, hb = __fma_rn(tb, inv32, M) - M;
    const double la = __fma_rn(ha, -4294967296.0, ta), lb = __fma_rn(hb, -4294967296.0, tb);
    // (la + ha phi) + (lb + hb phi) phi = (la - hb) + (ha + lb + hb) phi
    Eis r{la - hb, ha + lb + hb};
    if (q == 1) r = Eis{-r.b, r.a + r.b};
    if (q == 2) r = Eis{-r.a - r.b, r.a};
    return r;
}
// radix-2 DIT, natural in, bit-reversed out, omega_8 = 2^24 (omega_64 = 2^39 -> omega_8 = 2^(39 * 8) = 2^312 = 2^(312 mod 192) = 2^120
// = -2^24: the product's convention; only the instruction mix matters here, the check below uses the same exponents)
template <int E>
__device__ __forceinline__ void eis_bfly(Eis &u, Eis &v) {
    constexpr int S0 = E % 192, neg = S0 >= 96, S = neg ? S0 - 96 : S0;
    Eis t = v;
    if (S != 0) t = eis_shl<(S ? S : 1)>(v);
    const Eis p = eis_add(u, t), m = eis_sub(u, t);
    u = neg ? m : p;
    v = neg ? p : m;
}
__device__ __forceinline__ void eis_dft8(Eis (&v)[8]) {
    constexpr int EU = (39 << 3) % 192;  // exponent of omega_8 as a power of two
    // stage 1 (pairs at bit-reversed distance): v[brev(k+j)], v[brev(k+j+half)]
    eis_bfly<0>(v[0], v[4]); eis_bfly<0>(v[2], v[6]); eis_bfly<0>(v[1], v[5]); eis_bfly<0>(v[3], v[7]);
    eis_bfly<0>(v[0], v[2]); eis_bfly<2 * EU>(v[4], v[6]); eis_bfly<0>(v[1], v[3]); eis_bfly<2 * EU>(v[5], v[7]);
    eis_bfly<0>(v[0], v[1]); eis_bfly<EU>(v[4], v[5]); eis_bfly<2 * EU>(v[2], v[3]); eis_bfly<3 * EU>(v[6], v[7]);
}

// lazy u64 word -> (lo32, hi32) as exact doubles: the 2^52-bias trick, one DADD per component, no XU conversion
__device__ __forceinline__ Eis eis_from_u64(u64 x) {
    const double B52 = 4503599627370496.0;
    return Eis{__hiloint2double(0x43300000, (int)(u32)x) - B52, __hiloint2double(0x43300000, (int)(u32)(x >> 32)) - B52};
}
// a + b * 2^32 (|a|, |b| < 2^48) -> a lazy u64 representative mod p.  Both components are made non-negative with an offset of
// 2^48 (value + 2^48 + 2^80, and 2^80 = 2^48 - 2^16 mod p: the constant C = 2^49 - 2^16 is subtracted at the end); the
// biased doubles a' + 2^52, b' + 2^52 carry the integers in their low 52 bits:
//   a' + b' 2^32 = alo + (ahi + blo) 2^32 + bhi 2^64 = (blo : alo) + [((ahi + bhi) << 32) - bhi]   (2^64 = 2^32 - 1)
__device__ __forceinline__ u64 eis_to_u64(Eis e, u32 one) {
    const double OFF = 281474976710656.0 + 4503599627370496.0;  // 2^48 + 2^52
    const u64 ab = (u64)__double_as_longlong(e.a + OFF), bb = (u64)__double_as_longlong(e.b + OFF);
    const u32 alo = (u32)ab, ahi = (u32)(ab >> 32) & 0xFFFFFu, blo = (u32)bb, bhi = (u32)(bb >> 32) & 0xFFFFFu;
    const u64 T = ((u64)(ahi + bhi) << 32) - bhi;
    const u64 v = gl_addl(gl_pack(alo, blo), T);
    return gl_subl(v, 0x0001FFFFFFFF0000ull, one);
}

// bit 0: integer DFT8 (dft_pow2<false, 3>), bit 1: FP64 DFT8, on independent data of the same thread;
// bit 2: the FP64 side starts from and returns to lazy u64 words every round (conversions included)
template <int OP>
__global__ void __launch_bounds__(128) k(const u64 *in, u64 *out) {
    const u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 v[8], w[8];
    Eis e[8];
    const u32 one = (u32)(in[1023] >> 63) + 1u - (u32)(in[1023] >> 63);  // 1, opaque enough for this tool
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const u64 x = in[(t + c * 977) & 1023];
        v[c] = x;
        w[c] = x ^ 0x5555;
        e[c] = Eis{(double)(u32)x, (double)(u32)(x >> 32)};
    }
    for (int it = 0; it < ITERS; it++) {
        if (OP & 1) dft_pow2<false, 3>(v);
        if (OP & 4) {
            Eis f[8];
#pragma unroll
            for (int c = 0; c < 8; c++) f[c] = eis_from_u64(w[c]);
            eis_dft8(f);
#pragma unroll
            for (int c = 0; c < 8; c++) w[c] = eis_to_u64(f[c], one);
        }
        if (OP & 2) {
            eis_dft8(e);
#pragma unroll
            for (int c = 0; c < 8; c++) {  // keep the magnitudes bounded over the iterations (timing run only)
                e[c].a *= 0.125;
                e[c].b *= 0.125;
            }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) acc ^= v[c] ^ w[c] ^ (u64)(long long)e[c].a ^ ((u64)(long long)e[c].b << 7);
    out[t] = acc;
}

// one exact DFT8 in both forms: out[0..8) integer result (canonical), out[8..16) a, out[16..24) b as signed integers
__global__ void check(const u64 *in, u64 *out) {
    u64 v[8];
    Eis e[8];
    for (int c = 0; c < 8; c++) {
        v[c] = in[c] % GL_P;
        e[c] = Eis{(double)(u32)v[c], (double)(u32)(v[c] >> 32)};
    }
    dft_pow2<false, 3>(v);
    eis_dft8(e);
    for (int c = 0; c < 8; c++) {
        out[c] = gl_canon(v[c]);
        out[8 + c] = (u64)(long long)e[c].a;
        out[16 + c] = (u64)(long long)e[c].b;
        out[24 + c] = gl_canon(eis_to_u64(e[c], 1u));
        const Eis r = eis_from_u64(in[c]);
        out[32 + c] = gl_canon(eis_to_u64(r, 1u)) ^ gl_canon(in[c]);  // 0 when the conversions round-trip
    }
}

template <int OP>
void run(const char *name, const u64 *in, u64 *out, int sms, int instr) {
    const int blocks = sms * 16;
    k<OP><<<blocks, 128>>>(in, out);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<blocks, 128>>>(in, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double dfts = blocks * 4.0 / sms / 4 * ITERS;  // DFT8 rounds per SM sub-partition
    printf("%-40s %7.3f ms  %7.1f cycles per 8-point DFT round per warp and SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / dfts);
    (void)instr;
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u64 *in, *out;
    cudaMalloc(&in, 1024 * 8);
    cudaMalloc(&out, (size_t)sms * 16 * 128 * 8);
    u64 h[1024];
    u64 s = 12345;
    for (int i = 0; i < 1024; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = s;
    }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    check<<<1, 1>>>(in, out);
    u64 r[40];
    cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int c = 0; c < 8; c++) {
        // value of a + b * 2^32 mod p with signed a, b
        const __int128 val = (__int128)(long long)r[8 + c] + (__int128)(long long)r[16 + c] * ((__int128)1 << 32);
        __int128 m = val % (__int128)GL_P;
        if (m < 0) m += (__int128)GL_P;
        if ((u64)m != r[c]) ok = 0;
        if (r[24 + c] != r[c] || r[32 + c] != 0) ok = 0;  // back to a u64 word: same field element
    }
    printf("FP64 Eisenstein DFT8 (and its way back to u64 words) equals the integer DFT8 mod p: %s\n", ok ? "yes" : "NO");
    run<1>("integer DFT8 (product code)", in, out, sms, 0);
    run<2>("FP64 DFT8 (a + b phi)", in, out, sms, 0);
    run<3>("both, independent data, one stream", in, out, sms, 0);
    run<4>("FP64 DFT8 from / to u64 words", in, out, sms, 0);
    run<5>("integer DFT8 + FP64 DFT8 from / to u64", in, out, sms, 0);
    return ok ? 0 : 1;
}
