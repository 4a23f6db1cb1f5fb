// Integer-pipe micro-benchmarks for sm_100a (development aid; results are summarised in profiles/).
//   1. raw issue rates of IMAD / IMAD.HI / IMAD.WIDE / IADD3 chains  -> the "INT-ALU peak" the MSM/NTT rooflines use
//   2. Fq (12-limb) Montgomery multiplier variants: modmul/s per SM
//   3. XYZZ mixed addition throughput (the MSM inner loop)
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -I../aes_zero_knowledge_proof_circuit_b200/csrc ubench.cu -o ubench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ec.cuh"

using namespace zk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b = seed | 1, c = threadIdx.x * 2654435761u + seed;
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = c + i * 977; w[i] = ((uint64_t)a[i] << 32) | c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                // multiplier operand varies (low word of a neighbouring accumulator) so that ptxas cannot hoist the product
                if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"(c));
                if (MODE == 6) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(c)); a[i] = (uint32_t)t ^ (uint32_t)(t >> 32); }
                if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (MODE == 5) {  // carry-chained wide row as the multiplier emits it (8 limbs): 4 lo/hi pairs
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                    "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                    : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                    : "r"(b), "r"(c), "r"(b ^ 5), "r"(c ^ 9), "r"(b + c));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- Fq multiplier variants -------------------------------------------------------------------------------
__constant__ uint32_t c_mod[12];

// VAR 0: production operator* (modulus as immediates).  VAR 1: modulus from __constant__.  VAR 2: modulus in registers
// (loaded from global memory at run time).  VAR 3: portable CIOS.
template <int VAR, int ILP>
__global__ void __launch_bounds__(128) k_fqmul(uint32_t* out, const uint32_t* gmod, int iters) {
    using F = Fq377;
    F x[ILP], y;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 12; ++i) x[k].v[i] = (threadIdx.x + 1) * (i + 3 + k) * 2654435761u >> (i == 11 ? 8 : 0);
    for (int i = 0; i < 12; ++i) y.v[i] = (blockIdx.x + 7) * (i + 11) * 40503u >> (i == 11 ? 8 : 0);
    uint32_t m[12];
    for (int i = 0; i < 12; ++i) m[i] = gmod[i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (VAR == 0) {
                x[k] = x[k] * y;
            } else if (VAR == 1) {
                F r;
                mont_mul_raw_12(r.v, x[k].v, y.v, c_mod, zk_c_mont_inv[1], zk_c_zero);
                r.reduce_once();
                x[k] = r;
            } else if (VAR == 2) {
                F r;
                mont_mul_raw_12(r.v, x[k].v, y.v, m, zk_c_mont_inv[1], zk_c_zero);
                r.reduce_once();
                x[k] = r;
            } else {
                F r;
                F::mont_mul_portable(r.v, x[k].v, y.v);
                r.reduce_once();
                x[k] = r;
            }
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 12; ++i) s ^= x[k].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int BS>
__global__ void __launch_bounds__(BS) k_madd(const G1Affine377* pts, int npts, G1XYZZ377* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    G1XYZZ377 acc = G1XYZZ377::from_affine(pts[t % npts]);
    for (int it = 0; it < iters; ++it) {
        G1Affine377 p = pts[(t * 7 + it * 13 + 1) % npts];
        acc.madd(p);
    }
    out[t] = acc;
}

// same loop in the radix-2^29 internal form (fq29.cuh); MINB = min blocks per SM for __launch_bounds__ (register cap)
template <int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) k_madd29(const G1Affine377* pts, int npts, XYZZ<G1_377R29>* out, int iters) {
    using Fq = Fq29<Fq377R29Params>;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Affine<G1_377R29> p0;
    p0.x = Fq::from_std(pts[t % npts].x.v);
    p0.y = Fq::from_std(pts[t % npts].y.v);
    XYZZ<G1_377R29> acc = XYZZ<G1_377R29>::from_affine(p0);
    for (int it = 0; it < iters; ++it) {
        const G1Affine377& q = pts[(t * 7 + it * 13 + 1) % npts];
        Affine<G1_377R29> p;
        p.x = Fq::unpack(q.x.v);   // treat the stored words as already-internal representatives (any field elements do for timing)
        p.y = Fq::unpack(q.y.v);
        p.x.v[12] &= 0xffffff; p.y.v[12] &= 0xffffff;  // keep them < p
        acc.madd(p);
    }
    out[t] = acc;
}
template <int ILP>
__global__ void __launch_bounds__(128) k_fqmul29(uint32_t* out, int iters) {
    using F = Fq29<Fq377R29Params>;
    F x[ILP], y;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 13; ++i) x[k].v[i] = ((threadIdx.x + 1) * (i + 3 + k) * 2654435761u) >> (i == 12 ? 9 : 3);
    for (int i = 0; i < 13; ++i) y.v[i] = ((threadIdx.x * 31 + 7) * (i + 11) * 40503u) >> (i == 12 ? 9 : 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = x[k] * y;
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 13; ++i) s ^= x[k].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_make_pts(G1Affine377* pts, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    G1XYZZ377 p = G1XYZZ377::from_affine(G1Affine377::generator());
    G1XYZZ377 acc = G1XYZZ377::inf();
    for (int b = 0; b < 20; ++b) {
        if (((t + 1) >> b) & 1) acc.add(p);
        p = p.dbl();
    }
    pts[t] = acc.to_affine();
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, clk_khz / 1000);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint32_t* out; CK(cudaMalloc(&out, 1 << 26));
    const char* names[] = {"IMAD(lo)", "IMAD.HI", "IMAD.WIDE acc", "IADD", "LOP3", "madc lo/hi chain", "IMAD.WIDE (RZ) + LOP3"};
#define RUN_PIPE(M) { \
        int iters = 2000, blocks = sms * 8; \
        k_pipe<M><<<blocks, 256>>>(out, 10, 1); \
        cudaEventRecord(e0); k_pipe<M><<<blocks, 256>>>(out, iters, 3); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
        double ops = (double)blocks * 256 * iters * 64.0; float ms = time_ms(e0, e1); \
        printf("pipe %-18s: %.2f Tops/s  (%.1f ops/clk/SM at %d MHz nominal)\n", names[M], ops / ms / 1e9, ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000); }
    RUN_PIPE(0) RUN_PIPE(1) RUN_PIPE(2) RUN_PIPE(3) RUN_PIPE(4) RUN_PIPE(5) RUN_PIPE(6)

    uint32_t hmod[12];
    for (int i = 0; i < 12; ++i) hmod[i] = Fq377Params::MOD(i);
    CK(cudaMemcpyToSymbol(c_mod, hmod, sizeof(hmod)));
    uint32_t* gmod; CK(cudaMalloc(&gmod, 48)); CK(cudaMemcpy(gmod, hmod, 48, cudaMemcpyHostToDevice));
#define RUN_MUL(V, ILP, NAME) { \
        int iters = 2000; \
        for (int bps = 2; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_fqmul<V, ILP><<<blocks, 128>>>(out, gmod, 5); \
            cudaEventRecord(e0); k_fqmul<V, ILP><<<blocks, 128>>>(out, gmod, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double muls = (double)blocks * 128 * iters * ILP; float ms = time_ms(e0, e1); \
            printf("fqmul %-22s ilp=%d blocks/SM=%d: %.2f Gmul/s (%.1f clk/SM/mul)\n", NAME, ILP, bps, muls / ms / 1e6, (ms * 1e-3) * clk_khz * 1e3 * sms / muls); } }
    RUN_MUL(0, 1, "imm modulus") RUN_MUL(0, 2, "imm modulus")
    RUN_MUL(1, 1, "__constant__ modulus") RUN_MUL(1, 2, "__constant__ modulus")
    RUN_MUL(2, 1, "register modulus") RUN_MUL(2, 2, "register modulus")
    RUN_MUL(3, 1, "portable CIOS")

    int npts = 4096;
    G1Affine377* pts; CK(cudaMalloc(&pts, sizeof(G1Affine377) * npts));
    k_make_pts<<<npts / 64, 64>>>(pts, npts);
    CK(cudaDeviceSynchronize());
    G1XYZZ377* acc; CK(cudaMalloc(&acc, sizeof(G1XYZZ377) * sms * 16 * 256));
#define RUN_MADD(BS) { \
        int iters = 500; \
        for (int bps = 1; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_madd<BS><<<blocks, BS>>>(pts, npts, acc, 3); \
            cudaEventRecord(e0); k_madd<BS><<<blocks, BS>>>(pts, npts, acc, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double n = (double)blocks * BS * iters; float ms = time_ms(e0, e1); \
            printf("madd bs=%d blocks/SM=%d: %.1f Mmadd/s (%.0f clk/SM/madd)\n", BS, bps, n / ms / 1e3, (ms * 1e-3) * clk_khz * 1e3 * sms / n); } }
    RUN_MADD(64) RUN_MADD(128)
#define RUN_MUL29(ILP) { \
        int iters = 2000; \
        for (int bps = 2; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_fqmul29<ILP><<<blocks, 128>>>(out, 5); \
            cudaEventRecord(e0); k_fqmul29<ILP><<<blocks, 128>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double muls = (double)blocks * 128 * iters * ILP; float ms = time_ms(e0, e1); \
            printf("fqmul29 ilp=%d blocks/SM=%d: %.2f Gmul/s (%.1f clk/SM/mul)\n", ILP, bps, muls / ms / 1e6, (ms * 1e-3) * clk_khz * 1e3 * sms / muls); } }
    RUN_MUL29(1) RUN_MUL29(2)
    XYZZ<G1_377R29>* acc29; CK(cudaMalloc(&acc29, sizeof(XYZZ<G1_377R29>) * sms * 16 * 256));
#define RUN_MADD29(BS, MINB) { \
        int iters = 500; \
        for (int bps = 1; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_madd29<BS, MINB><<<blocks, BS>>>(pts, npts, acc29, 3); \
            cudaEventRecord(e0); k_madd29<BS, MINB><<<blocks, BS>>>(pts, npts, acc29, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double n = (double)blocks * BS * iters; float ms = time_ms(e0, e1); \
            printf("madd29 bs=%d minb=%d blocks/SM=%d: %.1f Mmadd/s (%.0f clk/SM/madd)\n", BS, MINB, bps, n / ms / 1e3, (ms * 1e-3) * clk_khz * 1e3 * sms / n); } }
    RUN_MADD29(128, 1) RUN_MADD29(128, 2) RUN_MADD29(128, 3) RUN_MADD29(128, 4) RUN_MADD29(64, 6)
    CK(cudaDeviceSynchronize());
    return 0;
}
