// Integer-pipe micro-benchmarks for sm_100a (development aid; results are summarised in profiles/).
//   1. raw issue rates of IMAD / IMAD.HI / IMAD.WIDE / IADD3 chains  -> the "INT-ALU peak" the MSM/NTT rooflines use
//   2. Fq (12-limb) Montgomery multiplier variants: modmul/s per SM
//   3. XYZZ mixed addition throughput (the MSM inner loop)
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -I../aes_zero_knowledge_proof_circuit_b200/csrc ubench.cu -o ubench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ec.cuh"
#include "fq52.cuh"

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

// ---- FP64 pipe (round 2): DFMA issue rate alone, with a 64-bit integer add per DFMA (the hi/lo limb-product scheme adds the
// raw bit patterns of the DFMA results as integers), and DFMA interleaved with IMAD.WIDE (do the two pipes overlap?)
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe64(double* out, int iters, double seed) {
    double d[8], b = seed + 1.0, c = seed * 3.0 + threadIdx.x;
    uint64_t w[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = c + i * 977.0; w[i] = (uint64_t)(threadIdx.x * 2654435761u + i) << 7; acc[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 3) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(b), "d"(c));
                if (MODE == 1) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((uint64_t)__double_as_longlong(d[i])));
                if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"((uint32_t)threadIdx.x | 1u));
                if (MODE == 3 && (i & 1)) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"((uint32_t)threadIdx.x | 1u));
                if (MODE == 4) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += d[i] + (double)(w[i] ^ acc[i]);
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

// CALL = 0: ten inlined products per mixed addition (round 1); 1: ten calls of one out-of-line multiplier (XYZZ::madd_call)
template <int BS, int CALL>
__global__ void __launch_bounds__(BS) k_madd(const G1Affine377* pts, int npts, G1XYZZ377* out, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    G1XYZZ377 acc = G1XYZZ377::from_affine(pts[t % npts]);
    for (int it = 0; it < iters; ++it) {
        G1Affine377 p = pts[(t * 7 + it * 13 + 1) % npts];
        if (CALL) acc.madd_call(p); else acc.madd(p);
    }
    out[t] = acc;
}
// the out-of-line multiplier alone (call overhead included), ILP independent chains per thread
template <int ILP>
__global__ void __launch_bounds__(128) k_fqmul_call(uint32_t* out, int iters) {
    using F = Fq377;
    F x[ILP], y;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 12; ++i) x[k].v[i] = (threadIdx.x + 1) * (i + 3 + k) * 2654435761u >> (i == 11 ? 8 : 0);
    for (int i = 0; i < 12; ++i) y.v[i] = (blockIdx.x + 7) * (i + 11) * 40503u >> (i == 11 ? 8 : 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = F::mul_call(x[k], y);
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k)
        for (int i = 0; i < 12; ++i) s ^= x[k].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- FP64-limb Fq product (csrc/fq52.cuh): alone, and warp-specialised next to the IMAD.WIDE product -----------------------------------
// DF warps of every block (warp index < DF of 4) run the 52-bit-limb DFMA product, the others the 32-bit-limb IMAD.WIDE product; each kind
// gets its own iteration count so that both finish together.  DF = 4: FP64 only; DF = 0: IMAD.WIDE only.
template <int DF>
__global__ void __launch_bounds__(128) k_fq_hybrid(uint32_t* out, int iters_imad, int iters_dfma) {
    const int warp = threadIdx.x >> 5;
    uint32_t s = 0;
    if (warp < DF) {
        Fq52 x, y;
        for (int i = 0; i < 8; ++i) {
            x.v[i] = (double)(((uint64_t)(threadIdx.x + 1) * (i + 3) * 0x9e3779b97f4aull) & (i == 7 ? 0xfffull : ZK52_MASK));
            y.v[i] = (double)(((uint64_t)(blockIdx.x + 7) * (i + 11) * 0xc2b2ae3d27d4ull) & (i == 7 ? 0xfffull : ZK52_MASK));
        }
        for (int it = 0; it < iters_dfma; ++it) x = fq52_mont_mul<Fq377P52>(x, y);
        for (int i = 0; i < 8; ++i) s ^= (uint32_t)__double_as_longlong(x.v[i]);
    } else {
        Fq377 x, y;
        for (int i = 0; i < 12; ++i) {
            x.v[i] = (threadIdx.x + 1) * (i + 3) * 2654435761u >> (i == 11 ? 8 : 0);
            y.v[i] = (blockIdx.x + 7) * (i + 11) * 40503u >> (i == 11 ? 8 : 0);
        }
        for (int it = 0; it < iters_imad; ++it) x = Fq377::mul_call(x, y);
        for (int i = 0; i < 12; ++i) s ^= x.v[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// bit-exactness of the FP64 product on the device: a, b as 12 words -> 52-bit limbs -> product -> 12 words (Montgomery form with R = 2^416)
__global__ void k_fq52_check(const uint32_t* a, const uint32_t* b, uint32_t* o, int n) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Fq52 r = fq52_mont_mul<Fq377P52>(fq52_from_words(a + 12 * t), fq52_from_words(b + 12 * t));
    fq52_to_words(r, o + 12 * t);
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

    {
        const char* n64[] = {"DFMA.RZ", "DFMA.RZ + IADD.64 each", "DFMA.RZ + IMAD.WIDE each", "DFMA.RZ + IMAD.WIDE per 2", "DADD.RZ"};
        double* o64 = reinterpret_cast<double*>(out);
#define RUN_PIPE64(M, PER) { \
        int iters = 2000, blocks = sms * 8; \
        k_pipe64<M><<<blocks, 256>>>(o64, 10, 1.0); \
        cudaEventRecord(e0); k_pipe64<M><<<blocks, 256>>>(o64, iters, 3.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
        double ops = (double)blocks * 256 * iters * 64.0; float ms = time_ms(e0, e1); \
        printf("pipe64 %-26s: %.2f T DFMA/s  (%.1f DFMA/clk/SM at %d MHz nominal; %s)\n", n64[M], ops / ms / 1e9, ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, PER); }
        RUN_PIPE64(0, "alone") RUN_PIPE64(1, "one 64-bit integer add per DFMA") RUN_PIPE64(2, "one IMAD.WIDE per DFMA") RUN_PIPE64(3, "one IMAD.WIDE per two DFMA") RUN_PIPE64(4, "DADD instead of DFMA")
    }
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

    {
        // device check of the FP64 product against the 32-bit-limb multiplier: both compute a*b/R mod q with their own R (2^416 vs 2^384),
        // so compare fq52(a, b) with imad(imad(a, b), 2^352 mod q in R=2^384 form)... simpler: the host recomputes with __int128-free schoolbook
        // in tests/test_fq52.py (CPU tier, same header under round-toward-zero); here only a smoke value is printed.
        int iters = 1000;
#define RUN_HYB(DF, II, ID) { \
        for (int bps = 4; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_fq_hybrid<DF><<<blocks, 128>>>(out, 5, 5); \
            cudaEventRecord(e0); k_fq_hybrid<DF><<<blocks, 128>>>(out, II, ID); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double muls = (double)blocks * 32.0 * ((4 - DF) * (double)(II) + DF * (double)(ID)); float ms = time_ms(e0, e1); \
            printf("fq product, %d of 4 warps on the FP64 pipe (iters imad %d / dfma %d) blocks/SM=%d: %.2f Gmul/s (%.2f clk/SM/mul)\n", DF, II, ID, bps, muls / ms / 1e6, (ms * 1e-3) * clk_khz * 1e3 * sms / muls); } }
        RUN_HYB(0, iters, 0) RUN_HYB(4, 0, iters)
        RUN_HYB(1, iters, iters) RUN_HYB(1, iters, 2 * iters) RUN_HYB(1, iters, 3 * iters)
        RUN_HYB(2, iters, iters) RUN_HYB(2, iters, iters / 2) RUN_HYB(2, iters, 3 * iters / 2)
        RUN_HYB(3, iters, iters / 2) RUN_HYB(3, iters, iters / 3)
    }
    int npts = 4096;
    G1Affine377* pts; CK(cudaMalloc(&pts, sizeof(G1Affine377) * npts));
    k_make_pts<<<npts / 64, 64>>>(pts, npts);
    CK(cudaDeviceSynchronize());
    G1XYZZ377* acc; CK(cudaMalloc(&acc, sizeof(G1XYZZ377) * sms * 16 * 256));
#define RUN_MADD(BS, CALL) { \
        int iters = 500; \
        for (int bps = 1; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_madd<BS, CALL><<<blocks, BS>>>(pts, npts, acc, 3); \
            cudaEventRecord(e0); k_madd<BS, CALL><<<blocks, BS>>>(pts, npts, acc, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double n = (double)blocks * BS * iters; float ms = time_ms(e0, e1); \
            printf("madd %s bs=%d blocks/SM=%d: %.1f Mmadd/s (%.0f clk/SM/madd)\n", CALL ? "out-of-line products" : "inlined products    ", BS, bps, n / ms / 1e3, (ms * 1e-3) * clk_khz * 1e3 * sms / n); } }
    RUN_MADD(64, 0) RUN_MADD(128, 0) RUN_MADD(64, 1) RUN_MADD(128, 1)
#define RUN_MULC(ILP) { \
        int iters = 2000; \
        for (int bps = 2; bps <= 8; bps *= 2) { \
            int blocks = sms * bps; \
            k_fqmul_call<ILP><<<blocks, 128>>>(out, 5); \
            cudaEventRecord(e0); k_fqmul_call<ILP><<<blocks, 128>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
            double muls = (double)blocks * 128 * iters * ILP; float ms = time_ms(e0, e1); \
            printf("fqmul out-of-line ilp=%d blocks/SM=%d: %.2f Gmul/s (%.1f clk/SM/mul)\n", ILP, bps, muls / ms / 1e6, (ms * 1e-3) * clk_khz * 1e3 * sms / muls); } }
    RUN_MULC(1) RUN_MULC(2)
    CK(cudaDeviceSynchronize());
    return 0;
}
