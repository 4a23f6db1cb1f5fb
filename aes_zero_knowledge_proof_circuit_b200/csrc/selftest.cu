// Self-test instantiation with the production (generated PTX) multiplier + G1 formula checks.
#define ZK_SELFTEST_NAME(x) x##_asm
#include "selftest_impl.cuh"
#include "fq52.cuh"
namespace zk {
template int selftest_field_asm<Fr377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_asm<Fq377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_asm<Fr381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_asm<Fq381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);

// variant 2: the out-of-line multiplier the MSM inner loop calls (Fp::mul_call); add / sub as in variant 0
template <class F>
__global__ void k_selftest_field_call(const F* a, const F* b, F* out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i], r;
    if (op == 0) r = x + y;
    else if (op == 1) r = x - y;
    else if (op == 3) r = F::sqr_call(x);
    else r = F::mul_call(x, y);
    out[i] = r;
}
template <class F>
int selftest_field_call(zkaes_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
    cudaStream_t st = ctx->stream;
    DevBuf da, db, dout;
    size_t bytes = sizeof(F) * count;
    ZK_CUDA(ctx, da.alloc(bytes, st));
    ZK_CUDA(ctx, db.alloc(bytes, st));
    ZK_CUDA(ctx, dout.alloc(bytes, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    k_selftest_field_call<F><<<cdiv(count, 128), 128, 0, st>>>(da.as<F>(), db.as<F>(), dout.as<F>(), count, op);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return ZK_OK;
}
template int selftest_field_call<Fr377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_call<Fq377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_call<Fr381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_call<Fq381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);

// variant 3: the FP64-limb product of csrc/fq52.cuh (BLS12-377 Fq only; an experiment, not used by the prover): a, b as 12 words ->
// 8 x 52-bit limbs in doubles -> DFMA hi/lo Montgomery product -> 12 words.  Its Montgomery radix is 2^416: out = a b 2^-416 mod q.
__global__ void k_selftest_fq52(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fq52 r = fq52_mont_mul<Fq377P52>(fq52_from_words(a + 12 * i), fq52_from_words(b + 12 * i));
    fq52_to_words(r, out + 12 * i);
}
int selftest_fq52(zkaes_ctx* ctx, const void* a, const void* b, void* out, size_t count) {
    cudaStream_t st = ctx->stream;
    DevBuf da, db, dout;
    const size_t bytes = 48 * count;
    ZK_CUDA(ctx, da.alloc(bytes, st));
    ZK_CUDA(ctx, db.alloc(bytes, st));
    ZK_CUDA(ctx, dout.alloc(bytes, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    k_selftest_fq52<<<cdiv(count, 128), 128, 0, st>>>(da.as<uint32_t>(), db.as<uint32_t>(), dout.as<uint32_t>(), count);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return ZK_OK;
}

template <class C>
__global__ void k_selftest_g1(const Affine<C>* a, const Affine<C>* b, Affine<C>* out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<C> acc = XYZZ<C>::from_affine(a[i]);
    if (op == 0) {
        acc.madd(b[i]);
    } else if (op == 1) {
        // make the second operand non-trivially projective: (b + a) - a computed as full adds
        XYZZ<C> q = XYZZ<C>::from_affine(b[i]);
        q = q.dbl();              // 2b
        q.add(XYZZ<C>::from_affine(b[i]).neg());  // 2b - b = b, with ZZ != 1
        acc.add(q);
    } else {
        acc = acc.dbl();
    }
    out[i] = acc.to_affine();
}

template <class C>
int selftest_g1(zkaes_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
    cudaStream_t st = ctx->stream;
    DevBuf da, db, dout;
    size_t bytes = sizeof(Affine<C>) * count;
    ZK_CUDA(ctx, da.alloc(bytes, st));
    ZK_CUDA(ctx, db.alloc(bytes, st));
    ZK_CUDA(ctx, dout.alloc(bytes, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    k_selftest_g1<C><<<cdiv(count, 64), 64, 0, st>>>(da.as<Affine<C>>(), db.as<Affine<C>>(), dout.as<Affine<C>>(), count, op);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return ZK_OK;
}
template int selftest_g1<G1_377Params>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_g1<G1_381Params>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
}
