// On-device arithmetic self-test kernels (included twice: generated-PTX multiplier and portable CIOS).
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace zk {

template <class F>
__global__ void ZK_SELFTEST_NAME(k_selftest_field)(const F* a, const F* b, F* out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i], r;
    if (op == 0) r = x + y;
    else if (op == 1) r = x - y;
    else if (op == 3) r = x.sqr();
    else r = x * y;
    out[i] = r;
}

template <class F>
int ZK_SELFTEST_NAME(selftest_field)(zkaes_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
    cudaStream_t st = ctx->stream;
    DevBuf da, db, dout;
    size_t bytes = sizeof(F) * count;
    ZK_CUDA(ctx, da.alloc(bytes, st));
    ZK_CUDA(ctx, db.alloc(bytes, st));
    ZK_CUDA(ctx, dout.alloc(bytes, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(da.p, a, bytes, cudaMemcpyHostToDevice, st));
    ZK_CUDA(ctx, cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    ZK_SELFTEST_NAME(k_selftest_field)<F><<<cdiv(count, 128), 128, 0, st>>>(da.as<F>(), db.as<F>(), dout.as<F>(), count, op);
    ctx->launches++;
    ZK_CUDA(ctx, cudaGetLastError());
    ZK_CUDA(ctx, cudaMemcpyAsync(out, dout.p, bytes, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(ctx, cudaStreamSynchronize(st));
    return ZK_OK;
}

}  // namespace zk
