// Declarations for the device-side test-SRS generator (srs.cu).
#pragma once
#include "common.cuh"
#include "ec.cuh"
namespace zk {
template <class C>
// out[i] = tau^(start + i * stride) * G, i < n
int srs_powers_device(zkaes_ctx* ctx, const uint8_t seed32[32], size_t n, void* d_out, size_t start = 0, size_t stride = 1);
template <class C>
// out[i] = scalars[start + i * stride] * G, i < n; scalars: Montgomery Fr elements on the device
int fb_mul_scalars_device(zkaes_ctx* ctx, const void* d_scalars, size_t n, void* d_out, size_t start = 0, size_t stride = 1);
}
