// Declarations for the device-side test-SRS generator (srs.cu).
#pragma once
#include "common.cuh"
#include "ec.cuh"
namespace zk {
template <class C>
int srs_powers_device(zkaes_ctx* ctx, const uint8_t seed32[32], size_t n, void* d_out);
}
