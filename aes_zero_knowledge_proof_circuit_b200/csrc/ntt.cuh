// Declarations for the Fr NTT (ntt.cu).
#pragma once
#include "common.cuh"
#include "ff.cuh"

namespace zk {
// In-place transform of 2^log_n Montgomery Fr elements resident on the device (natural order in and out).
template <class FrP>
int ntt_device(zkaes_ctx* ctx, int curve_id, void* d_data, int log_n, int inverse, int coset);
}  // namespace zk
