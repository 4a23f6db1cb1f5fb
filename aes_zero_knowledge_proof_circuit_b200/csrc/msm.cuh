// Declarations for the G1 MSM pipeline (msm.cu).
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "msm_core.cuh"  // MsmPlan, msm_make_plan, the per-thread kernel bodies

namespace zk {

// Device-resident inputs -> W window sums (XYZZ, device).  scalars: n x 8 u32 (canonical, or Montgomery if scalars_mont).
template <class C>
// scalar_stride: distance between consecutive scalars in elements (cyclic sharding reads every N-th coefficient).
// bases_internal != 0: bases are already in the internal packed form (msm_bases_to_internal); otherwise arkworks form.
int msm_window_sums(zkaes_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, int scalars_mont, const MsmPlan& p,
                    void* d_window_sums, int bases_internal = 0, size_t scalar_stride = 1);
template <class C>
int msm_bases_to_internal(zkaes_ctx* ctx, const void* src, void* dst, size_t n);

// Host Horner fold over n_sets sets of W window sums (one set per rank), returns the affine result.
template <class C>
Affine<C> msm_fold_windows_host(const XYZZ<C>* sums, int n_sets, const MsmPlan& p);

// Small-scalar MSM (signed integers |v| <= 2^(c-1), c <= 13): S pseudo-window sums (XYZZ, device) whose PLAIN sum is
// sum_j vals[val_start + j * val_stride] * bases[j], j < n  (msm.cu, "small-scalar MSM")
template <class C>
int msm_small_window_sums(zkaes_ctx* ctx, const void* d_bases, const int32_t* d_vals, size_t n, size_t val_start, size_t val_stride, int c, int S,
                          void* d_window_sums);
template <class C>
Affine<C> msm_sum_windows_host(const XYZZ<C>* sums, size_t count);

// out[i] = sum_{j < i} in[j]  (n <= 2^30)
int exclusive_scan_u32(zkaes_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n);

// Whole MSM: device-resident bases / scalars -> affine result on the host (window sums on the device, Horner fold on the host).
template <class C>
int msm_to_affine(zkaes_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, int scalars_mont, Affine<C>* out, int bases_internal = 0);

}  // namespace zk
