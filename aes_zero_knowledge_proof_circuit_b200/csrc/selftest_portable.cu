// Self-test instantiation with the portable (CIOS) multiplier forced on the device.
#define ZK_FF_PORTABLE 1
#define ZK_SELFTEST_NAME(x) x##_portable
#include "selftest_impl.cuh"
namespace zk {
template int selftest_field_portable<Fr377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_portable<Fq377>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_portable<Fr381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
template int selftest_field_portable<Fq381>(zkaes_ctx*, int, const void*, const void*, void*, size_t);
}
