// Shared device/host helpers for the sm_100a kernels: error plumbing, mbarrier + TMA (cp.async.bulk.tensor)
// PTX wrappers, the FP64 tensor-core MMA (DMMA.8x8x4) wrapper and the tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ldm {

// ----------------------------------------------------------------------------------------------------------
// error handling: every C-ABI entry returns 0 or a negative code; the message is kept per thread
// ----------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
const char* last_error();

#define LDM_CUDA_OK(expr)                                                                           \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::ldm::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                             __FILE__ + ":" + std::to_string(__LINE__));                            \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

#define LDM_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            ::ldm::set_error(std::string("invalid argument: ") + msg + " (" #cond ") at " +         \
                             __FILE__ + ":" + std::to_string(__LINE__));                            \
            return -2;                                                                              \
        }                                                                                           \
    } while (0)

// Encode a rank-3 FP64 tensor map.  dims/strides are given in doubles / bytes, innermost first;
// box = (box0, box1, 1).  swizzle: 0 none, 1 = 128-byte swizzle (box0 must be 16 doubles), 2 = 64-byte swizzle
// (box0 must be 8 doubles).
int encode_tmap_f64_3d(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                       int swizzle);

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// TMA: rank-3 tiled load global -> shared, completion signalled on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// bulk copy global -> shared of a contiguous, 16-byte aligned range (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy / TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col)   (SASS: DMMA.8x8x4)
//   lane = 4*g + t :  a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    // volatile: keeps the MMAs in program order with the (volatile) fragment loads and barrier operations
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double xor_hi(double x, uint32_t mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
}

__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
    return x;
}
#endif  // __CUDACC__

}  // namespace ldm
