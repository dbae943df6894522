// DMMA + shared-memory-load probe (development aid): the inner loop of the 3M complex GEMM without TMA, barriers or
// epilogues.  Per "stage" a warp loads 2 A fragments and, per column fragment, NL 128-bit B words, and issues 6 DMMAs
// per fragment (FB fragments).  Reports clocks per DMMA and sub-partition for 1 - 3 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dmma_probe2 tools/dmma_probe2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}

template <int FB, int NL, int JP, int BAR>
__global__ void __launch_bounds__(384, 1) probe(double* out, long long* clk, int iters) {
    extern __shared__ __align__(16) uint8_t smem[];
    for (int i = threadIdx.x; i < 48 * 1024 / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1e-9 * i;
    // BAR: per stage a try_wait on an mbarrier whose phase is already complete (1), plus an arrive on another (2)
    __shared__ __align__(8) unsigned long long bars[2];
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]), bar1 = bar0 + 8;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar1), "r"(1 << 19));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0) : "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t a_addr = base + (warp * 8 + g) * 128 + t * 16;
    const uint32_t b_addr = base + 8192 + g * 64 + t * 16;
    double cr[FB][2], ci[FB][2], cs[FB][2];
#pragma unroll
    for (int j = 0; j < FB; ++j) cr[j][0] = cr[j][1] = ci[j][0] = ci[j][1] = cs[j][0] = cs[j][1] = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t so = (uint32_t)(it & 1) * 24576;      // two "stages"
        if (BAR >= 1) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "WAIT_LOOP:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE;\n\t"
                "bra WAIT_LOOP;\n\t"
                "DONE:\n\t"
                "}" ::"r"(bar0), "r"(0) : "memory");
        }
        if (BAR >= 2) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar1) : "memory");
        }
        double ar[2], ai[2], an[2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            lds128(a_addr + so + kk * 64, ar[kk], ai[kk]);
            an[kk] = ar[kk] + ai[kk];
        }
        constexpr int NG = (FB + JP - 1) / JP;
        double f[JP][3][2], gn[JP][3][2];
#pragma unroll
        for (int jj = 0; jj < JP; ++jj)
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                gn[jj][pl][0] = gn[jj][pl][1] = 0.0;
                if (pl < NL) lds128(b_addr + so + jj * 512 + pl * 5120, f[jj][pl][0], f[jj][pl][1]);
                else { f[jj][pl][0] = ar[0]; f[jj][pl][1] = ai[0]; }
            }
#pragma unroll
        for (int jg = 0; jg < NG; ++jg) {
#pragma unroll
            for (int jj = 0; jj < JP; ++jj)
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                    if ((jg + 1) * JP + jj < FB) {
                        if (pl < NL) lds128(b_addr + so + ((jg + 1) * JP + jj) * 512 + pl * 5120, gn[jj][pl][0], gn[jj][pl][1]);
                        else { gn[jj][pl][0] = ar[1]; gn[jj][pl][1] = ai[1]; }
                    }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int jj = 0; jj < JP; ++jj) {
                    const int j = jg * JP + jj;
                    if (j < FB) {
                        dmma884(cr[j][0], cr[j][1], an[kk], f[jj][0][kk]);
                        dmma884(ci[j][0], ci[j][1], ar[kk], f[jj][1][kk]);
                        dmma884(cs[j][0], cs[j][1], ai[kk], f[jj][2][kk]);
                    }
                }
#pragma unroll
            for (int jj = 0; jj < JP; ++jj)
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    f[jj][pl][0] = gn[jj][pl][0];
                    f[jj][pl][1] = gn[jj][pl][1];
                }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < FB; ++j) s += cr[j][0] + cr[j][1] + ci[j][0] + ci[j][1] + cs[j][0] + cs[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int FB, int NL, int JP, int BAR>
static void run(int warps_per_smsp, double* out, long long* clk) {
    const int iters = 4000, threads = warps_per_smsp * 4 * 32;
    cudaFuncSetAttribute(probe<FB, NL, JP, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int r = 0; r < 2; ++r) {
        probe<FB, NL, JP, BAR><<<148, threads, 64 * 1024>>>(out, clk, iters);
        cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double per_warp = (double)iters * 6 * FB;
    printf("FB %2d  B loads per fragment %d  group %d  barrier ops %d  warps/smsp %d : %.2f clk per DMMA per sub-partition\n", FB, NL, JP,
           BAR, warps_per_smsp, avg / (per_warp * warps_per_smsp));
}

int main() {
    double* out;
    long long* clk;
    cudaMalloc(&out, 148 * 384 * 8);
    cudaMalloc(&clk, 148 * 8);
    for (int w = 1; w <= 3; ++w) {
        run<10, 3, 1, 0>(w, out, clk);
        run<10, 3, 2, 0>(w, out, clk);
        run<10, 3, 2, 1>(w, out, clk);
        run<10, 3, 2, 2>(w, out, clk);
        run<10, 3, 1, 2>(w, out, clk);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
