// DMMA.8x8x4 issue-rate probe (development aid): clocks per DMMA of one SM sub-partition as a function of the number
// of resident warps per sub-partition and of the distance (in DMMAs of the same warp) between two DMMAs that
// accumulate into the same registers.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int DIST>
__global__ void probe(double* out, long long* clk, int iters, double a, double b) {
    double c[DIST][2];
#pragma unroll
    for (int d = 0; d < DIST; ++d) c[d][0] = c[d][1] = threadIdx.x * 1e-3 + d;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 48 / DIST; ++rep)
#pragma unroll
            for (int d = 0; d < DIST; ++d) dmma884(c[d][0], c[d][1], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DIST; ++d) s += c[d][0] + c[d][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int DIST>
static void run(int warps_per_smsp, double* out, long long* clk) {
    const int iters = 2000, threads = warps_per_smsp * 4 * 32;
    probe<DIST><<<148, threads>>>(out, clk, iters, 1.0, 1e-9);
    cudaDeviceSynchronize();
    probe<DIST><<<148, threads>>>(out, clk, iters, 1.0, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double per_warp = (double)iters * (48 / DIST) * DIST;
    printf("dist %2d warps/smsp %d : %.2f clk per DMMA per sub-partition (%.2f per warp)\n", DIST, warps_per_smsp,
           avg / (per_warp * warps_per_smsp), avg / per_warp);
}

int main() {
    double* out;
    long long* clk;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&clk, 148 * 8);
    for (int w = 1; w <= 4; w *= 2) {
        run<1>(w, out, clk);
        run<2>(w, out, clk);
        run<3>(w, out, clk);
        run<4>(w, out, clk);
        run<6>(w, out, clk);
        run<8>(w, out, clk);
        run<12>(w, out, clk);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
