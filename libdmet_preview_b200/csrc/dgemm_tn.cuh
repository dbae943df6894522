// Real-FP64 "TN" GEMM / SYRK on the FP64 tensor cores (DMMA.8x8x4) with TMA-staged, 128B-swizzled tiles.
//
//   C[r, c] (+)= alpha * sum_{k < K} A[r, k] * B[c, k]          A: (M, lda) row-major, B: (N, ldb) row-major
//
// This is stage 3 of the embedding-ERI build: eri[P, Q] += w * sum_L X[L, P] X[L, Q] with X = Re / Im of the
// packed 3-index tensor (reference: `_Lij_s4_to_eri`, libdmet/basis_transform/eri_transform.py:436-485, which
// calls PySCF lib.dot = dgemm).  The caller stores X transposed and K-concatenated, XT[P][(kL, re|im, L)], so one
// launch covers several transfer momenta and both the real and the imaginary Gram product.
// With `lower_only` the kernel skips tiles strictly above the diagonal (syrk); a mirror kernel fills the upper
// triangle once at the end because the reference returns both triangles.
//
// Same skeleton as zgemm_tn.cuh: persistent CTAs, 8 consumer warps (4 x 2, warp tile 32 x 64) with the
// accumulators in registers, a producer warpgroup whose elected lane issues TMA, STAGES-deep mbarrier ring.
// One stage = 16 doubles of k: A tile 128 rows x 128 B and B tile 128 rows x 128 B, both 128-byte swizzled.
// Fragment loads are LDS.64; MMA row g reads tile row perm2(g) = ((g&3)<<1)|(g>>2), so the four rows a
// half-warp touches have distinct (row>>1)&3 and the swizzle spreads them over all 16 bank pairs.
#pragma once
#include "common.cuh"

namespace ldm {

struct DGemmArgs {
    int M, N, K;
    double* C;
    long long ldc;
    double alpha;
    int accumulate;
    int lower_only;      // 1: only tiles with tn <= tm (requires BM == BN)
    int tiles_m, tiles_n;
};

// Tile shapes: 128 x 128 (warp tile 32 x 64) for the large Gram products, 64 x 64 (warp tile 16 x 32) when the
// product has too few 128 x 128 tiles to occupy the SMs (npair of a few hundred to ~2000: BASELINE configs 1 - 3 and
// the lower end of the sweep).
template <int FA_, int FB_>
struct DTileT {
    static constexpr int WM = 4, WN = 2, FA = FA_, FB = FB_;
    static constexpr int BM = WM * FA * 8;
    static constexpr int BN = WN * FB * 8;
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCONS = WM * WN;
    static constexpr int THREADS = (NCONS + 4) * 32;
    static constexpr int STAGES = 6;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8;
};
using DTile = DTileT<4, 8>;       // 128 x 128
using DTileS = DTileT<2, 4>;      // 64 x 64

__device__ __forceinline__ void dsyrk_tile_coords(int lin, int lower_only, int tiles_n, int& tm, int& tn) {
    if (!lower_only) {
        tm = lin / tiles_n;
        tn = lin - tm * tiles_n;
    } else {   // lin enumerates (tm, tn <= tm) row by row: lin = tm(tm+1)/2 + tn
        int m = (int)((sqrt(8.0 * (double)lin + 1.0) - 1.0) * 0.5);
        while ((long long)(m + 1) * (m + 2) / 2 <= lin) ++m;
        while ((long long)m * (m + 1) / 2 > lin) --m;
        tm = m;
        tn = lin - m * (m + 1) / 2;
    }
}

template <class T>
__global__ void __launch_bounds__(T::THREADS, 1)
dgemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const DGemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + T::STAGES * T::STAGE_BYTES;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T::STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (T::STAGES + s), T::NCONS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int ktiles = (args.K + 15) >> 4;
    const int ntiles = args.lower_only ? args.tiles_m * (args.tiles_m + 1) / 2 : args.tiles_m * args.tiles_n;

    if (warp >= T::NCONS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == T::NCONS && lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int tm, tn;
                dsyrk_tile_coords(tile, args.lower_only, args.tiles_n, tm, tn);
                for (int kt = 0; kt < ktiles; ++kt) {
                    mbar_wait(bar_base + 8 * (T::STAGES + stage), phase ^ 1);
                    const uint32_t full = bar_base + 8 * stage;
                    const uint32_t dst = smem_base + stage * T::STAGE_BYTES;
                    mbar_expect_tx(full, T::STAGE_BYTES);
                    tma_load_3d(dst, &tmA, full, kt * 16, tm * T::BM, 0);
                    tma_load_3d(dst + T::A_BYTES, &tmB, full, kt * 16, tn * T::BN, 0);
                    if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = warp / T::WN, wn = warp % T::WN;
    const int g = lane >> 2, t = lane & 3;
    const int pg = ((g & 3) << 1) | (g >> 2);
    // element (row, k) of a tile lives at row*128 + ((k>>1) ^ (row&7))*16 + (k&1)*8
    const uint32_t a_row_off = (uint32_t)((wm * T::FA * 8 + pg) * 128);
    const uint32_t b_row_off = (uint32_t)(T::A_BYTES + (wn * T::FB * 8 + pg) * 128);
    const uint32_t sub = (uint32_t)((t & 1) * 8);
    const uint32_t ch = (uint32_t)(t >> 1);     // chunk of k = 4*kk + t is 2*kk + (t>>1)

    // A stage is handed back to the producer one iteration late, right after the wait for the NEXT stage: the
    // wait loop is a control-flow boundary behind all MMAs of the previous stage, so every fragment load of that
    // stage has returned its data (the MMAs that consume them have issued) before TMA may overwrite the buffer.
    // Releasing directly after the last fragment loads is not safe: the arrive does not wait for LDS in flight.
    int stage = 0, prev_stage = -1;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int tm, tn;
        dsyrk_tile_coords(tile, args.lower_only, args.tiles_n, tm, tn);
        double acc[T::FA][T::FB][2];
#pragma unroll
        for (int i = 0; i < T::FA; ++i)
#pragma unroll
            for (int j = 0; j < T::FB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < ktiles; ++kt) {
            mbar_wait(bar_base + 8 * stage, phase);
            if (prev_stage >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_base + 8 * (T::STAGES + prev_stage));
            }
            const uint32_t sbase = smem_base + stage * T::STAGE_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t koff = (((uint32_t)(2 * kk) + ch) ^ (uint32_t)pg) * 16 + sub;
                double a[T::FA];
#pragma unroll
                for (int i = 0; i < T::FA; ++i) a[i] = lds64(sbase + a_row_off + i * 1024 + koff);
#pragma unroll
                for (int j = 0; j < T::FB; ++j) {
                    const double b = lds64(sbase + b_row_off + j * 1024 + koff);
#pragma unroll
                    for (int i = 0; i < T::FA; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b);
                }
            }
            prev_stage = stage;
            if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
        }

        // epilogue.  Thread holds rows perm2(g), columns perm2(2t), perm2(2t+1) of every 8x8 fragment.
        const int pc0 = ((2 * t) & 3) << 1 | ((2 * t) >> 2);
        const int pc1 = ((2 * t + 1) & 3) << 1 | ((2 * t + 1) >> 2);
#pragma unroll
        for (int i = 0; i < T::FA; ++i) {
            const int r = tm * T::BM + wm * T::FA * 8 + i * 8 + pg;
            if (r >= args.M) continue;
            double* crow = args.C + (long long)r * args.ldc;
            // accumulate mode: all old values of the row first (16 independent loads), then the stores
            double old[T::FB][2];
            if (args.accumulate) {
#pragma unroll
                for (int j = 0; j < T::FB; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = tn * T::BN + wn * T::FB * 8 + j * 8 + (e ? pc1 : pc0);
                        old[j][e] = c < args.N ? crow[c] : 0.0;
                    }
            }
#pragma unroll
            for (int j = 0; j < T::FB; ++j) {
                const int cbase = tn * T::BN + wn * T::FB * 8 + j * 8;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = cbase + (e ? pc1 : pc0);
                    if (c >= args.N) continue;
                    double v = args.alpha * acc[i][j][e];
                    if (args.accumulate) v += old[j][e];
                    crow[c] = v;
                }
            }
        }
    }
}

}  // namespace ldm
