// Complex-FP64 "TN" GEMM on the FP64 tensor cores (DMMA.8x8x4 via mma.sync) with TMA-staged operand tiles.
//
//   C[b][r, c] (+)= alpha * sum_{s < nseg} sum_{k < K}  opA(A[az(b,s)][r, k]) * opB(B[bz(b,s)][c, k])
//
// Both operands are row-major with the contraction index k contiguous ("TN"): A is (Z_A, M, K), B is (Z_B, N, K),
// complex128 interleaved.  The contraction may be chained over `nseg` (A-slice, B-slice) segments per output
// tile -- this is how the sum over (k_i, k_j) blocks of the reference's inner loop
// (libdmet/basis_transform/eri_transform.py:344-378) is folded into one K loop.  opA/opB are optional complex
// conjugations, applied by flipping sign bits of the imaginary fragments.
//
// Replaces PySCF `_ao2mo.r_e2` (two zgemm per auxiliary row; call site eri_transform.py:432-433) and the
// numpy.dot loops of make_basis.py:548-557 / slater_helper.py:46.
//
// Structure: persistent CTAs (one per SM); warps 0..WM*WN-1 (two warpgroups) are MMA consumers holding the
// accumulators in registers (setmaxnreg.inc to 232), the third warpgroup gives its registers away and one
// elected lane of it is the TMA producer.  A STAGES-deep ring of smem stages is handed over with full/empty mbarriers.
// One stage = 8 complex k:  A tile  BM rows x 128 B, 128-byte swizzled by TMA;
//                           B tile  2 slabs (k 0-3, 4-7) of BN rows x 64 B, unswizzled.
// Fragment loads are LDS.128 (re, im) and bank-conflict free by construction:
//   A: MMA row g of a fragment reads tile row perm(g) = (g>>1)|((g&1)<<2), so the two rows a quarter-warp
//      touches differ in bit 2 and the swizzle sends them to disjoint bank groups;
//   B: rows 2q, 2q+1 of a 64-byte slab are one 128-byte line.
//
// M3 = true selects the 3-multiplication form of the complex product (Gauss / "3M"):
//   k1 = (Ar + Ai) Br,  k2 = Ar (Bi - Br),  k3 = Ai (Br + Bi);   Re = k1 - k3,  Im = k1 + k2
// i.e. three real DMMAs per complex fragment pair instead of four and a third accumulator set (so the 3M tiles are
// narrower: FB <= 13).  The kernel is DMMA-bound, so this is a 4/3 reduction of the executed tensor work; the result
// differs from the 4-multiplication form by rounding only (normwise bound of the same order, |error| ~ 1e-15
// relative to |A||B| here against the 1e-10 parity bar).  All linear forms of B are precomputed in memory
// (`zforms_kernel`: five real planes Br, D = Bi - Br, S = Br + Bi, -S, -D per slice; conj(B) uses Br, -S, -D), so
// the producer just picks three planes per segment and the only extra arithmetic in the loop is one DADD per A
// fragment -- DADDs share the FP64 pipe with the DMMAs and each one costs tensor issue slots.
// 3M stage: A tile as above; B = 3 planes of BN rows x 64 B (8 real k in the order 0,4,1,5,2,6,3,7), 64-byte
// swizzled by TMA; one LDS.128 per plane and column fragment delivers the words of both k-steps of the stage.
// JP = column fragments per group of the 3M main loop: within a group the k-step loop runs outside the fragment loop,
// so with JP = 2 (the default of the wide tiles) two DMMAs on the same accumulator are six instructions apart in the
// source instead of three.  The pipe takes one DMMA per 16 clocks and sub-partition, but a lone warp that issues two
// DMMAs on one accumulator back to back waits 26 (tools/dmma_probe.cu) -- the order ptxas otherwise derives.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace ldm {

struct ZSeg {          // one K-segment of one batch entry
    int az;            // slice of A (3rd tensor-map coordinate)
    int bz;            // slice of B
    uint32_t conjA;    // 0 or 0x80000000
    uint32_t conjB;    // 0 or 0x80000000
};

struct ZGemmArgs {
    int M, N, K;                 // per segment, complex elements
    int nseg, nbatch;
    const ZSeg* segs;            // [nbatch * nseg], device
    double2* C;
    const long long* c_off;      // [nbatch] element offsets into C, device (may be null -> 0)
    // output addressing (complex elements): off = (r / rdiv) * s_outer + (r % rdiv) * s_inner + c * s_col
    int rdiv;
    long long s_outer, s_inner, s_col;
    double alpha;
    int accumulate;              // 1: C += result
    int tiles_m, tiles_n;
    int short_last;              // 3M: the LAST n-tile runs with FB - 1 column fragments (its last fragment would be all
                                 // padding: N = 150 as 80 + 72 instead of 80 + 80); the n index is then rotated by the
                                 // wave number so that every CTA gets long and short tiles in turn
};

// tile index -> (batch entry, m-tile, n-tile).  n runs fastest, so the tiles_n CTAs that share an A tile run side by
// side and the second reader finds it in L2; with `short_last` the n index is shifted by the wave number (the grid
// is a multiple of tiles_n then, so the tiles of one A tile stay in one wave and the map stays one-to-one).
__device__ __forceinline__ void tile_coords(const ZGemmArgs& a, int tile, int tiles_per_batch, int& b, int& tm,
                                            int& tn) {
    b = tile / tiles_per_batch;
    const int rem = tile - b * tiles_per_batch;
    tm = rem / a.tiles_n;
    tn = rem - tm * a.tiles_n;
    if (a.short_last) tn = (tn + tile / (int)gridDim.x) % a.tiles_n;
}

constexpr int ZFORM_PLANES = 5;      // Br, D, S, -S, -D

template <int WM, int WN, int FA, int FB, bool M3 = false>
struct ZTile {
    static constexpr int BM = WM * FA * 8;
    static constexpr int BN = WN * FB * 8;
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_SLAB = BN * 64;
    static constexpr int TX_BYTES = A_BYTES + (M3 ? 3 : 2) * B_SLAB;      // what TMA delivers per stage
    static constexpr int STAGE_BYTES = (TX_BYTES + 1023) / 1024 * 1024;     // stages stay 1024-byte aligned (swizzle)
    static constexpr int NCONS = WM * WN;
    static constexpr int THREADS = (NCONS + 4) * 32;   // + one producer warpgroup
    // Staged epilogue (3M tiles): the consumers park a finished tile in shared memory (row stride BN + 1 complex:
    // conflict-free both along rows and along columns) and three otherwise idle warps of the producer warpgroup write it
    // to global memory while the consumers are already in the next tile's main loop.
    static constexpr int STG_LD = BN + 1;
    static constexpr int STG_BYTES_WANTED = BM * STG_LD * 16;
    static constexpr int SMEM_BUDGET = 224 * 1024;
    static constexpr bool STG = M3 && BM == 64 && (4 * STAGE_BYTES + STG_BYTES_WANTED + 2048 <= SMEM_BUDGET);
    static constexpr int STG_BYTES = STG ? STG_BYTES_WANTED : 0;
    static constexpr int STAGES_FIT = STG ? (SMEM_BUDGET - STG_BYTES - 2048) / STAGE_BYTES : 200 * 1024 / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int STG_WARPS = 3;
    static constexpr int SMEM = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align*/ + (2 * STAGES + 2) * 8;
};

template <int WM, int WN, int FA, int FB, bool M3, int JP = 1>
__global__ void __launch_bounds__((WM * WN + 4) * 32, 1)
zgemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const ZGemmArgs args) {
    using T = ZTile<WM, WN, FA, FB, M3>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stg_base = smem_base + T::STAGES * T::STAGE_BYTES;   // staging tile of the epilogue (if any)
    const uint32_t bar_base = stg_base + T::STG_BYTES;                  // full[s] at +8s, empty[s] at +8(STAGES+s)
    const uint32_t stg_full = bar_base + 8 * (2 * T::STAGES), stg_empty = stg_full + 8;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T::STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);
            mbar_init(bar_base + 8 * (T::STAGES + s), T::NCONS);
        }
        mbar_init(stg_full, T::NCONS * 32);          // every thread that writes the staging tile arrives itself
        mbar_init(stg_empty, T::STG_WARPS * 32);     // ... and so does every thread that reads it
        mbar_fence_init();
    }
    __syncthreads();

    const int ktiles = (args.K + 7) >> 3;
    const int tiles_per_batch = args.tiles_m * args.tiles_n;
    const int ntiles = tiles_per_batch * args.nbatch;

    if (warp >= T::NCONS) {
        // ===================== TMA producer warpgroup (one elected lane works) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == T::NCONS && lane == 0) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int b, tm, tn;
                tile_coords(args, tile, tiles_per_batch, b, tm, tn);
                const ZSeg* segs = args.segs + (size_t)b * args.nseg;
                for (int s = 0; s < args.nseg; ++s) {
                    const int az = segs[s].az, bz = segs[s].bz;
                    // 3M: planes (Br, D, S) of slice bz, or (Br, -S, -D) for conj(B)
                    const int pl0 = ZFORM_PLANES * bz, pl1 = pl0 + (segs[s].conjB ? 3 : 1), pl2 = pl1 + 1;
                    for (int kt = 0; kt < ktiles; ++kt) {
                        mbar_wait(bar_base + 8 * (T::STAGES + stage), phase ^ 1);
                        const uint32_t full = bar_base + 8 * stage;
                        const uint32_t dst = smem_base + stage * T::STAGE_BYTES;
                        mbar_expect_tx(full, T::TX_BYTES);
                        tma_load_3d(dst, &tmA, full, kt * 16, tm * T::BM, az);
                        if constexpr (M3) {
                            tma_load_3d(dst + T::A_BYTES, &tmB, full, kt * 8, tn * T::BN, pl0);
                            tma_load_3d(dst + T::A_BYTES + T::B_SLAB, &tmB, full, kt * 8, tn * T::BN, pl1);
                            tma_load_3d(dst + T::A_BYTES + 2 * T::B_SLAB, &tmB, full, kt * 8, tn * T::BN, pl2);
                        } else {
                            tma_load_3d(dst + T::A_BYTES, &tmB, full, kt * 16, tn * T::BN, bz);
                            tma_load_3d(dst + T::A_BYTES + T::B_SLAB, &tmB, full, kt * 16 + 8, tn * T::BN, bz);
                        }
                        if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        if constexpr (T::STG) {
            if (warp > T::NCONS) {
                // ===================== epilogue store warps =====================
                // tile by tile: wait until the consumers have parked the tile, write it out (read-modify-write in
                // accumulate mode) with the lanes along the contiguous direction of the output, hand the buffer back
                const int sw = warp - T::NCONS - 1;
                const double2* stg = reinterpret_cast<const double2*>(smem_raw + (stg_base - smem_u32(smem_raw)));
                uint32_t sphase = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    int b, tm, tn;
                    tile_coords(args, tile, tiles_per_batch, b, tm, tn);
                    double2* Cb = args.C + (args.c_off ? args.c_off[b] : 0ll);
                    const int rows = min(T::BM, args.M - tm * T::BM);
                    const int cols = min(T::BN, args.N - tn * T::BN);
                    const long long sc = args.s_col;
                    mbar_wait(stg_full, sphase);
                    if (sc == 1) {
                        // columns are contiguous in the output: a warp per row, lanes along the columns
                        for (int r = sw; r < rows; r += T::STG_WARPS) {
                            const int rg = tm * T::BM + r;
                            double2* dst = Cb + (long long)(rg / args.rdiv) * args.s_outer +
                                           (long long)(rg % args.rdiv) * args.s_inner + (long long)tn * T::BN;
                            const double2* src = stg + r * T::STG_LD;
                            for (int c = lane; c < cols; c += 32) {
                                double2 v = src[c];
                                if (args.accumulate) {
                                    const double2 o = dst[c];
                                    v.x += o.x;
                                    v.y += o.y;
                                }
                                dst[c] = v;
                            }
                        }
                    } else {
                        // rows are (piecewise) contiguous: a warp per column, lanes along the 64 rows
                        long long roff[2];
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int rg = tm * T::BM + lane + 32 * h2;
                            roff[h2] = (long long)(rg / args.rdiv) * args.s_outer + (long long)(rg % args.rdiv) * args.s_inner;
                        }
                        for (int c = sw; c < cols; c += T::STG_WARPS) {
                            double2* dst = Cb + (long long)(tn * T::BN + c) * sc;
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2) {
                                const int r = lane + 32 * h2;
                                if (r < rows) {
                                    double2 v = stg[r * T::STG_LD + c];
                                    if (args.accumulate) {
                                        const double2 o = dst[roff[h2]];
                                        v.x += o.x;
                                        v.y += o.y;
                                    }
                                    dst[roff[h2]] = v;
                                }
                            }
                        }
                    }
                    mbar_arrive(stg_empty);              // release: this thread's reads of the buffer are done
                    sphase ^= 1;
                }
            }
        }
        return;
    }

    // ===================== MMA consumers =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) | ((g & 1) << 2);
    // per-thread smem offsets inside a stage
    const uint32_t a_row_off = (uint32_t)((wm * FA * 8 + pg) * 128);
    const uint32_t a_c0 = (uint32_t)((t ^ pg) * 16);       // k-step 0 ; k-step 1 is a_c0 ^ 64
    // 4M: (re, im) of k = 4 kk + t in slab kk.
    // 3M: the planes are stored with the 8 k of a stage in the order 0,4,1,5,2,6,3,7 (zforms_kernel), so the 16-byte
    // chunk t of a 64-byte-swizzled row (chunk index xor ((row >> 1) & 3)) holds the words k = t and k = t + 4: ONE
    // 128-bit load per plane and fragment serves both k-steps of the stage (half the shared-load instructions of a
    // load per k-step -- every instruction a consumer warp issues between its DMMAs costs tensor-pipe time).  A
    // quarter-warp covers tile rows 2r, 2r + 1 = 128 contiguous bytes: conflict free.  MMA column g reads tile row
    // permB(g) = g with bits 1 and 2 swapped (kept from the 64-bit version; the epilogue applies the same
    // permutation to the output column).
    const int pgb = (g & 1) | ((g & 2) << 1) | ((g & 4) >> 1);
    const uint32_t b_off = M3 ? (uint32_t)(T::A_BYTES + (wn * FB * 8 + pgb) * 64 + ((t ^ ((pgb >> 1) & 3)) << 4))
                              : (uint32_t)(T::A_BYTES + (wn * FB * 8 + g) * 64 + t * 16);

    // A stage is handed back to the producer one iteration late, right after the wait for the NEXT stage: the
    // wait loop is a control-flow boundary behind all MMAs of the previous stage, so every fragment load of that
    // stage has returned its data (the MMAs that consume them have issued) before TMA may overwrite the buffer.
    // Releasing directly after the last fragment loads is not safe: the arrive does not wait for LDS in flight.
    int stage = 0, prev_stage = -1;
    uint32_t phase = 0, cphase = 0;
    // One tile, with FBE <= FB column fragments per warp (a compile-time count: the short last n-tile is a second
    // instance of the same code, chosen once per tile -- a run-time test per fragment costs more than the skipped
    // MMAs save, measured).
    auto run_tile = [&](auto fbe_tag, const int b, const int tm, const int tn) {
        constexpr int FBE = decltype(fbe_tag)::value;
        const ZSeg* segs = args.segs + (size_t)b * args.nseg;

        // 4M: cr = Re, ci = Im.   3M: cr = k1, ci = k2, cs = k3.
        double cr[FA][FBE][2], ci[FA][FBE][2], cs[M3 ? FA : 1][M3 ? FBE : 1][2];
#pragma unroll
        for (int i = 0; i < FA; ++i)
#pragma unroll
            for (int j = 0; j < FBE; ++j) {
                cr[i][j][0] = cr[i][j][1] = 0.0;
                ci[i][j][0] = ci[i][j][1] = 0.0;
                if constexpr (M3) cs[i][j][0] = cs[i][j][1] = 0.0;
            }

        for (int s = 0; s < args.nseg; ++s) {
            const uint32_t mA = segs[s].conjA;   // sa = -1 -> flip
            const uint32_t mB = segs[s].conjB;
            for (int kt = 0; kt < ktiles; ++kt) {
                mbar_wait(bar_base + 8 * stage, phase);
                if (prev_stage >= 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_base + 8 * (T::STAGES + prev_stage));
                }
                const uint32_t sbase = smem_base + stage * T::STAGE_BYTES;
                if constexpr (M3) {
                    // both k-steps of the stage: A fragments (re, im) of k = t and k = t + 4, then per column fragment
                    // three 128-bit loads (one per plane) and six DMMAs.  Fragment j + 1 is in flight while the
                    // MMAs of j issue.
                    double ar[2][FA], aip[2][FA], ain[2][FA];
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint32_t a_addr = sbase + a_row_off + (a_c0 ^ (kk * 64));
#pragma unroll
                        for (int i = 0; i < FA; ++i) {
                            double x, y;
                            lds128(a_addr + i * 1024, x, y);
                            ar[kk][i] = x;
                            aip[kk][i] = xor_hi(y, mA);               //  sa * Ai
                            ain[kk][i] = x + aip[kk][i];              // Ar + sa * Ai
                        }
                    }
                    const uint32_t b_addr = sbase + b_off;
                    // column fragments are processed in groups of JP: within a group the k-step loop is outside the
                    // fragment loop, so two DMMAs on the same accumulator are 3 * JP instructions apart
                    constexpr int NG = (FBE + JP - 1) / JP;
                    double f[JP][3][2], gn[JP][3][2];
#pragma unroll
                    for (int jj = 0; jj < JP; ++jj)
#pragma unroll
                        for (int pl = 0; pl < 3; ++pl) {
                            gn[jj][pl][0] = gn[jj][pl][1] = 0.0;
                            if (jj < FBE) lds128(b_addr + jj * 512 + pl * T::B_SLAB, f[jj][pl][0], f[jj][pl][1]);
                        }
#pragma unroll
                    for (int jg = 0; jg < NG; ++jg) {
#pragma unroll
                        for (int jj = 0; jj < JP; ++jj)
#pragma unroll
                            for (int pl = 0; pl < 3; ++pl)
                                if ((jg + 1) * JP + jj < FBE)
                                    lds128(b_addr + ((jg + 1) * JP + jj) * 512 + pl * T::B_SLAB, gn[jj][pl][0],
                                           gn[jj][pl][1]);
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                            for (int jj = 0; jj < JP; ++jj) {
                                const int j = jg * JP + jj;
                                if (j < FBE) {
#pragma unroll
                                    for (int i = 0; i < FA; ++i) {
                                        dmma884(cr[i][j][0], cr[i][j][1], ain[kk][i], f[jj][0][kk]);   // k1 = (Ar + Ai) Br
                                        dmma884(ci[i][j][0], ci[i][j][1], ar[kk][i], f[jj][1][kk]);    // k2 = Ar (Bi - Br)
                                        dmma884(cs[i][j][0], cs[i][j][1], aip[kk][i], f[jj][2][kk]);   // k3 = Ai (Br + Bi)
                                    }
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
                } else {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        double ar[FA], aip[FA], ain[FA];
                        const uint32_t a_addr = sbase + a_row_off + (a_c0 ^ (kk * 64));
#pragma unroll
                        for (int i = 0; i < FA; ++i) {
                            double x, y;
                            lds128(a_addr + i * 1024, x, y);
                            ar[i] = x;
                            aip[i] = xor_hi(y, mA);                   //  sa * Ai
                            ain[i] = xor_hi(y, mA ^ 0x80000000u);     // -sa * Ai
                        }
                        const uint32_t b_addr = sbase + b_off + kk * T::B_SLAB;
                        // software pipeline over the B fragments: fragment j+1 is in flight while the MMAs of j issue
                        double br, bi, br_n = 0.0, bi_n = 0.0;
                        lds128(b_addr, br, bi);
#pragma unroll
                        for (int j = 0; j < FBE; ++j) {
                            if (j + 1 < FBE) lds128(b_addr + (j + 1) * 512, br_n, bi_n);
                            bi = xor_hi(bi, mB);                  //  sb * Bi
#pragma unroll
                            for (int i = 0; i < FA; ++i) {
                                dmma884(cr[i][j][0], cr[i][j][1], ar[i], br);
                                dmma884(ci[i][j][0], ci[i][j][1], ar[i], bi);
                                dmma884(cr[i][j][0], cr[i][j][1], ain[i], bi);
                                dmma884(ci[i][j][0], ci[i][j][1], aip[i], br);
                            }
                            br = br_n;
                            bi = bi_n;
                        }
                    }
                }
                prev_stage = stage;
                if (++stage == T::STAGES) { stage = 0; phase ^= 1; }
            }
        }

        // ---------------- epilogue ----------------
        if constexpr (T::STG) {
            // park the tile (alpha and the 3M combination applied) in the staging buffer and go on; the store
            // warps write it out.  The buffer of the previous tile must have been drained.
            double2* stg = reinterpret_cast<double2*>(smem_raw + (stg_base - smem_u32(smem_raw)));
            mbar_wait(stg_empty, cphase ^ 1);
            const double alpha = args.alpha;
#pragma unroll
            for (int i = 0; i < FA; ++i) {
                double2* row = stg + (wm * FA * 8 + i * 8 + pg) * T::STG_LD + wn * FB * 8;
#pragma unroll
                for (int j = 0; j < FBE; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int q = 2 * t + e;
                        const int c = j * 8 + ((q & 1) | ((q & 2) << 1) | ((q & 4) >> 1));
                        row[c] = make_double2(alpha * (cr[i][j][e] - cs[i][j][e]), alpha * (cr[i][j][e] + ci[i][j][e]));
                    }
            }
            mbar_arrive(stg_full);                   // release: this thread's part of the tile is in place
            cphase ^= 1;
            return;
        }
        // ---------------- registers -> global ----------------
        // accumulate mode reads the old values of a whole chunk first (independent loads in flight together) and
        // only then stores: a load/store chain per element would expose one DRAM round trip per element.
        double2* Cb = args.C + (args.c_off ? args.c_off[b] : 0ll);
        const double alpha = args.alpha;
        auto col_in_frag = [](int q) { return M3 ? ((q & 1) | ((q & 2) << 1) | ((q & 4) >> 1)) : q; };
        constexpr int JB = FB > 23 ? 1 : (FB > 20 ? 2 : 4);   // the widest tiles have no registers to spare
        constexpr bool kBatch = FB <= 23;
        // addresses: one division per row, then pointer steps only (the epilogue is ~5 % of a consumer warp's time
        // at K = 200 and was dominated by 64-bit index arithmetic per stored element)
        const long long sc = args.s_col, step = 8 * sc;
        const int c_base = tn * T::BN + wn * FB * 8;
        const int ce0 = c_base + col_in_frag(2 * t), ce1 = c_base + col_in_frag(2 * t + 1);
#pragma unroll
        for (int i = 0; i < FA; ++i) {
            const int r = tm * T::BM + wm * FA * 8 + i * 8 + pg;
            if (r >= args.M) continue;
            const long long roff = (long long)(r / args.rdiv) * args.s_outer + (long long)(r % args.rdiv) * args.s_inner;
            double2* const pe0 = Cb + roff + (long long)ce0 * sc;
            double2* const pe1 = Cb + roff + (long long)ce1 * sc;
#pragma unroll
            for (int j0 = 0; j0 < FBE; j0 += JB) {
                double2 old[JB][2];
                if (kBatch && args.accumulate) {
#pragma unroll
                    for (int jj = 0; jj < JB; ++jj)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = j0 + jj;
                            const int c = (e ? ce1 : ce0) + 8 * j;
                            const double2* src = (e ? pe1 : pe0) + (long long)j * step;
                            old[jj][e] = (j < FBE && c < args.N) ? *src : make_double2(0.0, 0.0);
                        }
                }
#pragma unroll
                for (int jj = 0; jj < JB; ++jj) {
                    if (j0 + jj >= FBE) continue;
                    const int j = j0 + jj;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = (e ? ce1 : ce0) + 8 * j;
                        if (c >= args.N) continue;
                        double2* dst = (e ? pe1 : pe0) + (long long)j * step;
                        double2 v;
                        if constexpr (M3)
                            v = make_double2(alpha * (cr[i][j][e] - cs[i][j][e]),
                                             alpha * (cr[i][j][e] + ci[i][j][e]));
                        else
                            v = make_double2(alpha * cr[i][j][e], alpha * ci[i][j][e]);
                        if (args.accumulate) {
                            const double2 o = kBatch ? old[jj][e] : *dst;
                            v.x += o.x;
                            v.y += o.y;
                        }
                        *dst = v;
                    }
                }
            }
        }
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int b, tm, tn;
        tile_coords(args, tile, tiles_per_batch, b, tm, tn);
        if constexpr (M3 && FB > 1) {
            if (args.short_last && tn == args.tiles_n - 1) {
                run_tile(std::integral_constant<int, FB - 1>{}, b, tm, tn);
                continue;
            }
        }
        run_tile(std::integral_constant<int, FB>{}, b, tm, tn);
    }
}

}  // namespace ldm
