// HBM-bound kernels around the two tensor-core GEMMs: lattice phase transforms (k2R / R2k), batched transposes,
// symmetrise + pack + transpose of the 3-index tensor, s4 -> s1 / s8 re-layouts, triangle mirror, J/K contraction,
// k-point reduction and the synthetic GDF generator.  All are coalesced, grid-sized streaming kernels; the
// reference routines each one replaces are cited per kernel (paths relative to /root/reference).
#pragma once
#include "common.cuh"

namespace ldm {

// max over the warp, then one atomicMax per warp -- and only when the value beats what is already there: a plain
// (L2) read filters almost every warp once the running maximum has settled, so a single address is not hammered
// by one atomic per warp of a 100k-warp grid (non-negative doubles order like their bit patterns).
__device__ __forceinline__ void warp_atomic_max_abs(double v, unsigned long long* addr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0.0) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        if (bits > *reinterpret_cast<volatile unsigned long long*>(addr)) atomicMax(addr, bits);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Lattice Fourier transform as a dense phase-matrix product (Nk <= a few hundred):
//     out[b][k][x] = scale * sum_R W[k][R] * in[b][R][x]
// Replaces scipy fftn/ifftn in libdmet/system/fourier.py:160-177 (FFTtoK / FFTtoT) and the einsum of
// eri_transform.py:125 (get_basis_k).  in may be real (in_real) and out may keep only the real part (out_real,
// k2R returns .real, fourier.py:157); max|imag| of the discarded part is accumulated into *imag_max
// (the reference warns when it exceeds IMAG_DISCARD_TOL, fourier.py:174-175).
// Each thread owns one x and KT output k's; the phase matrix sits in shared memory.
// ----------------------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(128)
phase_transform_kernel(const double* __restrict__ in, double* __restrict__ out, const double2* __restrict__ W,
                       int nin, int nout, long long X, double scale, int in_real, int out_real,
                       unsigned long long* imag_max) {
    extern __shared__ double2 sW[];   // [KT][nin]
    const int k0 = blockIdx.y * KT;
    for (int idx = threadIdx.x; idx < KT * nin; idx += blockDim.x) {
        const int kk = idx / nin, r = idx - kk * nin;
        sW[idx] = (k0 + kk < nout) ? W[(size_t)(k0 + kk) * nin + r] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t bin = (size_t)blockIdx.z * nin * X;
    const size_t bout = (size_t)blockIdx.z * nout * X;
    double accr[KT], acci[KT];
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) accr[kk] = acci[kk] = 0.0;
    if (x < X) {
        for (int r = 0; r < nin; ++r) {
            double vr, vi;
            if (in_real) {
                vr = in[bin + (size_t)r * X + x];
                vi = 0.0;
            } else {
                const double2 v = reinterpret_cast<const double2*>(in)[bin + (size_t)r * X + x];
                vr = v.x;
                vi = v.y;
            }
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                const double2 w = sW[kk * nin + r];
                accr[kk] = fma(w.x, vr, accr[kk]);
                accr[kk] = fma(-w.y, vi, accr[kk]);
                acci[kk] = fma(w.x, vi, acci[kk]);
                acci[kk] = fma(w.y, vr, acci[kk]);
            }
        }
    }
    double im_max = 0.0;
    if (x < X) {
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            if (k0 + kk >= nout) break;
            const double re = accr[kk] * scale, im = acci[kk] * scale;
            if (out_real) {
                out[bout + (size_t)(k0 + kk) * X + x] = re;
                im_max = fmax(im_max, fabs(im));
            } else {
                reinterpret_cast<double2*>(out)[bout + (size_t)(k0 + kk) * X + x] = make_double2(re, im);
            }
        }
    }
    if (out_real && imag_max) warp_atomic_max_abs(im_max, imag_max);
}

// ----------------------------------------------------------------------------------------------------------
// Lattice Fourier transform on the k-mesh itself, factorised over the (<= 3) mesh axes:
//     out[b][k][x] = scale * sum_R exp(-+ 2 pi i k.R) in[b][R][x]        (R2k: minus, k2R: plus and scale = 1/Nk)
// One CTA stages a [Nk][TX] tile of TX consecutive x in shared memory (512-byte coalesced rows), runs one dense
// n_d-point DFT pass per mesh axis in place (each thread owns whole lines, one __syncthreads per pass) and streams
// the result out (the first pass reads straight from global memory, the last one writes straight to it).
// 12 complex MACs per element at 4x4x4 instead of 64 for the dense phase matrix, so the kernel is
// bound by HBM: algorithmic bytes 8|16 Nk X in + 8|16 Nk X out.  Replaces scipy fftn / ifftn
// (libdmet/system/fourier.py:160-177).  Mesh axes up to 8 points; larger meshes use phase_transform_kernel.
// ----------------------------------------------------------------------------------------------------------
struct DftTables {
    double2 w[3][64];     // w[d][k * n_d + r] = exp(-+ 2 pi i k r / n_d), filled on the host
};

// one n-point DFT over a line held in registers.  Lines of 2 and 4 points (the common meshes: 2x2x2 ... 4x4x4) are
// butterflies of additions only -- their twiddles are 1, -1, -+i -- so the kernel stays bound by HBM also for the
// real-input / real-output variants, which move 24 instead of 32 bytes per element; other lengths are a dense sum.
template <int ND>
__device__ __forceinline__ void dft_line(double2 (&v)[ND], const double2* __restrict__ tw, double2 (&o)[ND]) {
    if constexpr (ND == 1) {
        o[0] = v[0];
    } else if constexpr (ND == 2) {
        o[0] = make_double2(v[0].x + v[1].x, v[0].y + v[1].y);
        o[1] = make_double2(v[0].x - v[1].x, v[0].y - v[1].y);
    } else if constexpr (ND == 4) {
        const double s = tw[ND + 1].y;            // w = exp(-+ 2 pi i / 4) = (0, s), s = -1 (R -> k) or +1 (k -> R)
        const double2 t0 = make_double2(v[0].x + v[2].x, v[0].y + v[2].y);
        const double2 t1 = make_double2(v[0].x - v[2].x, v[0].y - v[2].y);
        const double2 t2 = make_double2(v[1].x + v[3].x, v[1].y + v[3].y);
        const double2 t3 = make_double2(s * (v[3].y - v[1].y), s * (v[1].x - v[3].x));      // w * (v1 - v3)
        o[0] = make_double2(t0.x + t2.x, t0.y + t2.y);
        o[2] = make_double2(t0.x - t2.x, t0.y - t2.y);
        o[1] = make_double2(t1.x + t3.x, t1.y + t3.y);
        o[3] = make_double2(t1.x - t3.x, t1.y - t3.y);
    } else {
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int r = 0; r < ND; ++r) {
                const double2 w = tw[k * ND + r];
                re = fma(w.x, v[r].x, re);
                re = fma(-w.y, v[r].y, re);
                im = fma(w.x, v[r].y, im);
                im = fma(w.y, v[r].x, im);
            }
            o[k] = make_double2(re, im);
        }
    }
}

// pass over mesh axis 0: global -> registers -> shared tile
template <int ND>
__device__ __forceinline__ void dft_pass_in(const double* __restrict__ in, double2* tile, const double2* tw, int s0,
                                            int TX, long long X, long long x0, size_t boff, int in_real) {
    for (int ln = threadIdx.x; ln < s0 * TX; ln += blockDim.x) {
        const int tx = ln % TX, inner = ln / TX;
        double2 v[ND], o[ND];
        const bool ok = x0 + tx < X;
#pragma unroll
        for (int r = 0; r < ND; ++r) {
            v[r] = make_double2(0.0, 0.0);
            if (ok) {
                const size_t g = boff + (size_t)(r * s0 + inner) * X + x0 + tx;
                if (in_real) v[r].x = in[g];
                else v[r] = reinterpret_cast<const double2*>(in)[g];
            }
        }
        dft_line<ND>(v, tw, o);
#pragma unroll
        for (int k = 0; k < ND; ++k) tile[(k * s0 + inner) * TX + tx] = o[k];
    }
}

// pass over mesh axis 1: shared -> shared, in place
template <int ND>
__device__ __forceinline__ void dft_pass_mid(double2* tile, const double2* tw, int n0, int s1, int TX) {
    for (int ln = threadIdx.x; ln < n0 * s1 * TX; ln += blockDim.x) {
        const int tx = ln % TX, rest = ln / TX;
        const int outer = rest / s1, inner = rest - outer * s1;
        const int base = (outer * ND * s1 + inner) * TX + tx;
        double2 v[ND], o[ND];
#pragma unroll
        for (int r = 0; r < ND; ++r) v[r] = tile[base + r * s1 * TX];
        dft_line<ND>(v, tw, o);
#pragma unroll
        for (int k = 0; k < ND; ++k) tile[base + k * s1 * TX] = o[k];
    }
}

// pass over mesh axis 2: shared -> registers -> global
template <int ND>
__device__ __forceinline__ double dft_pass_out(double* __restrict__ out, const double2* tile, const double2* tw,
                                               int nouter, int TX, long long X, long long x0, size_t boff,
                                               double scale, int out_real) {
    double im_max = 0.0;
    for (int ln = threadIdx.x; ln < nouter * TX; ln += blockDim.x) {
        const int tx = ln % TX, outer = ln / TX;
        double2 v[ND], o[ND];
#pragma unroll
        for (int r = 0; r < ND; ++r) v[r] = tile[(outer * ND + r) * TX + tx];
        dft_line<ND>(v, tw, o);
        if (x0 + tx < X) {
#pragma unroll
            for (int k = 0; k < ND; ++k) {
                const size_t g = boff + (size_t)(outer * ND + k) * X + x0 + tx;
                if (out_real) {
                    out[g] = o[k].x * scale;
                    im_max = fmax(im_max, fabs(o[k].y * scale));
                } else {
                    reinterpret_cast<double2*>(out)[g] = make_double2(o[k].x * scale, o[k].y * scale);
                }
            }
        }
    }
    return im_max;
}

// dispatch on the axis length; MAXND bounds the instantiated line lengths so that the common small meshes (axes of
// <= 4 points) compile to <= 64 registers and four CTAs share an SM (load, compute and store phases overlap)
#define LDM_DFT_SWITCH(nd, CALL)                                              \
    switch (nd) {                                                            \
        case 1: { constexpr int ND = 1; CALL; } break;                       \
        case 2: { constexpr int ND = 2; CALL; } break;                       \
        case 3: { constexpr int ND = 3; CALL; } break;                       \
        case 4: { constexpr int ND = 4; CALL; } break;                       \
        default:                                                             \
            if constexpr (MAXND > 4) {                                       \
                switch (nd) {                                                \
                    case 5: { constexpr int ND = 5; CALL; } break;           \
                    case 6: { constexpr int ND = 6; CALL; } break;           \
                    case 7: { constexpr int ND = 7; CALL; } break;           \
                    default: { constexpr int ND = 8; CALL; } break;          \
                }                                                            \
            }                                                                \
            break;                                                           \
    }

template <int MAXND>
__global__ void __launch_bounds__(256, MAXND <= 4 ? 4 : 1)
lattice_dft_kernel(const double* __restrict__ in, double* __restrict__ out, const __grid_constant__ DftTables tb,
                   int n0, int n1, int n2, long long X, int TX, double scale, int in_real, int out_real,
                   unsigned long long* imag_max) {
    extern __shared__ double2 dft_smem[];
    const int nk = n0 * n1 * n2;
    double2* tile = dft_smem;                 // [nk][TX]
    const long long x0 = (long long)blockIdx.x * TX;
    const size_t boff = (size_t)blockIdx.y * nk * X;
    LDM_DFT_SWITCH(n0, dft_pass_in<ND>(in, tile, tb.w[0], n1 * n2, TX, X, x0, boff, in_real));
    __syncthreads();
    if (n1 > 1) {
        LDM_DFT_SWITCH(n1, dft_pass_mid<ND>(tile, tb.w[1], n0, n2, TX));
        __syncthreads();
    }
    double im_max = 0.0;
    LDM_DFT_SWITCH(n2, im_max = dft_pass_out<ND>(out, tile, tb.w[2], n0 * n1, TX, X, x0, boff, scale, out_real));
    if (out_real && imag_max) warp_atomic_max_abs(im_max, imag_max);
}

// ----------------------------------------------------------------------------------------------------------
// Batched complex transpose with optional conjugation and real scale:  out[b][c][r] = scale * op(in[b][r][c]).
// Used to lay coefficient matrices out with the contraction index contiguous for the TN GEMM.
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ztranspose_kernel(const double2* __restrict__ in, double2* __restrict__ out, int rows, int cols, int conj,
                  double scale) {
    __shared__ double2 tile[32][33];
    const size_t boff = (size_t)blockIdx.z * rows * cols;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = in[boff + (size_t)r * cols + c];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        const int c = c0 + dy, r = r0 + threadIdx.x;
        if (r < rows && c < cols) {
            double2 v = tile[threadIdx.x][dy];
            v.x *= scale;
            v.y *= conj ? -scale : scale;
            out[boff + (size_t)c * rows + r] = v;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// Stored cderi entry -> GDF block (PySCF _load3c + sr_loop semantics on the device; gdf_file.py).
//   src   (rows, ncols) as it lies in the file: complex128, or float64 when `real`;
//         ncols = nao*nao (full) or nao*(nao+1)/2 (`packed`: Hermitian lower triangle, row-major pack_tril)
//   out   (naux, nao, nao) complex128; aux rows >= rows are zero ("aux basis drop")
//   swapped: the entry belongs to the pair (k_j, k_i): out[L][p][q] = conj(src[L][q][p])
//   conj_all: the entry belongs to the time-reversed pair (-k_i, -k_j) [or, with swapped, (-k_j, -k_i)]: the
//         result is conjugated once more (files that keep only one member of each time-reversal pair)
// One CTA per 32x32 output tile of one auxiliary row: the source tile (the mirrored one for swapped entries and for
// the upper triangle of packed entries) is read with its fast index along threadIdx.x, parked in shared memory and
// written out transposed / conjugated as needed, so that both the loads and the stores are coalesced.  HBM-bound:
// 8-16 B read + 16 B written per element.
// ----------------------------------------------------------------------------------------------------------
template <bool REAL>
__device__ __forceinline__ double2 stored_elem(const void* __restrict__ src, size_t idx) {
    if (REAL) return make_double2(static_cast<const double*>(src)[idx], 0.0);
    return static_cast<const double2*>(src)[idx];
}

template <bool REAL>
__global__ void __launch_bounds__(256)
unpack_stored_kernel(const void* __restrict__ src, double2* __restrict__ out, int naux, int rows, int nao,
                     long long ncols, int packed, int swapped, int conj_all) {
    __shared__ double2 tile[32][33];
    const int p0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
    // source tile: rows r0.., columns c0.. of the (nao x nao) matrix the entry describes
    const bool mirror = packed ? (p0 < q0) : (swapped != 0);       // read tile (q0, p0) and transpose it
    const int r0 = mirror ? q0 : p0, c0 = mirror ? p0 : q0;
    for (int L = blockIdx.z; L < naux; L += gridDim.z) {
        double2* o = out + (size_t)L * nao * nao;
        if (L >= rows) {
            for (int dy = threadIdx.y; dy < 32; dy += 8) {
                const int p = p0 + dy, q = q0 + threadIdx.x;
                if (p < nao && q < nao) o[(size_t)p * nao + q] = make_double2(0.0, 0.0);
            }
            continue;
        }
        const size_t base = (size_t)L * (size_t)ncols;
        for (int dy = threadIdx.y; dy < 32; dy += 8) {
            const int r = r0 + dy, c = c0 + threadIdx.x;
            double2 v = make_double2(0.0, 0.0);
            if (r < nao && c < nao) {
                if (!packed) v = stored_elem<REAL>(src, base + (size_t)r * nao + c);
                else if (r >= c) v = stored_elem<REAL>(src, base + (size_t)r * (r + 1) / 2 + c);
            }
            tile[dy][threadIdx.x] = v;
        }
        __syncthreads();
        for (int dy = threadIdx.y; dy < 32; dy += 8) {
            const int p = p0 + dy, q = q0 + threadIdx.x;
            if (p < nao && q < nao) {
                double2 v;
                if (packed) {
                    // lower triangle stored; upper = conj of the mirrored element (unpack_tril, HERMITIAN)
                    if (p0 > q0 || (p0 == q0 && dy >= (int)threadIdx.x)) v = tile[dy][threadIdx.x];
                    else { v = tile[threadIdx.x][dy]; v.y = -v.y; }
                    if (swapped) v.y = -v.y;               // packed entry of the swapped pair: plain conjugate
                } else if (swapped) {
                    v = tile[threadIdx.x][dy];
                    v.y = -v.y;
                } else {
                    v = tile[dy][threadIdx.x];
                }
                if (conj_all) v.y = -v.y;                  // entry of the time-reversed pair (-k_i, -k_j)
                o[(size_t)p * nao + q] = v;
            }
        }
        __syncthreads();
    }
}

// The five real planes of the B operand that the 3-multiplication complex GEMM reads (zgemm_tn.cuh):
//   F[5 z + 0] = Br, [5 z + 1] = Bi - Br, [5 z + 2] = Br + Bi, [5 z + 3] = -(Br + Bi), [5 z + 4] = Br - Bi
// for every slice z of B (zb, N, K) complex.  Rows are Kp = K rounded up to a multiple of 8 doubles, zero padded, and
// within every group of 8 the words are stored in the order k = 0,4,1,5,2,6,3,7, so that the two k the MMA lane
// t needs in the two k-steps of a pipeline stage (t and t + 4) are one aligned 16-byte chunk.
__global__ void zforms_kernel(const double2* __restrict__ B, double* __restrict__ F, long long rows, int K, int Kp,
                              long long N) {
    const long long total = rows * Kp;
    const long long ps = N * Kp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / Kp;
        const int k = (int)(idx - r * Kp);
        const long long z = r / N, n = r - z * N;
        const double2 b = k < K ? B[r * K + k] : make_double2(0.0, 0.0);
        const int kpos = (k & ~7) + 2 * (k & 3) + ((k >> 2) & 1);
        double* f = F + (5 * z * N + n) * Kp + kpos;
        const double sum = b.x + b.y, dif = b.y - b.x;
        f[0] = b.x;
        f[ps] = dif;
        f[2 * ps] = sum;
        f[3 * ps] = -sum;
        f[4 * ps] = -dif;
    }
}

// Lower-triangular packing of a stack of square complex matrices, out[L][m(m+1)/2 + n] = in[L][m][n] (n <= m) --
// PySCF `lib.pack_tril` as `transform_gdf_to_lo` applies it to the LO-basis GDF blocks it stores
// (eri_transform.py:1386-1390): complex for k_i == k_j, real part only (with max|imag| reported) when both k-points
// are Gamma.  One CTA per matrix; reads run along n, writes are contiguous.
template <bool OUT_REAL>
__global__ void __launch_bounds__(256)
pack_tril_kernel(const double2* __restrict__ in, void* __restrict__ out, int n, long long npair,
                 unsigned long long* imag_max) {
    const double2* src = in + (size_t)blockIdx.x * n * n;
    double im_max = 0.0;
    for (long long P = threadIdx.x; P < npair; P += blockDim.x) {
        int m = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
        while ((long long)(m + 1) * (m + 2) / 2 <= P) ++m;
        while ((long long)m * (m + 1) / 2 > P) --m;
        const int c = (int)(P - (long long)m * (m + 1) / 2);
        const double2 v = src[(size_t)m * n + c];
        if (OUT_REAL) {
            static_cast<double*>(out)[(size_t)blockIdx.x * npair + P] = v.x;
            im_max = fmax(im_max, fabs(v.y));
        } else {
            static_cast<double2*>(out)[(size_t)blockIdx.x * npair + P] = v;
        }
    }
    if (OUT_REAL && imag_max) warp_atomic_max_abs(im_max, imag_max);
}

// real -> complex widening copy (basis in R-space is real; the GEMM operands are complex)
__global__ void d2z_kernel(const double* __restrict__ in, double2* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_double2(in[i], 0.0);
}

// ----------------------------------------------------------------------------------------------------------
// Symmetrise + pack + transpose of the accumulated embedding-orbital 3-index tensor of one transfer momentum.
//   S_sym [L][x][y]  : sum over the blocks the reference symmetrises (hermi_sum SYMMETRIC, eri_transform.py:371-373)
//   S_pln [L][n][m]  : sum over the blocks it does not, stored transposed, S_pln[L][n][m] = T[L][m][n]
//   Lambda[L, tri(m,n)] = S_sym[L][m][n] + S_sym[L][n][m] + S_pln[L][n][m]     (m >= n; pack_tril, l.375)
// Output goes K-contiguous for the stage-3 GEMM:  XT[P][col_re + L] = Re Lambda, XT[P][col_im + L] = Im Lambda.
// One CTA handles a 16x16 (m, n) tile (n-tile <= m-tile only: the grid's x index runs over those pairs) for 16
// consecutive L: reads are 256-byte rows, writes are 128-byte rows.  The loads of 4 consecutive L are issued together
// (12 independent 16-byte loads per thread in flight, two CTAs per SM) and each batch costs one barrier: after it,
// thread (tx, ty) is the only one that touches element [l][tx][ty].
// ----------------------------------------------------------------------------------------------------------
template <bool GSO>
__global__ void __launch_bounds__(256, 2)
pack_sym_kernel(const double2* __restrict__ S_sym, const double2* __restrict__ S_pln,
                const double2* __restrict__ S_sym2, const double2* __restrict__ S_pln2, double* __restrict__ XT,
                int naux, int neo, long long ldx, long long col_re, long long col_im) {
    // GSO: the second set (S_sym2, S_pln2) is SUBTRACTED: Lambda_a - Lambda_b of the generalised-spin-orbital
    // embedding ERI (reference: _Lij_s4_to_eri_gso, eri_transform.py:1252-1284, whose four signed Gram
    // products equal one Gram product of the difference)
    extern __shared__ double2 v_raw[];
    double2 (*v)[16][17] = reinterpret_cast<double2 (*)[16][17]>(v_raw);   // [L][m][n]
    int tm = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
    while ((tm + 1) * (tm + 2) / 2 <= (int)blockIdx.x) ++tm;
    while (tm * (tm + 1) / 2 > (int)blockIdx.x) --tm;
    const int tn = (int)blockIdx.x - tm * (tm + 1) / 2;
    const int L0 = blockIdx.y * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t n2 = (size_t)neo * neo;
    const int m = tm * 16 + ty, n = tn * 16 + tx;      // contiguous reads along n:  S[L][m][n]
    const int m2 = tm * 16 + tx, n2i = tn * 16 + ty;    // contiguous reads along m:  S[L][n][m]
    const bool in1 = m < neo && n < neo, in2 = m2 < neo && n2i < neo;
    const size_t o1 = (size_t)m * neo + n, o2 = (size_t)n2i * neo + m2;
    const double2 zero = make_double2(0.0, 0.0);
    constexpr int LB = 4;
#pragma unroll 1
    for (int lb = 0; lb < 16; lb += LB) {
        double2 a1[LB], b1[LB], b2[LB], a2[GSO ? LB : 1], b3[GSO ? LB : 1], b4[GSO ? LB : 1];
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const bool ok = L0 + lb + l < naux;
            const size_t base = (size_t)(L0 + lb + l) * n2;
            a1[l] = (ok && in1 && S_sym) ? S_sym[base + o1] : zero;
            b1[l] = (ok && in2 && S_sym) ? S_sym[base + o2] : zero;
            b2[l] = (ok && in2 && S_pln) ? S_pln[base + o2] : zero;
            if constexpr (GSO) {
                a2[l] = (ok && in1 && S_sym2) ? S_sym2[base + o1] : zero;
                b3[l] = (ok && in2 && S_sym2) ? S_sym2[base + o2] : zero;
                b4[l] = (ok && in2 && S_pln2) ? S_pln2[base + o2] : zero;
            }
        }
        double2 bs[LB];
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            double2 a = a1[l];
            bs[l] = make_double2(b1[l].x + b2[l].x, b1[l].y + b2[l].y);
            if constexpr (GSO) {
                a.x -= a2[l].x;
                a.y -= a2[l].y;
                bs[l].x -= b3[l].x + b4[l].x;
                bs[l].y -= b3[l].y + b4[l].y;
            }
            v[lb + l][ty][tx] = a;            // (m = ty, n = tx)
        }
        __syncthreads();
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            double2 acc = v[lb + l][tx][ty];  // element (m = tx, n = ty), whose transposed-read part is this thread's b
            acc.x += bs[l].x;
            acc.y += bs[l].y;
            v[lb + l][tx][ty] = acc;
        }
    }
    __syncthreads();
    // write: 16 pairs per pass, 16 L each (128-byte rows of XT)
    const int l = threadIdx.x & 15;
    for (int pp = threadIdx.x >> 4; pp < 256; pp += 16) {
        const int mi = pp >> 4, ni = pp & 15;
        const int mm = tm * 16 + mi, nn = tn * 16 + ni;
        if (mm >= neo || nn > mm || L0 + l >= naux) continue;
        const long long P = (long long)mm * (mm + 1) / 2 + nn;
        const double2 val = v[l][mi][ni];
        XT[P * ldx + col_re + L0 + l] = val.x;
        if (col_im >= 0) XT[P * ldx + col_im + L0 + l] = val.y;
    }
}

// zero-fill columns [c0, c1) of the rows of XT (K-padding so stale data never enters a Gram product)
__global__ void fill_cols_kernel(double* __restrict__ XT, long long rows, long long ldx, long long c0, long long c1) {
    const long long w = c1 - c0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * w;
         i += (long long)gridDim.x * blockDim.x)
        XT[(i / w) * ldx + c0 + (i % w)] = 0.0;
}

// ----------------------------------------------------------------------------------------------------------
// eri[P][Q] = eri[Q][P] for Q > P : fills the upper triangle from the lower one (the reference's lib.dot
// produces both, eri_transform.py:455-459).
// ----------------------------------------------------------------------------------------------------------
// One CTA per 64x64 tile on or below the diagonal (1-D grid over the nt(nt+1)/2 such tiles, no idle CTAs); 16 words
// per thread in flight, 512-byte row segments on both sides, transposition through a padded shared-memory tile.
constexpr int MIRROR_T = 64;
__global__ void __launch_bounds__(256)
mirror_lower_kernel(double* __restrict__ E, int n, long long ld) {
    __shared__ double tile[MIRROR_T][MIRROR_T + 1];
    // tile (tp, tq), tq <= tp, from the linear index
    const long long t = blockIdx.x;
    int tp = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)(tp + 1) * (tp + 2) / 2 <= t) ++tp;
    while ((long long)tp * (tp + 1) / 2 > t) --tp;
    const int tq = (int)(t - (long long)tp * (tp + 1) / 2);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    double v[2 * MIRROR_T / 8];
#pragma unroll
    for (int i = 0; i < MIRROR_T / 8; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = tp * MIRROR_T + ty + 8 * i, c = tq * MIRROR_T + tx + 32 * h;
            v[2 * i + h] = (r < n && c < n) ? E[(long long)r * ld + c] : 0.0;
        }
#pragma unroll
    for (int i = 0; i < MIRROR_T / 8; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) tile[ty + 8 * i][tx + 32 * h] = v[2 * i + h];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MIRROR_T / 8; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = tq * MIRROR_T + ty + 8 * i, c = tp * MIRROR_T + tx + 32 * h;   // destination (rows tq, cols tp)
            if (r < n && c < n && c > r) E[(long long)r * ld + c] = tile[tx + 32 * h][ty + 8 * i];
        }
}

// ----------------------------------------------------------------------------------------------------------
// s4 -> s1 : out[i][j][k][l] = eri4[tri(i,j)][tri(k,l)]   (pyscf ao2mo.restore(1, ...), eri_transform.py:529,543)
// s4 -> s8 : out[tri(P,Q)]   = eri4[P][Q], P >= Q          (ao2mo.restore(8, ...))
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri_idx(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

template <bool STAGED>
__global__ void __launch_bounds__(256)
restore_s1_kernel(const double* __restrict__ eri4, double* __restrict__ out, int n, long long npair) {
    // one CTA per packed row P = (i >= j): the row is staged in shared memory once (STAGED; rows beyond 200 KB are
    // gathered from global memory / L2 instead) and unpacked into the two (n, n) slabs out[i][j][:][:] and
    // out[j][i][:][:] with coalesced writes
    extern __shared__ double s1row[];
    const long long P = blockIdx.x;
    int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
    while ((long long)i * (i + 1) / 2 > P) --i;
    const int j = (int)(P - (long long)i * (i + 1) / 2);
    const double* row = eri4 + P * npair;
    if (STAGED) {
        for (long long q = threadIdx.x; q < npair; q += blockDim.x) s1row[q] = row[q];
        __syncthreads();
    }
    const double* src = STAGED ? s1row : row;
    double* dst_ij = out + ((long long)i * n + j) * n * n;
    double* dst_ji = out + ((long long)j * n + i) * n * n;
    for (int kl = threadIdx.x; kl < n * n; kl += blockDim.x) {
        const int k = kl / n, l = kl - k * n;
        const double v = src[tri_idx(k, l)];
        dst_ij[kl] = v;
        if (i != j) dst_ji[kl] = v;
    }
}

__global__ void __launch_bounds__(256)
restore_s8_kernel(const double* __restrict__ eri4, double* __restrict__ out, long long npair) {
    const long long P = blockIdx.x;
    const double* row = eri4 + P * npair;
    double* dst = out + P * (P + 1) / 2;
    for (long long Q = threadIdx.x; Q <= P; Q += blockDim.x) dst[Q] = row[Q];
}

// ----------------------------------------------------------------------------------------------------------
// J/K contraction of an s4 ERI with density matrices (PySCF hf.dot_eri_dm; call site libdmet/solver/scf.py:300-326;
// convention J_ij = sum_kl (ij|kl) D_kl, K_jk = sum_il (ij|kl) D_il, scf.py:269-271).
// One CTA per packed row P = (i >= j): the row is staged in shared memory once and used for
//    vj[P]      = sum_Q row[Q] * dd[Q]                       dd[Q=(k>=l)] = D_kl + D_lk (k != l), D_kk
//    y1[k]      = sum_l M[k][l] D[i][l]   -> contributes to K[j][k]
//    y2[k]      = sum_l M[k][l] D[j][l]   -> contributes to K[i][k]   (i != j)
// partial rows are written out and reduced in fixed order by jk_reduce_kernel (deterministic, no atomics).
// ----------------------------------------------------------------------------------------------------------
// dd[Q = (k >= l)] = D_kl + D_lk (k != l), D_kk : the packed density the J dot product runs against
__global__ void jk_pack_dm_kernel(const double* __restrict__ D, double* __restrict__ dd, int n) {
    const long long npair = (long long)n * (n + 1) / 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npair;
         q += (long long)gridDim.x * blockDim.x) {
        int k = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
        while ((long long)(k + 1) * (k + 2) / 2 <= q) ++k;
        while ((long long)k * (k + 1) / 2 > q) --k;
        const int l = (int)(q - (long long)k * (k + 1) / 2);
        dd[q] = k == l ? D[(size_t)k * n + k] : D[(size_t)k * n + l] + D[(size_t)l * n + k];
    }
}

template <bool STAGED>
__global__ void __launch_bounds__(256)
jk_rows_kernel(const double* __restrict__ eri4, const double* __restrict__ D, const double* __restrict__ dd,
               double* __restrict__ vj_packed, double* __restrict__ kpart, int n, long long npair, int with_k) {
    // STAGED: the packed row lives in shared memory; otherwise (rows beyond 200 KB) it is re-read from global / L2
    extern __shared__ double jk_smem[];  // [npair if STAGED] packed row, [n] D[i,:], [n] D[j,:], [8] reduction
    double* srow = jk_smem;
    double* dI = jk_smem + (STAGED ? npair : 0);
    double* dJ = dI + n;
    double* red = dJ + n;
    const long long P = blockIdx.x;
    int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
    while ((long long)i * (i + 1) / 2 > P) --i;
    const int j = (int)(P - (long long)i * (i + 1) / 2);
    const double* row = eri4 + P * npair;
    // stage the row (coalesced) and form the J dot product on the fly
    double part = 0.0;
    for (long long q = threadIdx.x; q < npair; q += blockDim.x) {
        const double v = row[q];
        if (STAGED) srow[q] = v;
        part = fma(v, dd[q], part);
    }
    const double* src = STAGED ? srow : row;
    for (int l = threadIdx.x; l < n; l += blockDim.x) {
        dI[l] = D[(size_t)i * n + l];
        dJ[l] = D[(size_t)j * n + l];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        vj_packed[P] = s;
    }
    if (!with_k) return;
    // K partials: y1[k] = sum_l M[k][l] D[i][l], y2[k] = sum_l M[k][l] D[j][l] with M the symmetric unpacking of the
    // row.  Thread k walks its own packed row segment (l <= k), then the column below the diagonal (l > k), where
    // neighbouring threads read neighbouring words.
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double y1 = 0.0, y2 = 0.0;
        const double* seg = src + (size_t)k * (k + 1) / 2;
        for (int l = 0; l <= k; ++l) {
            const double m = seg[l];
            y1 = fma(m, dI[l], y1);
            y2 = fma(m, dJ[l], y2);
        }
        size_t off = (size_t)(k + 1) * (k + 2) / 2 + k;       // element (l = k + 1, k)
        for (int l = k + 1; l < n; ++l) {
            const double m = src[off];
            y1 = fma(m, dI[l], y1);
            y2 = fma(m, dJ[l], y2);
            off += (size_t)l + 1;
        }
        kpart[(P * 2 + 0) * n + k] = y1;
        kpart[(P * 2 + 1) * n + k] = y2;
    }
}

// Streaming variant for n <= 32 * NCH (<= 256): nothing is staged.  Warp w owns the rows k = w, w + 8, ... of the
// symmetric matrix M packed in eri4[P][:]; lane reads the words l = lane + 32 c <= k of the row segment straight from
// global memory (each word is used once).  A word M[k][l] feeds the row sum y[k] += M[k][l] d[l] (warp reduction) and,
// for l < k, the column sum y[l] += M[k][l] d[k], which lives in the lane's registers for all rows of the warp.
template <int NCH>
__global__ void __launch_bounds__(256)
jk_rows_seg_kernel(const double* __restrict__ eri4, const double* __restrict__ D, const double* __restrict__ dd,
                   double* __restrict__ vj_packed, double* __restrict__ kpart, int n, long long npair, int with_k) {
    constexpr int NP = NCH * 32;
    __shared__ double sDI[NP], sDJ[NP];
    __shared__ double sRow[2][NP];
    __shared__ double sCol[8][2][NP];
    __shared__ double sJ[8];
    const long long P = blockIdx.x;
    int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
    while ((long long)i * (i + 1) / 2 > P) --i;
    const int j = (int)(P - (long long)i * (i + 1) / 2);
    for (int l = threadIdx.x; l < NP; l += blockDim.x) {
        sDI[l] = l < n ? D[(size_t)i * n + l] : 0.0;
        sDJ[l] = l < n ? D[(size_t)j * n + l] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double dIl[NCH], dJl[NCH], c1[NCH], c2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        dIl[c] = sDI[lane + 32 * c];
        dJl[c] = sDJ[lane + 32 * c];
        c1[c] = 0.0;
        c2[c] = 0.0;
    }
    const double* row = eri4 + P * npair;
    double jacc = 0.0;
    for (int k = w; k < n; k += 8) {
        const size_t base = (size_t)k * (k + 1) / 2;
        const double* seg = row + base;
        const double* dseg = dd + base;
        double m[NCH], g[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int l = lane + 32 * c;
            const bool in = l <= k;
            m[c] = in ? seg[l] : 0.0;
            g[c] = in ? dseg[l] : 0.0;
        }
        const double dik = sDI[k], djk = sDJ[k];
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            jacc = fma(m[c], g[c], jacc);
            if (with_k) {
                r1 = fma(m[c], dIl[c], r1);
                r2 = fma(m[c], dJl[c], r2);
                const double mc = (lane + 32 * c < k) ? m[c] : 0.0;      // the diagonal word enters once
                c1[c] = fma(mc, dik, c1[c]);
                c2[c] = fma(mc, djk, c2[c]);
            }
        }
        if (with_k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                r1 += __shfl_xor_sync(0xffffffffu, r1, o);
                r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            }
            if (lane == 0) {
                sRow[0][k] = r1;
                sRow[1][k] = r2;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) jacc += __shfl_xor_sync(0xffffffffu, jacc, o);
    if (lane == 0) sJ[w] = jacc;
    if (with_k) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            sCol[w][0][lane + 32 * c] = c1[c];
            sCol[w][1][lane + 32 * c] = c2[c];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sj = 0.0;
        for (int x = 0; x < 8; ++x) sj += sJ[x];
        vj_packed[P] = sj;
    }
    if (!with_k) return;
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
        const int v = t >= n, k = t - v * n;
        double y = sRow[v][k];
#pragma unroll
        for (int x = 0; x < 8; ++x) y += sCol[x][v][k];
        kpart[(P * 2 + v) * n + k] = y;
    }
}

// Bulk-copy variant (n <= NP, row fits in shared memory): one thread brings the whole packed row in with
// cp.async.bulk -- a single transaction in flight while the SM's other resident CTA computes.  J is a flat dot product
// against dd (streamed from L2).  K: thread k of group g walks row k of the symmetric matrix M over its half of the
// l range -- M[k][l] sits at k(k+1)/2 + l for l <= k and at l(l+1)/2 + k beyond the diagonal, both bank-conflict free
// across consecutive k -- so y[k] needs no cross-thread reduction; (D[i][l], D[j][l]) is one broadcast 16-byte load.
// Rows start on 8-byte boundaries only, so the copy covers the enclosing 16-byte aligned range; a tail that would
// cross the end of the tensor is fetched with plain loads.
template <int NP>
__global__ void __launch_bounds__(2 * NP)
jk_rows_bulk_kernel(const double* __restrict__ eri4, const double* __restrict__ D, const double* __restrict__ dd,
                    double* __restrict__ vj_packed, double* __restrict__ kpart, int n, long long npair, int with_k) {
    extern __shared__ __align__(16) double jkb_row[];          // npair + 2 words
    __shared__ double2 sD[NP];                                 // (D[i][l], D[j][l])
    __shared__ double sY[2][2][NP];
    __shared__ double sJ[2 * NP / 32];
    __shared__ __align__(8) unsigned long long bar_store;
    const long long P = blockIdx.x;
    const long long off = P * npair;
    const long long a0 = off & ~1LL;
    const int head = (int)(off - a0);
    long long cnt = ((long long)head + npair + 1) & ~1LL;
    if (a0 + cnt > npair * npair) cnt -= 2;                     // stay inside the tensor; the tail comes below
    const uint32_t bar = smem_u32(&bar_store);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, (uint32_t)(cnt * 8));
        const uint32_t dst = smem_u32(jkb_row);
        for (long long c = 0; c < cnt; c += 4096)
            bulk_load_1d(dst + (uint32_t)(c * 8), eri4 + a0 + c, (uint32_t)(min(4096LL, cnt - c) * 8), bar);
    }
    int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
    while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
    while ((long long)i * (i + 1) / 2 > P) --i;
    const int j = (int)(P - (long long)i * (i + 1) / 2);
    for (int l = threadIdx.x; l < NP; l += blockDim.x)
        sD[l] = l < n ? make_double2(D[(size_t)i * n + l], D[(size_t)j * n + l]) : make_double2(0.0, 0.0);
    double* srow = jkb_row + head;
    for (long long q = cnt - head + threadIdx.x; q < npair; q += blockDim.x) srow[q] = eri4[off + q];
    __syncthreads();                                            // barrier initialised, tail and D rows in place
    mbar_wait(bar, 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double jacc = 0.0, jacc2 = 0.0;
    {
        // eight independent dd loads (L2) in flight per thread; a two-way unrolled chain exposed their latency
        constexpr int CH = 8;
        for (long long q0 = threadIdx.x; q0 < npair; q0 += (long long)CH * blockDim.x) {
            double e[CH], gq[CH];
#pragma unroll
            for (int x = 0; x < CH; ++x) {
                const long long q = q0 + (long long)x * blockDim.x;
                const bool in = q < npair;
                e[x] = in ? srow[q] : 0.0;
                gq[x] = in ? __ldg(dd + q) : 0.0;
            }
#pragma unroll
            for (int x = 0; x < CH; x += 2) {
                jacc = fma(e[x], gq[x], jacc);
                jacc2 = fma(e[x + 1], gq[x + 1], jacc2);
            }
        }
        jacc += jacc2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) jacc += __shfl_xor_sync(0xffffffffu, jacc, o);
    if (lane == 0) sJ[w] = jacc;
    if (with_k) {
        const int g = threadIdx.x >= NP, k = threadIdx.x - g * NP;
        const int half = (n + 1) >> 1;
        const int l0 = g * half, l1 = min(n, l0 + half);
        double y1 = 0.0, y2 = 0.0, z1 = 0.0, z2 = 0.0;
        if (k < n) {
            int offA = k * (k + 1) / 2 + l0;                    // M[k][l], l <= k
            int offB = l0 * (l0 + 1) / 2 + k;                   // M[l][k], l > k
            int l = l0;
            for (; l + 1 < l1; l += 2) {
                const double m0 = srow[l <= k ? offA : offB];
                const double m1 = srow[l + 1 <= k ? offA + 1 : offB + l + 1];
                const double2 d0 = sD[l], d1 = sD[l + 1];
                y1 = fma(m0, d0.x, y1);
                y2 = fma(m0, d0.y, y2);
                z1 = fma(m1, d1.x, z1);
                z2 = fma(m1, d1.y, z2);
                offA += 2;
                offB += 2 * l + 3;
            }
            if (l < l1) {
                const double m0 = srow[l <= k ? offA : offB];
                const double2 d0 = sD[l];
                y1 = fma(m0, d0.x, y1);
                y2 = fma(m0, d0.y, y2);
            }
        }
        sY[g][0][k] = y1 + z1;
        sY[g][1][k] = y2 + z2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sj = 0.0;
        for (int x = 0; x < 2 * NP / 32; ++x) sj += sJ[x];
        vj_packed[P] = sj;
    }
    if (!with_k) return;
    for (int t = threadIdx.x; t < 2 * n; t += blockDim.x) {
        const int v = t >= n, k = t - v * n;
        kpart[(P * 2 + v) * n + k] = sY[0][v][k] + sY[1][v][k];
    }
}

// ----------------------------------------------------------------------------------------------------------
// J/K from the LOWER TRIANGLE of a symmetric s4 ERI (eri4[P][Q] == eri4[Q][P]: the restricted / aa / bb blocks, which
// are Gram matrices) and a symmetric density matrix (the reference always passes hermi=1, solver/scf.py:300-326):
// only the words Q <= P of row P are read, i.e. half of the tensor crosses HBM.
//   J:  vj[P] = sum_{Q <= P} E[P][Q] dd[Q]  (row part, kept per row)
//             + sum_{Q >  P} E[Q][P] dd[Q]  (column part: every row P' adds dd[P'] E[P'][Q] to the columns Q < P';
//                                            accumulated in registers over the rows of a CTA, one partial vector
//                                            per CTA, summed in a fixed order afterwards)
//   K:  with M_P the symmetric matrix packed in row P, truncated to the words Q <= P (the word Q = P halved),
//       S[j][:] += M_P D[i][:],  S[i][:] += M_P D[j][:] (i != j),   K = S + S^T.
// One CTA walks JKT_ROWS consecutive rows (heaviest row blocks are scheduled first); a row arrives in shared memory
// by cp.async.bulk into one of two buffers, so the next row streams in while the current one is used.  Per word: one
// conflict-free LDS + one cached load of dd for J; the K walk of the bulk kernel above on the truncated matrix (the
// unused tail of packed row i is zero-filled, so the walk needs no predicates).  Deterministic: no atomics.
// ----------------------------------------------------------------------------------------------------------
constexpr int JKT_ROWS = 32;
constexpr int JKT_MAXBUF = 8;

template <int NP>
__global__ void __launch_bounds__(2 * NP, 1)
jk_tri_kernel(const double* __restrict__ eri4, const double* __restrict__ D, const double* __restrict__ dd,
              double* __restrict__ vj_row, double* __restrict__ jpart, double* __restrict__ kpart, int n,
              long long npair, int with_k, int arena_words) {
    constexpr int T = 2 * NP;
    constexpr int NCMAX = (NP * (NP + 1) / 2 + T - 1) / T;
    extern __shared__ __align__(16) double jkt_rows[];         // arena_words doubles: a ring of row slots
    __shared__ double2 sD[NP];                                 // (D[i][l], D[j][l])
    __shared__ double sY[2][2][NP];
    __shared__ double sJ[T / 32];
    __shared__ __align__(8) unsigned long long bars[JKT_MAXBUF];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long nblk = (npair + JKT_ROWS - 1) / JKT_ROWS;
    const long long B = nblk - 1 - blockIdx.x;                  // heaviest blocks first
    const long long Phi = min(npair, (B + 1) * JKT_ROWS) - 1, Plo = B * JKT_ROWS;
    const int nrows = (int)(Phi - Plo + 1);
    const long long total = npair * npair;
    // Ring of row slots sized for the longest row of this block (packed row i of the truncated matrix may be zero
    // filled up to its end, hence the extra i + 1 words).  Short rows get a deep ring: a row costs at least one DRAM
    // round trip, so enough of them must be in flight to keep the SM's share of the bandwidth busy.
    const int slot_words = (int)((Phi + 2 * NP + 6) & ~1LL);
    const int nbuf = max(2, min(JKT_MAXBUF, arena_words / slot_words));
    if (t == 0) {
        for (int x = 0; x < JKT_MAXBUF; ++x) mbar_init(smem_u32(&bars[x]), 1);
        mbar_fence_init();
    }
    __syncthreads();

    // bring the words 0..P of row P into slot b (16-byte aligned enclosing range by bulk copy, a tail that would
    // cross the end of the tensor by plain loads)
    auto issue = [&](long long P, int b) {
        const long long off = P * npair, a0 = off & ~1LL;
        const int head = (int)(off - a0);
        long long cnt = ((long long)head + P + 2) & ~1LL;
        if (a0 + cnt > total) cnt -= 2;
        double* buf = jkt_rows + (size_t)b * slot_words;
        if (t == 0) {
            fence_proxy_async_smem();
            const uint32_t bar = smem_u32(&bars[b]);
            mbar_expect_tx(bar, (uint32_t)(cnt * 8));
            const uint32_t dst = smem_u32(buf);
            for (long long c = 0; c < cnt; c += 4096)
                bulk_load_1d(dst + (uint32_t)(c * 8), eri4 + a0 + c, (uint32_t)(min(4096LL, cnt - c) * 8), bar);
        }
        for (long long q = cnt - head + t; q <= P; q += T) buf[head + q] = eri4[off + q];
    };

    double colacc[NCMAX];
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) colacc[c] = 0.0;

    for (int r = 0; r < min(nbuf - 1, nrows); ++r) issue(Phi - r, r);
    __syncthreads();                                                // plain-load tail of the very last row
    for (int r = 0; r < nrows; ++r) {
        const long long P = Phi - r;
        const int b = r % nbuf;
        if (r + nbuf - 1 < nrows) issue(P - (nbuf - 1), (r + nbuf - 1) % nbuf);   // the slot row r - 1 has just left
        int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
        while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
        while ((long long)i * (i + 1) / 2 > P) --i;
        const int j = (int)(P - (long long)i * (i + 1) / 2);
        for (int l = t; l < NP; l += T)
            sD[l] = l < n ? make_double2(D[(size_t)i * n + l], D[(size_t)j * n + l]) : make_double2(0.0, 0.0);
        double* srow = jkt_rows + (size_t)b * slot_words + (int)((P * npair) & 1LL);
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((r / nbuf) & 1));
        // ---- J: row dot product and column updates ----
        const double ddP = dd[P];
        double racc = 0.0;
        const int Pi = (int)P;
        // eight independent (row word, dd) load pairs are issued before their first use: a load / FMA chain per
        // word would expose one L2 round trip per word (dd does not fit beside the row slots in L1)
        constexpr int CH = 8;
#pragma unroll
        for (int c0 = 0; c0 < NCMAX; c0 += CH) {
            if (c0 * T > Pi) break;                                 // warp-uniform: the row ends before this chunk
            double e[CH], gq[CH];
#pragma unroll
            for (int x = 0; x < CH; ++x) {
                const int q = t + (c0 + x) * T;
                const bool in = (c0 + x < NCMAX) && q <= Pi;
                e[x] = in ? srow[q] : 0.0;
                gq[x] = in ? __ldg(dd + q) : 0.0;
            }
#pragma unroll
            for (int x = 0; x < CH; ++x) {
                if (c0 + x < NCMAX) {
                    const int q = t + (c0 + x) * T;
                    racc = fma(e[x], gq[x], racc);
                    colacc[c0 + x] = fma(q < Pi ? e[x] : 0.0, ddP, colacc[c0 + x]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
        if (lane == 0) sJ[w] = racc;
        __syncthreads();
        if (t == 0) {
            double sj = 0.0;
            for (int x = 0; x < T / 32; ++x) sj += sJ[x];
            vj_row[P] = sj;
        }
        if (with_k) {
            // the rest of packed row i is not part of the truncated matrix; the word Q = P counts half
            const int rowend = (i + 1) * (i + 2) / 2;
            for (int q = Pi + 1 + t; q < rowend; q += T) srow[q] = 0.0;
            if (t == T - 1) srow[P] *= 0.5;
            __syncthreads();
            const int ne = i + 1;
            const int g = t >= NP, k = t - g * NP;
            const int half = (ne + 1) >> 1;
            const int l0 = g * half, l1 = min(ne, l0 + half);
            double y1 = 0.0, y2 = 0.0, z1 = 0.0, z2 = 0.0;
            if (k < ne) {
                int offA = k * (k + 1) / 2 + l0;                    // M[k][l], l <= k
                int offB = l0 * (l0 + 1) / 2 + k;                   // M[l][k], l > k
                int l = l0;
                for (; l + 1 < l1; l += 2) {
                    const double m0 = srow[l <= k ? offA : offB];
                    const double m1 = srow[l + 1 <= k ? offA + 1 : offB + l + 1];
                    const double2 d0 = sD[l], d1 = sD[l + 1];
                    y1 = fma(m0, d0.x, y1);
                    y2 = fma(m0, d0.y, y2);
                    z1 = fma(m1, d1.x, z1);
                    z2 = fma(m1, d1.y, z2);
                    offA += 2;
                    offB += 2 * l + 3;
                }
                if (l < l1) {
                    const double m0 = srow[l <= k ? offA : offB];
                    const double2 d0 = sD[l];
                    y1 = fma(m0, d0.x, y1);
                    y2 = fma(m0, d0.y, y2);
                }
            }
            sY[g][0][k] = y1 + z1;
            sY[g][1][k] = y2 + z2;
            __syncthreads();
            for (int x = t; x < 2 * n; x += T) {
                const int v = x >= n, kk = x - v * n;
                kpart[(P * 2 + v) * n + kk] = sY[0][v][kk] + sY[1][v][kk];
            }
        }
        fence_proxy_async_smem();        // generic writes to this slot (zero fill, halving) before the next bulk copy
        __syncthreads();                                            // slot, sD, sY, sJ free for the next row
    }
    // column partials of this block: every column below the block's last row
    double* jp = jpart + (size_t)B * npair;
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
        const long long q = t + (long long)c * T;
        if (q < Phi) jp[q] = colacc[c];
    }
}

// The same kernel with the packed density dd held in REGISTERS: thread t owns the words q = t + c T of every row, so
// dd[q] is loaded once per CTA instead of once per row from L2 (the J pass of jk_tri_kernel is bound by the latency of
// those loads: 8 in flight per thread, 4.5 round trips per long row).  T threads (>= 2 NP; only the first 2 NP walk
// the K matrix) are chosen so that the column accumulators and dd -- 2 NCMAX doubles per thread -- leave room for the
// rest: NP = 152, T = 384 gives 2 x 31 doubles.  The K walk is unrolled four-fold (four independent load / FMA chains).
template <int NP, int T>
__global__ void __launch_bounds__(T, 1)
jk_tri2_kernel(const double* __restrict__ eri4, const double* __restrict__ D, const double* __restrict__ dd,
               double* __restrict__ vj_row, double* __restrict__ jpart, double* __restrict__ kpart, int n,
               long long npair, int with_k, int arena_words) {
    static_assert(T >= 2 * NP, "the K walk needs at least two groups of NP threads");
    constexpr int NG = T / NP;                                 // thread groups of the K walk
    constexpr int NCMAX = (NP * (NP + 1) / 2 + T - 1) / T;
    extern __shared__ __align__(16) double jkt_rows[];         // arena_words doubles: a ring of row slots
    __shared__ double2 sD[NP];                                 // (D[i][l], D[j][l])
    __shared__ double sY[NG][2][NP];
    __shared__ double sJ[T / 32];
    __shared__ __align__(8) unsigned long long bars[JKT_MAXBUF];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long nblk = (npair + JKT_ROWS - 1) / JKT_ROWS;
    const long long B = nblk - 1 - blockIdx.x;                  // heaviest blocks first
    const long long Phi = min(npair, (B + 1) * JKT_ROWS) - 1, Plo = B * JKT_ROWS;
    const int nrows = (int)(Phi - Plo + 1);
    const long long total = npair * npair;
    const int slot_words = (int)((Phi + 2 * NP + 6) & ~1LL);
    const int nbuf = max(2, min(JKT_MAXBUF, arena_words / slot_words));
    if (t == 0) {
        for (int x = 0; x < JKT_MAXBUF; ++x) mbar_init(smem_u32(&bars[x]), 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](long long P, int b) {
        const long long off = P * npair, a0 = off & ~1LL;
        const int head = (int)(off - a0);
        long long cnt = ((long long)head + P + 2) & ~1LL;
        if (a0 + cnt > total) cnt -= 2;
        double* buf = jkt_rows + (size_t)b * slot_words;
        if (t == 0) {
            fence_proxy_async_smem();
            const uint32_t bar = smem_u32(&bars[b]);
            mbar_expect_tx(bar, (uint32_t)(cnt * 8));
            const uint32_t dst = smem_u32(buf);
            for (long long c = 0; c < cnt; c += 4096)
                bulk_load_1d(dst + (uint32_t)(c * 8), eri4 + a0 + c, (uint32_t)(min(4096LL, cnt - c) * 8), bar);
        }
        for (long long q = cnt - head + t; q <= P; q += T) buf[head + q] = eri4[off + q];
    };

    for (int r = 0; r < min(nbuf - 1, nrows); ++r) issue(Phi - r, r);
    double colacc[NCMAX], ddreg[NCMAX];
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
        const long long q = t + (long long)c * T;
        colacc[c] = 0.0;
        ddreg[c] = q <= Phi ? dd[q] : 0.0;                           // the rows of this block end at Phi
    }
    __syncthreads();                                                // plain-load tail of the very last row
    for (int r = 0; r < nrows; ++r) {
        const long long P = Phi - r;
        const int b = r % nbuf;
        if (r + nbuf - 1 < nrows) issue(P - (nbuf - 1), (r + nbuf - 1) % nbuf);   // the slot row r - 1 has just left
        int i = (int)((sqrt(8.0 * (double)P + 1.0) - 1.0) * 0.5);
        while ((long long)(i + 1) * (i + 2) / 2 <= P) ++i;
        while ((long long)i * (i + 1) / 2 > P) --i;
        const int j = (int)(P - (long long)i * (i + 1) / 2);
        for (int l = t; l < NP; l += T)
            sD[l] = l < n ? make_double2(D[(size_t)i * n + l], D[(size_t)j * n + l]) : make_double2(0.0, 0.0);
        double* srow = jkt_rows + (size_t)b * slot_words + (int)((P * npair) & 1LL);
        const double ddP = dd[P];
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((r / nbuf) & 1));
        // ---- J: row dot product and column updates, no global loads ----
        double racc = 0.0, racc2 = 0.0;
        const int Pi = (int)P;
#pragma unroll
        for (int c = 0; c < NCMAX; c += 2) {
            if (c * T > Pi) break;                                  // warp-uniform: the row ends before this chunk
            const int q0 = t + c * T, q1 = q0 + T;
            const double e0 = q0 <= Pi ? srow[q0] : 0.0;
            const double e1 = (c + 1 < NCMAX && q1 <= Pi) ? srow[q1] : 0.0;
            racc = fma(e0, ddreg[c], racc);
            colacc[c] = fma(q0 < Pi ? e0 : 0.0, ddP, colacc[c]);
            if (c + 1 < NCMAX) {
                racc2 = fma(e1, ddreg[c + 1], racc2);
                colacc[c + 1] = fma(q1 < Pi ? e1 : 0.0, ddP, colacc[c + 1]);
            }
        }
        racc += racc2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) racc += __shfl_xor_sync(0xffffffffu, racc, o);
        if (lane == 0) sJ[w] = racc;
        __syncthreads();
        if (t == 0) {
            double sj = 0.0;
            for (int x = 0; x < T / 32; ++x) sj += sJ[x];
            vj_row[P] = sj;
        }
        if (with_k) {
            // the rest of packed row i is not part of the truncated matrix; the word Q = P counts half
            const int rowend = (i + 1) * (i + 2) / 2;
            for (int q = Pi + 1 + t; q < rowend; q += T) srow[q] = 0.0;
            if (t == T - 1) srow[P] *= 0.5;
            __syncthreads();
            // K walk: NG = T / NP groups of NP threads share the l range; thread (g, k) walks row k of the symmetric
            // matrix over its part of the l range -- M[k][l] sits at k (k + 1) / 2 + l for l <= k and at
            // l (l + 1) / 2 + k beyond the diagonal, both conflict free across consecutive k; every lane of a warp does
            // useful work in every trip (uniform l), (D[i][l], D[j][l]) is one broadcast 16-byte load
            const int ne = i + 1;
            const int g = t / NP, k = t - g * NP;
            const int part = (ne + NG - 1) / NG;
            const int l0 = g * part, l1 = min(ne, l0 + part);
            double y1 = 0.0, y2 = 0.0, z1 = 0.0, z2 = 0.0;
            if (g < NG && k < ne) {
                int offA = k * (k + 1) / 2 + l0;                    // M[k][l], l <= k
                int offB = l0 * (l0 + 1) / 2 + k;                   // M[l][k], l > k
                int l = l0;
                for (; l + 1 < l1; l += 2) {
                    const double m0 = srow[l <= k ? offA : offB];
                    const double m1 = srow[l + 1 <= k ? offA + 1 : offB + l + 1];
                    const double2 d0 = sD[l], d1 = sD[l + 1];
                    y1 = fma(m0, d0.x, y1);
                    y2 = fma(m0, d0.y, y2);
                    z1 = fma(m1, d1.x, z1);
                    z2 = fma(m1, d1.y, z2);
                    offA += 2;
                    offB += 2 * l + 3;
                }
                if (l < l1) {
                    const double m0 = srow[l <= k ? offA : offB];
                    const double2 d0 = sD[l];
                    y1 = fma(m0, d0.x, y1);
                    y2 = fma(m0, d0.y, y2);
                }
            }
            if (g < NG) {
                sY[g][0][k] = y1 + z1;
                sY[g][1][k] = y2 + z2;
            }
            __syncthreads();
            for (int x = t; x < 2 * n; x += T) {
                const int v = x >= n, kk = x - v * n;
                double sum = sY[0][v][kk];
#pragma unroll
                for (int gg = 1; gg < NG; ++gg) sum += sY[gg][v][kk];
                kpart[(P * 2 + v) * n + kk] = sum;
            }
        }
        fence_proxy_async_smem();        // generic writes to this slot (zero fill, halving) before the next bulk copy
        __syncthreads();                                            // slot, sD, sY, sJ free for the next row
    }
    double* jp = jpart + (size_t)B * npair;
#pragma unroll
    for (int c = 0; c < NCMAX; ++c) {
        const long long q = t + (long long)c * T;
        if (q < Phi) jp[q] = colacc[c];
    }
}

// vj_packed[q] = vj_row[q] + sum over the row blocks B >= q / JKT_ROWS of their column partials (fixed order)
__global__ void __launch_bounds__(256)
jk_tri_jsum_kernel(const double* __restrict__ vj_row, const double* __restrict__ jpart,
                   double* __restrict__ vj_packed, long long npair) {
    // 32 consecutive columns per CTA (threadIdx.x), the row blocks dealt to the 8 warps (threadIdx.y) with four
    // independent loads in flight each; the eight partial sums are added in a fixed order
    __shared__ double part[8][33];
    const long long nblk = (npair + JKT_ROWS - 1) / JKT_ROWS;
    const long long q = (long long)blockIdx.x * 32 + threadIdx.x;
    const int w = threadIdx.y;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (q < npair) {
        // block B holds a value for column q iff q < (last row of B) ; B >= q / JKT_ROWS, and for B == q / JKT_ROWS
        // only if q is not that block's last row
        long long B = q / JKT_ROWS;
        if (q >= min(npair, (B + 1) * JKT_ROWS) - 1) ++B;
        B += w;
        for (; B + 24 < nblk; B += 32) {
            s0 += jpart[B * npair + q];
            s1 += jpart[(B + 8) * npair + q];
            s2 += jpart[(B + 16) * npair + q];
            s3 += jpart[(B + 24) * npair + q];
        }
        for (; B < nblk; B += 8) s0 += jpart[B * npair + q];
    }
    part[w][threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0 && q < npair) {
        double s = vj_row[q];
#pragma unroll
        for (int x = 0; x < 8; ++x) s += part[x][threadIdx.x];
        vj_packed[q] = s;
    }
}

// K = S + S^T
__global__ void symmetrise_add_kernel(const double* __restrict__ S, double* __restrict__ K, int n) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int a = idx / n, b = idx - a * n;
        K[idx] = S[idx] + S[(size_t)b * n + a];
    }
}

// K[a][k] = sum_{i >= a} y1(P(i, a))[k] + sum_{j < a} y2(P(a, j))[k], terms dealt to the 8 warps and summed in a
// fixed order (deterministic)
__global__ void __launch_bounds__(256)
jk_reduce_kernel(const double* __restrict__ kpart, double* __restrict__ vk, int n) {
    extern __shared__ double jkr_smem[];          // [8][n]
    const int a = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = lane; k < n; k += 32) {
        double s = 0.0;
        for (int t = w; t < n; t += 8) {
            const long long src = t >= a ? ((long long)t * (t + 1) / 2 + a) * 2 + 0
                                         : ((long long)a * (a + 1) / 2 + t) * 2 + 1;
            s += kpart[src * n + k];
        }
        jkr_smem[(size_t)w * n + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int x = 0; x < 8; ++x) s += jkr_smem[(size_t)x * n + k];
        vk[(size_t)a * n + k] = s;
    }
}

// ----------------------------------------------------------------------------------------------------------
// DMET energy weights (reference: get_H2_scaled, libdmet/routine/slater.py:1734-1778): every two-electron
// integral is scaled by (number of impurity indices among its four) / 4.
//   s4: E[P][Q] *= (w[P] + w[Q]) / 4,  w[P] = impurity count of the pair P in {0, 1, 2}
//   s1: E[i][j][k][l] *= (m[i] + m[j] + m[k] + m[l]) / 4,  m = 0/1 impurity flag
// ----------------------------------------------------------------------------------------------------------
__global__ void scale_s4_kernel(double* __restrict__ E, const int* __restrict__ w, long long npair) {
    const long long P = blockIdx.x;
    const double wp = (double)w[P];
    double* row = E + P * npair;
    for (long long Q = threadIdx.x; Q < npair; Q += blockDim.x) row[Q] *= (wp + (double)w[Q]) * 0.25;
}

__global__ void scale_s1_kernel(double* __restrict__ E, const int* __restrict__ m, int n) {
    const int ij = blockIdx.x;
    const int i = ij / n, j = ij - i * n;
    const double mij = (double)(m[i] + m[j]);
    double* dst = E + (long long)ij * n * n;
    for (int kl = threadIdx.x; kl < n * n; kl += blockDim.x) {
        const int k = kl / n, l = kl - k * n;
        dst[kl] *= (mij + (double)(m[k] + m[l])) * 0.25;
    }
}

// max |x| of a real array (non-negative doubles order like their bit patterns)
__global__ void max_abs_kernel(const double* __restrict__ x, long long n, unsigned long long* out) {
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, fabs(x[i]));
    warp_atomic_max_abs(m, out);
}

// unpack a packed symmetric vector to a full (n, n) matrix
__global__ void unpack_sym_kernel(const double* __restrict__ packed, double* __restrict__ full, int n) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        full[idx] = packed[tri_idx(r, c)];
    }
}

// ----------------------------------------------------------------------------------------------------------
// res[x] = scale * Re sum_k in[k][x]  (+ max|Im|)  -- the sum over k-points of transform_trans_inv_k
// (libdmet/routine/slater_helper.py:37-50): one thread per x, fixed summation order.
// ----------------------------------------------------------------------------------------------------------
__global__ void ksum_real_kernel(const double2* __restrict__ in, double* __restrict__ out, int nk, long long X,
                                 double scale, unsigned long long* imag_max) {
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double im_max = 0.0;
    if (x < X) {
        double re = 0.0, im = 0.0;
        for (int k = 0; k < nk; ++k) {
            const double2 v = in[(size_t)k * X + x];
            re += v.x;
            im += v.y;
        }
        out[x] = re * scale;
        im_max = fabs(im);
    }
    if (imag_max) warp_atomic_max_abs(im_max, imag_max);
}

// ----------------------------------------------------------------------------------------------------------
// Synthetic GDF generator (benchmarks and parity tests; the host twin is libdmet_preview_b200/synthetic.py).
//   L(ki,kj)[L,p,q] = scale/4 * ( u_ij[L,p,q] + conj(u_ji[L,q,p]) + conj(u_-i-j[L,p,q]) + u_-j-i[L,q,p] )
// which satisfies L(kj,ki) = L(ki,kj)^H and L(-ki,-kj) = conj L(ki,kj) by construction; u is a counter-based
// 32-bit hash mapped to [-1, 1) so host and device produce identical bits.
// ----------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ double synth_u(uint32_t key, uint32_t idx) {
    return (double)(int32_t)lowbias32(idx ^ key) * 4.656612873077393e-10;   // 2^-31
}
__global__ void __launch_bounds__(256)
synth_block_kernel(double2* __restrict__ out, int naux, int nao, int aux_offset, uint32_t key_ij, uint32_t key_ji,
                   uint32_t key_mij, uint32_t key_mji, double scale) {
    const size_t total = (size_t)naux * nao * nao;
    const double s = 0.25 * scale;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = (uint32_t)(e % nao);
        const uint32_t lp = (uint32_t)(e / nao);
        const uint32_t p = lp % nao, L = lp / nao + (uint32_t)aux_offset;
        const uint32_t d = 2u * ((L * nao + p) * nao + q);                // (L, p, q)
        const uint32_t tt = 2u * ((L * nao + q) * nao + p);               // (L, q, p)
        const double re = synth_u(key_ij, d) + synth_u(key_ji, tt) + synth_u(key_mij, d) + synth_u(key_mji, tt);
        const double im = synth_u(key_ij, d + 1) - synth_u(key_ji, tt + 1) - synth_u(key_mij, d + 1) +
                          synth_u(key_mji, tt + 1);
        out[e] = make_double2(s * re, s * im);
    }
}

}  // namespace ldm
