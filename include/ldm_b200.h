/* ldm_b200.h -- C ABI of libldm_b200.so, the sm_100a implementation of libDMET's embedding-Hamiltonian hot path.
 *
 * The reference (gkclab/libdmet_preview) has no FFI layer: the seam is a set of module-level Python functions
 * whose arithmetic runs in PySCF's C libraries.  Each entry point below names the reference routine (file:line
 * under /root/reference) whose work it takes over; the Python modules of `libdmet_preview_b200` bind them with
 * ctypes and re-expose the reference's Python signatures.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; ldm_last_error() gives the message
 *     (thread-local);
 *   - all matrices are C-contiguous; complex data are interleaved (re, im) doubles ("double2");
 *   - `*_d` pointers are DEVICE pointers, `*_h` pointers are HOST pointers; nothing else crosses the ABI;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls are asynchronous on
 *     that stream unless stated otherwise;
 *   - the caller owns every buffer it passes in; the handle owns its internal workspaces.
 */
#ifndef LDM_B200_H
#define LDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldm_context* ldm_handle;

/* ---- library / handle ---------------------------------------------------------------------------------- */
int ldm_version(void);
const char* ldm_last_error(void);
/* binds the handle to CUDA device `device` (one handle per GPU / per process rank) */
int ldm_create(int device, ldm_handle* out);
int ldm_destroy(ldm_handle h);
/* run-time options of a handle.  "zgemm_3m": 1 (default; environment LDM_ZGEMM_3M=0 changes the default) evaluates
 * complex products on the tensor cores with three real multiplications -- k1 = (Ar + Ai) Br, k2 = Ar (Bi - Br),
 * k3 = Ai (Br + Bi), Re = k1 - k3, Im = k1 + k2, as in csrc/zgemm_tn.cuh -- 0 with the classical four.  Takes effect for the next ldm_zgemm_tn / ldm_eri_begin.  Returns the previous value in
 * *old_value when that pointer is not NULL. */
int ldm_set_option(ldm_handle h, const char* name, int value, int* old_value);
/* pinned host memory for staging GDF blocks (cudaHostAlloc / cudaFreeHost) */
int ldm_host_alloc(size_t bytes, void** out_h);
int ldm_host_free(void* p_h);
/* device memory for callers that do not bring their own allocator (cudaMalloc / cudaFree / cudaMemcpyAsync) */
int ldm_dev_alloc(ldm_handle h, size_t bytes, void** out_d);
int ldm_dev_free(ldm_handle h, void* p_d);
int ldm_memcpy_h2d(ldm_handle h, void* dst_d, const void* src_h, size_t bytes, void* stream);
int ldm_memcpy_d2h(ldm_handle h, void* dst_h, const void* src_d, size_t bytes, void* stream);
int ldm_memset(ldm_handle h, void* dst_d, int value, size_t bytes, void* stream);
int ldm_stream_sync(ldm_handle h, void* stream);

/* ---- complex TN GEMM on the FP64 tensor cores -----------------------------------------------------------
 * C[b](r, c) (+)= alpha * sum_{s<nseg} sum_{k<K} opA(A[az(b,s)](r, k)) * opB(B[bz(b,s)](c, k))
 * A: (za_count, M, K) complex, B: (zb_count, N, K) complex, both k-contiguous.
 * segs_h: nbatch*nseg records {int az, int bz, int conjA, int conjB} (host; copied internally).
 * Output element (r, c) of batch b goes to C_d[c_off_h[b] + (r / rdiv) * s_outer + (r % rdiv) * s_inner + c * s_col]
 * (complex elements).  Replaces PySCF _ao2mo.r_e2 (libdmet/basis_transform/eri_transform.py:432-433), the numpy
 * dot loops of make_basis.py:548-557 (transform_h1_to_lo), misc.py:58 (kdot) and slater_helper.py:46.        */
int ldm_zgemm_tn(ldm_handle h, void* stream, const void* A_d, int za_count, const void* B_d, int zb_count, int M,
                 int N, int K, int nseg, int nbatch, const int32_t* segs_h, void* C_d, const int64_t* c_off_h,
                 int rdiv, int64_t s_outer, int64_t s_inner, int64_t s_col, double alpha, int accumulate);

/* ---- real TN GEMM / SYRK on the FP64 tensor cores -------------------------------------------------------
 * C(r, c) (+)= alpha * sum_{k<K} A(r, k) * B(c, k);  A: (M, lda), B: (N, ldb) row-major doubles; lower_only=1
 * computes only the tiles on or below the diagonal (A == B, syrk): the lower triangle is valid afterwards, elements
 * above it are either computed too (inside a diagonal tile) or left as they were.  Tiles are 128 x 128, or 64 x 64
 * when the product has fewer 128 x 128 tiles than the device has SMs.  Replaces PySCF lib.dot in `_Lij_s4_to_eri`
 * (eri_transform.py:450-485).                                                                               */
int ldm_dgemm_tn(ldm_handle h, void* stream, const double* A_d, int64_t lda, const double* B_d, int64_t ldb, int M,
                 int N, int K, double* C_d, int64_t ldc, double alpha, int accumulate, int lower_only);
/* C(r, c) = C(c, r) for c > r  (fills the upper triangle after lower_only products) */
int ldm_mirror_lower(ldm_handle h, void* stream, double* C_d, int n, int64_t ldc);

/* ---- lattice Fourier transforms --------------------------------------------------------------------------
 * out[b][k][x] = scale * sum_R W[k][R] in[b][R][x],  W: (nout, nin) complex phase matrix on the device.
 * in_real: input is real doubles;  out_real: keep the real part only and return max|imag| in *imag_max_h
 * (synchronises the stream in that case).  Replaces scipy fftn/ifftn in libdmet/system/fourier.py:160-177
 * (FFTtoK / FFTtoT behind R2k / k2R, fourier.py:129-158) and the einsum of eri_transform.py:125.             */
int ldm_phase_transform(ldm_handle h, void* stream, const void* in_d, void* out_d, const void* W_d, int nin,
                        int nout, int64_t X, int batch, double scale, int in_real, int out_real,
                        double* imag_max_h);
/* The same transform for the k-points of the mesh itself, factorised over the mesh axes (kmesh3 = {n0,n1,n2}, each
 * 1..8): out[b][k][x] = scale * sum_R exp(-+2 pi i k.R) in[b][R][x], forward != 0 -> minus sign (R2k / FFTtoK),
 * forward == 0 -> plus sign (k2R / FFTtoT, pass scale = 1/Nk).  HBM-bound.                                    */
int ldm_lattice_dft(ldm_handle h, void* stream, const void* in_d, void* out_d, const int32_t* kmesh3, int64_t X,
                    int batch, int forward, double scale, int in_real, int out_real, double* imag_max_h);
/* out[b][c][r] = scale * op(in[b][r][c]) for complex matrices (op = conj if conj != 0) */
int ldm_ztranspose(ldm_handle h, void* stream, const void* in_d, void* out_d, int batch, int rows, int cols,
                   int conj, double scale);
int ldm_d2z(ldm_handle h, void* stream, const double* in_d, void* out_d, int64_t n);
/* out[x] = scale * Re sum_k in[k][x]; max|Im sum| -> *imag_max_h (synchronises).  The k-sum of
 * transform_trans_inv_k (libdmet/routine/slater_helper.py:37-50).                                            */
int ldm_ksum_real(ldm_handle h, void* stream, const void* in_d, double* out_d, int nk, int64_t X, double scale,
                  double* imag_max_h);
/* out[L][m(m+1)/2 + c] = in[L][m][c], c <= m, for a stack of `rows` complex (n, n) matrices: PySCF lib.pack_tril
 * as transform_gdf_to_lo applies it to the blocks it stores (eri_transform.py:1386-1390).  out_real: keep the real
 * part only (double output) and report max|imag| (both k-points Gamma).                                          */
int ldm_pack_tril(ldm_handle h, void* stream, const void* in_d, void* out_d, int rows, int n, int out_real,
                  double* imag_max_h);

/* ---- ERI re-layouts and J/K ------------------------------------------------------------------------------
 * restore: s4 (npair, npair) -> s1 (n,n,n,n) or s8 (npair(npair+1)/2)   (pyscf ao2mo.restore; call sites
 * eri_transform.py:529,543).  jk: vj (n,n), vk (n,n) from an s4 ERI and one density matrix (PySCF
 * hf.dot_eri_dm; call site libdmet/solver/scf.py:300-326).  vk_d may be NULL.                                */
int ldm_restore_s1(ldm_handle h, void* stream, const double* eri4_d, double* out_d, int n);
int ldm_restore_s8(ldm_handle h, void* stream, const double* eri4_d, double* out_d, int n);
int ldm_jk_s4(ldm_handle h, void* stream, const double* eri4_d, const double* dm_d, double* vj_d, double* vk_d,
              int n);
/* Same result for a SYMMETRIC s4 block (eri4[P][Q] == eri4[Q][P]: the restricted / aa / bb blocks) and a symmetric
 * density matrix -- what the reference always passes (hermi=1, solver/scf.py:300, 309, 313-316): only the lower
 * triangle of eri4 is read, half of the tensor crosses HBM.  Shapes outside the kernel's range (n <= 16, n > 160)
 * are served by ldm_jk_s4.                                                                                     */
int ldm_jk_s4_symm(ldm_handle h, void* stream, const double* eri4_d, const double* dm_d, double* vj_d, double* vk_d,
                   int n);

/* DMET energy weights, in place (reference: get_H2_scaled, libdmet/routine/slater.py:1734-1778).
 * symmetry 4: eri (npair, npair) *= (w[P] + w[Q]) / 4 with weights_d[npair] = impurity count of each pair (0,1,2);
 * symmetry 1: eri (n,n,n,n) *= (m[i]+m[j]+m[k]+m[l]) / 4 with weights_d[n] = 0/1 impurity flags.              */
int ldm_scale_eri(ldm_handle h, void* stream, double* eri_d, int n, int symmetry, const int32_t* weights_d);

/* ---- synthetic GDF block generator -----------------------------------------------------------------------
 * Writes rows [aux_offset, aux_offset + naux) of L(k_i,k_j) (.., nao, nao) complex for the seeded synthetic
 * provider (host twin: libdmet_preview_b200/synthetic.py); keys are the four 32-bit pair keys of that scheme. */
int ldm_synth_block(ldm_handle h, void* stream, void* out_d, int naux, int nao, int aux_offset, uint32_t key_ij,
                    uint32_t key_ji, uint32_t key_mij, uint32_t key_mji, double scale);

/* ---- embedding ERI from GDF blocks: the get_emb_eri_fast_gdf pipeline ------------------------------------
 * Replaces the loop nest of libdmet/basis_transform/eri_transform.py:338-386 (transform_ao_to_emb, hermi_sum,
 * pack_tril, accumulation, _Lij_s4_to_eri).  The k-point schedule (which (i, j) blocks belong to which transfer
 * momentum kL, which are symmetrised, the weights) is replayed on the host by the caller and fed block by block:
 *
 *   ldm_eri_begin(h, ...)                      once; CT_d = (nspin, nkpts, neo, nao) complex = C_ao_emb^T
 *   for kL in schedule:
 *       for (i, j, sym) in blocks(kL):  ldm_eri_block_*(h, i, j, sym, ...)
 *       ldm_eri_end_kl(h, weight)               weight 1: eri += Re^T Re ; weight 2: eri += 2 (Re^T Re + Im^T Im)
 *                                               weight 0: complex Lambda^dagger Lambda, real part (no time reversal)
 *   ldm_eri_finish(h)                           flushes pending products; eri_d then holds the LOWER triangles of
 *                                               the aa, (ab,) (bb) s4 blocks in the reference's incore order
 *   ldm_eri_end(h)                              closes the build (workspaces stay cached in the handle)
 *
 * eri_d: (nspin*(nspin+1)/2, npair, npair) doubles owned by the caller, accumulated in place (zero it first;
 * partial results of several ranks can be summed with NCCL before ldm_mirror_lower).
 * naux may be a slice of the auxiliary index: rows of Lambda are independent through stage 1 and additive in
 * stage 3, so (kL, aux-range) work items can be dealt to different GPUs and their eri_d summed.
 * max_group: how many (i, j) blocks are staged and processed per kernel launch (>= 1).
 * Block sources: a host pointer (pageable or pinned; consumed before the call returns), a slot of a resident
 * device store registered with ldm_eri_set_store, or the synthetic generator.                                */
int ldm_eri_begin(ldm_handle h, void* stream, int nkpts, int nao, int naux, int neo, int nspin, const void* CT_d,
                  double* eri_d, int max_group, int kl_group);
int ldm_eri_set_store(ldm_handle h, const void* store_d, int nslots);
/* Optional diagnostic of the path without time reversal (weight-0 units): accumulate
 * Im(Lambda^dagger Lambda) into imag_d (same shape as eri_d, zeroed by the caller), whose max-abs the reference logs
 * and warns about above ERI_IMAG_TOL (eri_transform.py:390-395).  Call after ldm_eri_begin.                   */
int ldm_eri_set_imag(ldm_handle h, double* imag_d);
/* max |x_i| of a device array -> *out_h (synchronises the stream) */
int ldm_max_abs(ldm_handle h, void* stream, const double* x_d, int64_t n, double* out_h);
/* gso = 1 (nspin must be 2): the two "spins" are the alpha / beta halves of generalised spin orbitals and ONE ERI
 * block eri_d (1, npair, npair) is built from Lambda_a - Lambda_b  (reference: get_emb_eri_gso / _Lij_s4_to_eri_gso,
 * eri_transform.py:1104-1284).  Call right after ldm_eri_begin.                                              */
int ldm_eri_set_mode(ldm_handle h, int gso);
int ldm_eri_block_host(ldm_handle h, int ki, int kj, int sym, const void* L_h);
/* A block as it is STORED in a PySCF cderi file (reference: sr_loop -> PySCF _load3c + lib.unpack_tril + cast,
 * libdmet/basis_transform/eri_transform.py:195-227): src_h is the (rows, ncols) C-contiguous entry, complex128 or --
 * LDM_STORED_REAL -- float64; ncols = nao*nao, or nao*(nao+1)/2 for a Hermitian-packed k_i == k_j entry;
 * LDM_STORED_SWAPPED: the entry is the one of the pair (k_j, k_i) and is conjugate-transposed in its two orbital
 * indices (plain conjugate when packed); rows < naux of the build leaves the remaining auxiliary rows zero.  The
 * raw bytes cross PCIe once and are widened / unpacked / transposed on the device (ldm_unpack_stored).        */
#define LDM_STORED_SWAPPED 1
#define LDM_STORED_REAL 2
#define LDM_STORED_CONJ 4      /* the entry is the one of the time-reversed pair (-k_i, -k_j): plain conjugate */
int ldm_eri_block_stored(ldm_handle h, int ki, int kj, int sym, const void* src_h, int rows, int64_t ncols,
                         int flags);
/* device -> device form of the same conversion: src_d (rows, ncols) -> out_d (naux, nao, nao) complex128 */
int ldm_unpack_stored(ldm_handle h, void* stream, const void* src_d, void* out_d, int naux, int rows, int nao,
                      int64_t ncols, int flags);
int ldm_eri_block_store(ldm_handle h, int ki, int kj, int sym, int slot);
int ldm_eri_block_synth(ldm_handle h, int ki, int kj, int sym, int aux_offset, uint32_t key_ij, uint32_t key_ji,
                        uint32_t key_mij, uint32_t key_mji, double scale);
int ldm_eri_end_kl(ldm_handle h, int weight);
int ldm_eri_finish(ldm_handle h);
int ldm_eri_end(ldm_handle h);
/* the pipeline's device workspaces are kept in the handle between builds (grow-only); this frees them */
int ldm_release_workspaces(ldm_handle h);
/* counters since ldm_eri_begin: kernels launched by this library, bytes copied host->device */
int ldm_eri_stats(ldm_handle h, int64_t* launches, int64_t* h2d_bytes);
/* total kernels launched through this handle since creation */
int64_t ldm_launch_count(ldm_handle h);
/* time between two internal events around the last dominant-kernel launches (ms): kind 0 = stage-1 zgemm,
 * 1 = stage-3 dgemm; returns accumulated ms and number of launches since ldm_eri_begin (synchronises)       */
int ldm_eri_kernel_time(ldm_handle h, int kind, double* ms, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* LDM_B200_H */
