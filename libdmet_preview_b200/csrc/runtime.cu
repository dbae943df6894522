// libldm_b200.so -- host runtime and C ABI (include/ldm_b200.h) around the sm_100a kernels.
// Native code on purpose: the block scheduler, the staging ring, the tensor-map set-up and the kernel launches of
// the embedding-ERI pipeline all live here; Python only replays the k-point schedule and hands over pointers.
#include "../../include/ldm_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "aux_kernels.cuh"
#include "common.cuh"
#include "dgemm_tn.cuh"
#include "zgemm_tn.cuh"

namespace ldm {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

// ----------------------------------------------------------------------------------------------------------
// tensor maps
// ----------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int encode_tmap_f64_3d(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                       int swizzle) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return -3;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15)) {
        set_error("tensor map: base and strides must be 16-byte aligned");
        return -2;
    }
    cuuint64_t dims[3] = {dim0, dim1, dim2 ? dim2 : 1};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes ? stride2_bytes : stride1_bytes * dim1};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " dims=(" +
                  std::to_string(dim0) + "," + std::to_string(dim1) + "," + std::to_string(dim2) + ") strides=(" +
                  std::to_string(stride1_bytes) + "," + std::to_string(stride2_bytes) + ") box=(" +
                  std::to_string(box0) + "," + std::to_string(box1) + ")");
        return -3;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------------------
// zgemm tile configurations
// ----------------------------------------------------------------------------------------------------------
struct ZConfig {
    int BM, BN, threads, smem;
    void (*kernel)(const CUtensorMap, const CUtensorMap, const ZGemmArgs);
    bool m3;
    int jp;                      // 3M: column fragments per group of the main loop (zgemm_tn.cuh)
    bool wide;                   // 64 x (8 FB) tile, one 8-row fragment per warp
};

template <int WM, int WN, int FA, int FB, bool M3 = false, int JP = 1>
static ZConfig make_zconfig() {
    using T = ZTile<WM, WN, FA, FB, M3>;
    return ZConfig{T::BM, T::BN, T::THREADS, T::SMEM, zgemm_tn_kernel<WM, WN, FA, FB, M3, JP>, M3, JP,
                   WM == 8 && WN == 1 && FA == 1};
}

template <int FB, bool M3 = false, int JP = 1>
static void add_wide(std::vector<ZConfig>& v) {
    v.push_back(make_zconfig<8, 1, 1, FB, M3, JP>());
    if constexpr (FB > 1) add_wide<FB - 1, M3, JP>(v);
}

// Tile family.  First the register-blocked 32x(8*FB) warp tiles (least shared-memory traffic), then the "wide"
// family 64 x (8*FB) with one 8-row fragment per warp and FB up to 25 column fragments, which removes N padding
// for any neo <= 200 (neo = 150 runs as 64x152 instead of 64x160).
static const std::vector<ZConfig>& zconfigs() {
    static std::vector<ZConfig> v = [] {
        std::vector<ZConfig> c = {
            make_zconfig<2, 4, 4, 5>(), make_zconfig<4, 2, 4, 5>(), make_zconfig<8, 1, 4, 5>(),
            make_zconfig<2, 4, 4, 4>(), make_zconfig<4, 2, 4, 4>(), make_zconfig<8, 1, 4, 4>(),
            make_zconfig<2, 4, 4, 3>(), make_zconfig<4, 2, 4, 3>(), make_zconfig<8, 1, 4, 3>(),
        };
        add_wide<25>(c);
        // 3-multiplication family (three accumulator sets): 64 x (8*FB) wide tiles up to FB = 13 (measured 2 % faster
        // than the register-blocked tiles of equal shape), then 64x80 / 64x96 / 64x64 with two row fragments per warp
        add_wide<13, true>(c);
        // the wide 3M family again with the column fragments of the main loop taken two at a time (k-step loop outside
        // the pair: DMMAs on the same accumulator are 6 instead of 3 instructions apart; a lone warp issues
        // back-to-back dependent DMMAs 26 clocks apart instead of 16, tools/dmma_probe.cu).  LDM_Z3M_JP=1 selects the
        // single-fragment order.
        add_wide<13, true, 2>(c);
        c.push_back(make_zconfig<4, 2, 2, 5, true>());
        c.push_back(make_zconfig<4, 2, 2, 6, true>());
        c.push_back(make_zconfig<4, 2, 2, 4, true>());
        return c;
    }();
    return v;
}

static const ZConfig& pick_zconfig(int N, bool m3) {
    const auto& v = zconfigs();
    if (const char* f = getenv("LDM_FORCE_ZCFG")) {       // development aid: force a tile configuration
        int idx = atoi(f);
        if (idx >= 0 && idx < (int)v.size()) return v[idx];
    }
    // cost of a configuration: padded N, inflated by a per-tile overhead that shrinks with the tile width (a 64x8
    // tile pads N = 150 to 152 but spends its time in fragment loads and epilogues); ties: the earlier (more
    // register-blocked) entry wins
    static const int want_jp = getenv("LDM_Z3M_JP") ? atoi(getenv("LDM_Z3M_JP")) : 2;
    int best = -1;
    double best_cost = 0.0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].m3 != m3) continue;
        if (m3 && v[i].wide && v[i].jp != want_jp) continue;   // wide family: one order
        const int tiles = (N + v[i].BN - 1) / v[i].BN;
        double pad = (double)tiles * v[i].BN;
        // 3M tiles run their last n-tile with one column fragment less when that fragment would be all padding
        if (v[i].m3 && v[i].wide && tiles > 1 && v[i].BN > 8 && (N - (tiles - 1) * v[i].BN + 7) / 8 == v[i].BN / 8 - 1)
            pad -= 8.0;
        const double cost = pad * (1.0 + 24.0 / v[i].BN);
        if (best < 0 || cost < best_cost) {
            best = (int)i;
            best_cost = cost;
        }
    }
    return v[best];
}

}  // namespace ldm

using namespace ldm;

// ----------------------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------------------
struct EriPlan;

struct ldm_context {
    int device = 0;
    int num_sms = 148;
    int64_t launches = 0;
    // small reusable device scratch for segment tables / offsets
    void* scratch_d = nullptr;
    void* scratch_h = nullptr;   // pinned twin of scratch_d: tables are written here, then copied asynchronously
    size_t scratch_bytes = 0;
    size_t scratch_used = 0;
    cudaEvent_t scratch_ev = nullptr;
    unsigned long long* imag_d = nullptr;
    void* jk_part_d = nullptr;
    size_t jk_part_bytes = 0;
    EriPlan* plan = nullptr;
    bool attrs_set = false;
    bool zgemm_3m = true;        // complex products with three real multiplications (see zgemm_tn.cuh)
    // grow-only workspace pool of the ERI pipeline (X, S_sym, S_pln, panel, ring): cudaMalloc/cudaFree of GB-sized
    // buffers costs tens of ms per build and cudaFree synchronises the device, so they are kept across builds
    void* ws[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t ws_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t bforms_ev = nullptr;   // last use of the WS_BFORMS planes (ldm_zgemm_tn may be called on any stream)
    cudaStream_t copy_st = nullptr;    // host -> device staging stream of the ERI pipeline, created once per handle
};

enum { WS_XT = 0, WS_SSYM = 1, WS_SPLN = 2, WS_PANEL = 3, WS_RING = 4, WS_BFORMS = 5, WS_CTFORMS = 6, WS_STORED = 7,
       WS_COUNT = 8 };

static int ws_get(ldm_handle h, int slot, size_t bytes, void** out) {
    if (h->ws_bytes[slot] < bytes) {
        if (h->ws[slot]) {
            LDM_CUDA_OK(cudaDeviceSynchronize());
            LDM_CUDA_OK(cudaFree(h->ws[slot]));
            h->ws[slot] = nullptr;
            h->ws_bytes[slot] = 0;
        }
        LDM_CUDA_OK(cudaMalloc(&h->ws[slot], bytes));
        h->ws_bytes[slot] = bytes;
    }
    *out = h->ws[slot];
    return 0;
}

static void ws_release(ldm_handle h) {
    for (int i = 0; i < WS_COUNT; ++i) {
        if (h->ws[i]) cudaFree(h->ws[i]);
        h->ws[i] = nullptr;
        h->ws_bytes[i] = 0;
    }
}

static int ensure_attrs(ldm_handle h) {
    if (h->attrs_set) return 0;
    for (const auto& c : zconfigs())
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)c.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)dgemm_tn_kernel<DTileS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     DTileS::SMEM));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)dgemm_tn_kernel<DTile>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     DTile::SMEM));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     200 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     200 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_rows_bulk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     210 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_rows_bulk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     210 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_rows_bulk_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     210 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_rows_bulk_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     210 * 1024));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)pack_sym_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 16 * 17 * 16));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)pack_sym_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 16 * 17 * 16));
    LDM_CUDA_OK(cudaFuncSetAttribute((const void*)restore_s1_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    h->attrs_set = true;
    return 0;
}

// Segment tables / offset tables are tiny.  Each launch writes its table into a fresh region of a pinned host
// buffer and copies it asynchronously to the same offset of the device twin (a pageable source would make
// cudaMemcpyAsync synchronise the stream); the regions wrap after a device sync.
static int scratch_put(ldm_handle h, cudaStream_t st, const void* src, size_t bytes, void** out) {
    const size_t padded = (bytes + 255) & ~size_t(255);
    if (!h->scratch_d) {
        h->scratch_bytes = 8 << 20;
        LDM_CUDA_OK(cudaMalloc(&h->scratch_d, h->scratch_bytes));
        LDM_CUDA_OK(cudaHostAlloc(&h->scratch_h, h->scratch_bytes, cudaHostAllocDefault));
    }
    LDM_REQUIRE(padded <= h->scratch_bytes, "segment table too large");
    if (h->scratch_used + padded > h->scratch_bytes) {
        LDM_CUDA_OK(cudaDeviceSynchronize());   // rare: everything that used older tables has finished
        h->scratch_used = 0;
    }
    *out = static_cast<char*>(h->scratch_d) + h->scratch_used;
    void* stage = static_cast<char*>(h->scratch_h) + h->scratch_used;
    h->scratch_used += padded;
    std::memcpy(stage, src, bytes);
    LDM_CUDA_OK(cudaMemcpyAsync(*out, stage, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

extern "C" {

int ldm_version(void) { return 100; }
const char* ldm_last_error(void) { return ldm::last_error(); }

int ldm_create(int device, ldm_handle* out) {
    LDM_REQUIRE(out != nullptr, "out");
    int count = 0;
    LDM_CUDA_OK(cudaGetDeviceCount(&count));
    LDM_REQUIRE(device >= 0 && device < count, "device index out of range");
    LDM_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    LDM_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("libldm_b200 is built for sm_100a only; device is ") + prop.name + " (sm_" +
                  std::to_string(prop.major) + std::to_string(prop.minor) + ")");
        return -4;
    }
    ldm_context* h = new ldm_context();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("LDM_ZGEMM_3M")) h->zgemm_3m = atoi(e) != 0;
    LDM_CUDA_OK(cudaMalloc(&h->imag_d, sizeof(unsigned long long)));
    *out = h;
    return ensure_attrs(h);
}

int ldm_eri_end(ldm_handle h);

int ldm_destroy(ldm_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->plan) ldm_eri_end(h);
    ws_release(h);
    if (h->scratch_d) cudaFree(h->scratch_d);
    if (h->scratch_h) cudaFreeHost(h->scratch_h);
    if (h->imag_d) cudaFree(h->imag_d);
    if (h->bforms_ev) cudaEventDestroy(h->bforms_ev);
    if (h->copy_st) cudaStreamDestroy(h->copy_st);
    if (h->jk_part_d) cudaFree(h->jk_part_d);
    delete h;
    return 0;
}

int ldm_set_option(ldm_handle h, const char* name, int value, int* old_value) {
    LDM_REQUIRE(h && name, "null pointer");
    if (std::strcmp(name, "zgemm_3m") == 0) {
        if (old_value) *old_value = h->zgemm_3m ? 1 : 0;
        h->zgemm_3m = value != 0;
        return 0;
    }
    set_error(std::string("unknown option: ") + name);
    return -2;
}

int ldm_host_alloc(size_t bytes, void** out_h) {
    LDM_CUDA_OK(cudaHostAlloc(out_h, bytes, cudaHostAllocDefault));
    return 0;
}
int ldm_host_free(void* p_h) {
    LDM_CUDA_OK(cudaFreeHost(p_h));
    return 0;
}
int ldm_dev_alloc(ldm_handle h, size_t bytes, void** out_d) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    LDM_CUDA_OK(cudaMalloc(out_d, bytes));
    return 0;
}
int ldm_dev_free(ldm_handle h, void* p_d) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    LDM_CUDA_OK(cudaFree(p_d));
    return 0;
}
int ldm_memcpy_h2d(ldm_handle h, void* dst_d, const void* src_h, size_t bytes, void* stream) {
    LDM_CUDA_OK(cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}
int ldm_memcpy_d2h(ldm_handle h, void* dst_h, const void* src_d, size_t bytes, void* stream) {
    LDM_CUDA_OK(cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return 0;
}
int ldm_memset(ldm_handle h, void* dst_d, int value, size_t bytes, void* stream) {
    LDM_CUDA_OK(cudaMemsetAsync(dst_d, value, bytes, (cudaStream_t)stream));
    return 0;
}
int ldm_stream_sync(ldm_handle h, void* stream) {
    LDM_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int64_t ldm_launch_count(ldm_handle h) { return h ? h->launches : 0; }

}  // extern "C"

// ----------------------------------------------------------------------------------------------------------
// GEMM launchers (internal; tensor maps supplied by the caller so the ERI pipeline can cache them)
// ----------------------------------------------------------------------------------------------------------
static int launch_zgemm(ldm_handle h, cudaStream_t st, const ZConfig& cfg, const CUtensorMap& tmA,
                        const CUtensorMap& tmB, int M, int N, int K, int nseg, int nbatch, const ZSeg* segs_h,
                        double2* C, const long long* c_off_h, int rdiv, long long s_outer, long long s_inner,
                        long long s_col, double alpha, int accumulate) {
    if (M <= 0 || N <= 0 || nbatch <= 0) return 0;
    void* segs_d = nullptr;
    void* off_d = nullptr;
    int rc = scratch_put(h, st, segs_h, sizeof(ZSeg) * (size_t)nseg * nbatch, &segs_d);
    if (rc) return rc;
    if (c_off_h) {
        rc = scratch_put(h, st, c_off_h, sizeof(long long) * (size_t)nbatch, &off_d);
        if (rc) return rc;
    }
    ZGemmArgs a;
    a.M = M; a.N = N; a.K = K; a.nseg = nseg; a.nbatch = nbatch;
    a.segs = static_cast<const ZSeg*>(segs_d);
    a.C = C;
    a.c_off = static_cast<const long long*>(off_d);
    a.rdiv = rdiv > 0 ? rdiv : 1;
    a.s_outer = s_outer; a.s_inner = s_inner; a.s_col = s_col;
    a.alpha = alpha; a.accumulate = accumulate;
    a.tiles_m = (M + cfg.BM - 1) / cfg.BM;
    a.tiles_n = (N + cfg.BN - 1) / cfg.BN;
    long long ntiles = (long long)a.tiles_m * a.tiles_n * nbatch;
    LDM_REQUIRE(ntiles < (1ll << 31), "too many tiles");
    int grid = (int)std::min<long long>(ntiles, h->num_sms);
    // the last n-tile needs one column fragment less than the others: run it with FB - 1 fragments (zgemm_tn.cuh)
    const int fb_full = cfg.BN / 8, fb_last = (N - (a.tiles_n - 1) * cfg.BN + 7) / 8;
    static const bool allow_short = getenv("LDM_ZGEMM_SHORT_LAST") ? atoi(getenv("LDM_ZGEMM_SHORT_LAST")) != 0 : true;
    a.short_last = (allow_short && cfg.m3 && cfg.wide && fb_full > 1 && fb_last == fb_full - 1 && a.tiles_n > 1 &&
                    grid % a.tiles_n == 0) ? 1 : 0;
    cfg.kernel<<<grid, cfg.threads, cfg.smem, st>>>(tmA, tmB, a);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

// B operand of a 3M launch: build the five real planes in workspace `slot` and encode their tensor map
static int make_bforms(ldm_handle h, cudaStream_t st, int slot, const void* B_d, int zb_count, int N, int K, int BN,
                       CUtensorMap* tm) {
    const int Kp = (K + 7) & ~7;        // zero-padded rows of whole 8-k groups (permuted order, see zforms_kernel)
    void* F = nullptr;
    int rc = ws_get(h, slot, (size_t)ZFORM_PLANES * zb_count * N * Kp * 8, &F);
    if (rc) return rc;
    const long long rows = (long long)zb_count * N;
    const long long total = rows * Kp;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16);
    zforms_kernel<<<grid, 256, 0, st>>>(static_cast<const double2*>(B_d), static_cast<double*>(F), rows, K, Kp, N);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return encode_tmap_f64_3d(tm, F, (uint64_t)Kp, (uint64_t)N, (uint64_t)ZFORM_PLANES * zb_count, 8ull * Kp,
                              8ull * Kp * N, 8, BN, 2);
}

template <class T>
static int launch_dgemm_t(ldm_handle h, cudaStream_t st, const double* A, long long lda, const double* B, long long ldb,
                          int M, int N, int K, double* C, long long ldc, double alpha, int accumulate, int lower_only) {
    CUtensorMap tmA, tmB;
    int rc = encode_tmap_f64_3d(&tmA, A, (uint64_t)K, (uint64_t)M, 1, (uint64_t)lda * 8, 0, 16, T::BM, true);
    if (rc) return rc;
    rc = encode_tmap_f64_3d(&tmB, B, (uint64_t)K, (uint64_t)N, 1, (uint64_t)ldb * 8, 0, 16, T::BN, true);
    if (rc) return rc;
    DGemmArgs a;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.alpha = alpha; a.accumulate = accumulate;
    a.lower_only = lower_only;
    a.tiles_m = (M + T::BM - 1) / T::BM;
    a.tiles_n = (N + T::BN - 1) / T::BN;
    long long ntiles = lower_only ? (long long)a.tiles_m * (a.tiles_m + 1) / 2 : (long long)a.tiles_m * a.tiles_n;
    int grid = (int)std::min<long long>(ntiles, h->num_sms);
    dgemm_tn_kernel<T><<<grid, T::THREADS, T::SMEM, st>>>(tmA, tmB, a);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

static int launch_dgemm(ldm_handle h, cudaStream_t st, const double* A, long long lda, const double* B, long long ldb,
                        int M, int N, int K, double* C, long long ldc, double alpha, int accumulate, int lower_only) {
    if (M <= 0 || N <= 0) return 0;
    LDM_REQUIRE(!lower_only || M == N, "lower_only needs a square product");
    // fewer 128 x 128 tiles than SMs: 64 x 64 tiles occupy the machine (four times as many, each a quarter of the work)
    const long long tm = (M + DTile::BM - 1) / DTile::BM, tn = (N + DTile::BN - 1) / DTile::BN;
    const long long big = lower_only ? tm * (tm + 1) / 2 : tm * tn;
    static const int force = getenv("LDM_DGEMM_TILE") ? atoi(getenv("LDM_DGEMM_TILE")) : 0;   // development aid: 64 / 128
    const bool small = force ? force == 64 : big < h->num_sms;
    if (small) return launch_dgemm_t<DTileS>(h, st, A, lda, B, ldb, M, N, K, C, ldc, alpha, accumulate, lower_only);
    return launch_dgemm_t<DTile>(h, st, A, lda, B, ldb, M, N, K, C, ldc, alpha, accumulate, lower_only);
}

extern "C" {

int ldm_zgemm_tn(ldm_handle h, void* stream, const void* A_d, int za_count, const void* B_d, int zb_count, int M,
                 int N, int K, int nseg, int nbatch, const int32_t* segs_h, void* C_d, const int64_t* c_off_h,
                 int rdiv, int64_t s_outer, int64_t s_inner, int64_t s_col, double alpha, int accumulate) {
    LDM_REQUIRE(h && A_d && B_d && C_d && segs_h, "null pointer");
    LDM_REQUIRE(M > 0 && N > 0 && K > 0 && nseg > 0 && nbatch > 0, "shape");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    const ZConfig& cfg = pick_zconfig(N, h->zgemm_3m);
    CUtensorMap tmA, tmB;
    int rc = encode_tmap_f64_3d(&tmA, A_d, 2ull * K, (uint64_t)M, (uint64_t)za_count, 16ull * K, 16ull * K * M, 16,
                                cfg.BM, true);
    if (rc) return rc;
    if (cfg.m3) {
        // the planes live in one workspace of the handle: order this call behind the previous user's GEMM
        if (!h->bforms_ev) LDM_CUDA_OK(cudaEventCreateWithFlags(&h->bforms_ev, cudaEventDisableTiming));
        else LDM_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, h->bforms_ev, 0));
        rc = make_bforms(h, (cudaStream_t)stream, WS_BFORMS, B_d, zb_count, N, K, cfg.BN, &tmB);
    } else {
        rc = encode_tmap_f64_3d(&tmB, B_d, 2ull * K, (uint64_t)N, (uint64_t)zb_count, 16ull * K, 16ull * K * N, 8,
                                cfg.BN, false);
    }
    if (rc) return rc;
    std::vector<ZSeg> segs((size_t)nseg * nbatch);
    for (size_t i = 0; i < segs.size(); ++i) {
        segs[i].az = segs_h[4 * i + 0];
        segs[i].bz = segs_h[4 * i + 1];
        segs[i].conjA = segs_h[4 * i + 2] ? 0x80000000u : 0u;
        segs[i].conjB = segs_h[4 * i + 3] ? 0x80000000u : 0u;
        LDM_REQUIRE(segs[i].az >= 0 && segs[i].az < za_count && segs[i].bz >= 0 && segs[i].bz < zb_count,
                    "segment slice out of range");
    }
    std::vector<long long> offs;
    if (c_off_h) offs.assign(c_off_h, c_off_h + nbatch);
    rc = launch_zgemm(h, (cudaStream_t)stream, cfg, tmA, tmB, M, N, K, nseg, nbatch, segs.data(),
                      static_cast<double2*>(C_d), c_off_h ? offs.data() : nullptr, rdiv, s_outer, s_inner, s_col,
                      alpha, accumulate);
    if (rc) return rc;
    if (cfg.m3) LDM_CUDA_OK(cudaEventRecord(h->bforms_ev, (cudaStream_t)stream));
    return 0;
}

int ldm_dgemm_tn(ldm_handle h, void* stream, const double* A_d, int64_t lda, const double* B_d, int64_t ldb, int M,
                 int N, int K, double* C_d, int64_t ldc, double alpha, int accumulate, int lower_only) {
    LDM_REQUIRE(h && A_d && B_d && C_d, "null pointer");
    LDM_REQUIRE((lda % 2) == 0 && (ldb % 2) == 0, "leading dimensions must be even (16-byte rows for TMA)");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    return launch_dgemm(h, (cudaStream_t)stream, A_d, lda, B_d, ldb, M, N, K, C_d, ldc, alpha, accumulate,
                        lower_only);
}

int ldm_mirror_lower(ldm_handle h, void* stream, double* C_d, int n, int64_t ldc) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    const long long t = (n + MIRROR_T - 1) / MIRROR_T;
    mirror_lower_kernel<<<(unsigned)(t * (t + 1) / 2), 256, 0, (cudaStream_t)stream>>>(C_d, n, ldc);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_phase_transform(ldm_handle h, void* stream, const void* in_d, void* out_d, const void* W_d, int nin,
                        int nout, int64_t X, int batch, double scale, int in_real, int out_real,
                        double* imag_max_h) {
    LDM_REQUIRE(h && in_d && out_d && W_d, "null pointer");
    LDM_REQUIRE(nin > 0 && nout > 0 && X > 0 && batch > 0, "shape");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int KT = 8;
    LDM_REQUIRE((size_t)KT * nin * 16 <= 48 * 1024, "too many cells for the phase-matrix kernel");
    if (out_real) LDM_CUDA_OK(cudaMemsetAsync(h->imag_d, 0, sizeof(unsigned long long), st));
    dim3 grid((unsigned)((X + 127) / 128), (unsigned)((nout + KT - 1) / KT), (unsigned)batch);
    phase_transform_kernel<KT><<<grid, 128, (size_t)KT * nin * 16, st>>>(
        static_cast<const double*>(in_d), static_cast<double*>(out_d), static_cast<const double2*>(W_d), nin, nout,
        (long long)X, scale, in_real, out_real, out_real ? h->imag_d : nullptr);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    if (out_real && imag_max_h) {
        unsigned long long bits = 0;
        LDM_CUDA_OK(cudaMemcpyAsync(&bits, h->imag_d, sizeof(bits), cudaMemcpyDeviceToHost, st));
        LDM_CUDA_OK(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, &bits, sizeof(v));
        *imag_max_h = v;
    }
    return 0;
}

int ldm_lattice_dft(ldm_handle h, void* stream, const void* in_d, void* out_d, const int32_t* kmesh3, int64_t X,
                    int batch, int forward, double scale, int in_real, int out_real, double* imag_max_h) {
    LDM_REQUIRE(h && in_d && out_d && kmesh3, "null pointer");
    const int n0 = kmesh3[0], n1 = kmesh3[1], n2 = kmesh3[2];
    LDM_REQUIRE(n0 >= 1 && n1 >= 1 && n2 >= 1 && n0 <= 8 && n1 <= 8 && n2 <= 8, "mesh axes must have 1..8 points");
    LDM_REQUIRE(X > 0 && batch > 0, "shape");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nk = n0 * n1 * n2;
    int TX = 32;
    while (TX > 1 && (size_t)nk * TX * 16 > 96 * 1024) TX >>= 1;
    const size_t smem = (size_t)nk * TX * 16;
    DftTables tb;
    {
        const int nd[3] = {n0, n1, n2};
        const double sgn = forward ? -1.0 : 1.0;
        for (int d = 0; d < 3; ++d)
            for (int k = 0; k < nd[d]; ++k)
                for (int r = 0; r < nd[d]; ++r) {
                    const int m = (k * r) % nd[d];
                    double c = std::cos(2.0 * M_PI * m / nd[d]), sn = std::sin(2.0 * M_PI * m / nd[d]);
                    if (4 * m == nd[d]) { c = 0.0; sn = 1.0; }            // exact quarter turns
                    else if (2 * m == nd[d]) { c = -1.0; sn = 0.0; }
                    else if (4 * m == 3 * nd[d]) { c = 0.0; sn = -1.0; }
                    else if (m == 0) { c = 1.0; sn = 0.0; }
                    tb.w[d][k * nd[d] + r] = make_double2(c, sgn * sn);
                }
    }
    LDM_REQUIRE(smem <= 200 * 1024, "mesh too large for the shared-memory DFT");
    static bool attr = false;
    if (!attr) {
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)lattice_dft_kernel<4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)lattice_dft_kernel<8>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (out_real) LDM_CUDA_OK(cudaMemsetAsync(h->imag_d, 0, sizeof(unsigned long long), st));
    dim3 grid((unsigned)((X + TX - 1) / TX), (unsigned)batch);
    if (n0 <= 4 && n1 <= 4 && n2 <= 4)
        lattice_dft_kernel<4><<<grid, 256, smem, st>>>(static_cast<const double*>(in_d), static_cast<double*>(out_d), tb,
                                                       n0, n1, n2, (long long)X, TX, scale, in_real, out_real,
                                                       out_real ? h->imag_d : nullptr);
    else
        lattice_dft_kernel<8><<<grid, 256, smem, st>>>(static_cast<const double*>(in_d), static_cast<double*>(out_d), tb,
                                                       n0, n1, n2, (long long)X, TX, scale, in_real, out_real,
                                                       out_real ? h->imag_d : nullptr);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    if (out_real && imag_max_h) {
        unsigned long long bits = 0;
        LDM_CUDA_OK(cudaMemcpyAsync(&bits, h->imag_d, sizeof(bits), cudaMemcpyDeviceToHost, st));
        LDM_CUDA_OK(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, &bits, sizeof(v));
        *imag_max_h = v;
    }
    return 0;
}

int ldm_ztranspose(ldm_handle h, void* stream, const void* in_d, void* out_d, int batch, int rows, int cols,
                   int conj, double scale) {
    LDM_REQUIRE(h && in_d && out_d && batch > 0 && rows > 0 && cols > 0, "arguments");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    ztranspose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(static_cast<const double2*>(in_d),
                                                                       static_cast<double2*>(out_d), rows, cols, conj,
                                                                       scale);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

static int check_stored_shape(int naux, int rows, int nao, int64_t ncols, int flags, int* packed) {
    LDM_REQUIRE(naux > 0 && nao > 0 && rows >= 0 && rows <= naux, "stored entry: rows must lie in [0, naux]");
    LDM_REQUIRE((flags & ~(LDM_STORED_SWAPPED | LDM_STORED_REAL | LDM_STORED_CONJ)) == 0, "stored entry: unknown flags");
    const int64_t full = (int64_t)nao * nao, tri = (int64_t)nao * (nao + 1) / 2;
    LDM_REQUIRE(ncols == full || ncols == tri, "stored entry: ncols must be nao*nao or nao*(nao+1)/2");
    *packed = (ncols != full) ? 1 : 0;
    return 0;
}

int ldm_unpack_stored(ldm_handle h, void* stream, const void* src_d, void* out_d, int naux, int rows, int nao,
                      int64_t ncols, int flags) {
    LDM_REQUIRE(h && out_d && (src_d || rows == 0), "arguments");
    int packed = 0;
    int rc = check_stored_shape(naux, rows, nao, ncols, flags, &packed);
    if (rc) return rc;
    LDM_CUDA_OK(cudaSetDevice(h->device));
    const int nt = (nao + 31) / 32;
    dim3 grid(nt, nt, std::min(naux, 32768));
    const int swapped = (flags & LDM_STORED_SWAPPED) ? 1 : 0;
    const int conj = (flags & LDM_STORED_CONJ) ? 1 : 0;
    if (flags & LDM_STORED_REAL)
        unpack_stored_kernel<true><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
            src_d, static_cast<double2*>(out_d), naux, rows, nao, (long long)ncols, packed, swapped, conj);
    else
        unpack_stored_kernel<false><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
            src_d, static_cast<double2*>(out_d), naux, rows, nao, (long long)ncols, packed, swapped, conj);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_pack_tril(ldm_handle h, void* stream, const void* in_d, void* out_d, int rows, int n, int out_real,
                  double* imag_max_h) {
    LDM_REQUIRE(h && in_d && out_d && rows > 0 && n > 0, "arguments");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long npair = (long long)n * (n + 1) / 2;
    if (out_real) {
        LDM_CUDA_OK(cudaMemsetAsync(h->imag_d, 0, sizeof(unsigned long long), st));
        pack_tril_kernel<true><<<rows, 256, 0, st>>>(static_cast<const double2*>(in_d), out_d, n, npair, h->imag_d);
    } else {
        pack_tril_kernel<false><<<rows, 256, 0, st>>>(static_cast<const double2*>(in_d), out_d, n, npair, nullptr);
    }
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    if (out_real && imag_max_h) {
        unsigned long long bits = 0;
        LDM_CUDA_OK(cudaMemcpyAsync(&bits, h->imag_d, sizeof(bits), cudaMemcpyDeviceToHost, st));
        LDM_CUDA_OK(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, &bits, sizeof(v));
        *imag_max_h = v;
    }
    return 0;
}

int ldm_d2z(ldm_handle h, void* stream, const double* in_d, void* out_d, int64_t n) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    d2z_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in_d, static_cast<double2*>(out_d), (size_t)n);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_ksum_real(ldm_handle h, void* stream, const void* in_d, double* out_d, int nk, int64_t X, double scale,
                  double* imag_max_h) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    LDM_CUDA_OK(cudaMemsetAsync(h->imag_d, 0, sizeof(unsigned long long), st));
    ksum_real_kernel<<<(unsigned)((X + 255) / 256), 256, 0, st>>>(static_cast<const double2*>(in_d), out_d, nk,
                                                                  (long long)X, scale, h->imag_d);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    if (imag_max_h) {
        unsigned long long bits = 0;
        LDM_CUDA_OK(cudaMemcpyAsync(&bits, h->imag_d, sizeof(bits), cudaMemcpyDeviceToHost, st));
        LDM_CUDA_OK(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, &bits, sizeof(v));
        *imag_max_h = v;
    }
    return 0;
}

int ldm_restore_s1(ldm_handle h, void* stream, const double* eri4_d, double* out_d, int n) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    long long npair = (long long)n * (n + 1) / 2;
    if ((size_t)npair * 8 <= 200 * 1024)
        restore_s1_kernel<true><<<(unsigned)npair, 256, (size_t)npair * 8, (cudaStream_t)stream>>>(eri4_d, out_d, n,
                                                                                                 npair);
    else
        restore_s1_kernel<false><<<(unsigned)npair, 256, 0, (cudaStream_t)stream>>>(eri4_d, out_d, n, npair);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_restore_s8(ldm_handle h, void* stream, const double* eri4_d, double* out_d, int n) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    long long npair = (long long)n * (n + 1) / 2;
    restore_s8_kernel<<<(unsigned)npair, 256, 0, (cudaStream_t)stream>>>(eri4_d, out_d, npair);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_jk_s4(ldm_handle h, void* stream, const double* eri4_d, const double* dm_d, double* vj_d, double* vk_d,
              int n) {
    LDM_REQUIRE(h && eri4_d && dm_d && vj_d, "null pointer");
    LDM_REQUIRE(n > 0 && n <= 3200, "orbital count out of range");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    long long npair = (long long)n * (n + 1) / 2;
    const bool staged = (size_t)(npair + 2 * n + 8) * 8 <= 200 * 1024;
    size_t smem = (size_t)((staged ? npair : 0) + 2 * n + 8) * 8;
    size_t need = (size_t)npair * 16 + (size_t)npair * 2 * n * 8;
    if (h->jk_part_bytes < need) {
        if (h->jk_part_d) LDM_CUDA_OK(cudaFree(h->jk_part_d));
        h->jk_part_d = nullptr;
        h->jk_part_bytes = 0;
        LDM_CUDA_OK(cudaMalloc(&h->jk_part_d, need));
        h->jk_part_bytes = need;
    }
    double* vj_packed = static_cast<double*>(h->jk_part_d);
    double* dd = vj_packed + npair;
    double* kpart = dd + npair;
    jk_pack_dm_kernel<<<(unsigned)std::min<long long>((npair + 255) / 256, 1024), 256, 0, st>>>(dm_d, dd, n);
    LDM_CUDA_OK(cudaGetLastError());
    const int wk = vk_d != nullptr;
    const size_t rowb = (size_t)(npair + 2) * 8;      // bulk-copy variant: the packed row in shared memory
#define LDM_JK_LAUNCH(NCH)                                                                                         \
    do {                                                                                                           \
        if (n > 16 && rowb + 16 * 1024 <= 227 * 1024)                                                              \
            jk_rows_bulk_kernel<32 * NCH><<<(unsigned)npair, 64 * NCH, rowb, st>>>(eri4_d, dm_d, dd, vj_packed,   \
                                                                                   kpart, n, npair, wk);           \
        else                                                                                                       \
            jk_rows_seg_kernel<NCH><<<(unsigned)npair, 256, 0, st>>>(eri4_d, dm_d, dd, vj_packed, kpart, n, npair, \
                                                                     wk);                                          \
    } while (0)
    if (n <= 64)
        LDM_JK_LAUNCH(2);
    else if (n <= 128)
        LDM_JK_LAUNCH(4);
    else if (n <= 160)
        LDM_JK_LAUNCH(5);
    else if (n <= 256)
        LDM_JK_LAUNCH(8);
#undef LDM_JK_LAUNCH
    else if (staged)
        jk_rows_kernel<true><<<(unsigned)npair, 256, smem, st>>>(eri4_d, dm_d, dd, vj_packed, kpart, n, npair, wk);
    else
        jk_rows_kernel<false><<<(unsigned)npair, 256, smem, st>>>(eri4_d, dm_d, dd, vj_packed, kpart, n, npair, wk);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    unpack_sym_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(vj_packed, vj_d, n);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches += 2;
    if (vk_d) {
        jk_reduce_kernel<<<n, 256, (size_t)8 * n * 8, st>>>(kpart, vk_d, n);
        LDM_CUDA_OK(cudaGetLastError());
        h->launches++;
    }
    return 0;
}

int ldm_jk_s4_symm(ldm_handle h, void* stream, const double* eri4_d, const double* dm_d, double* vj_d, double* vk_d,
                   int n) {
    LDM_REQUIRE(h && eri4_d && dm_d && vj_d, "null pointer");
    LDM_REQUIRE(n > 0 && n <= 3200, "orbital count out of range");
    const long long npair = (long long)n * (n + 1) / 2;
    // ring of row slots in shared memory: at least two slots of the longest row (+ its zero-filled tail) must fit
    // beside ~10 KB of static shared memory; otherwise the general kernels serve the call
    const int NPt = n <= 64 ? 64 : (n <= 128 ? 128 : 160);
    const int slot_max = (int)((npair - 1 + 2 * NPt + 6) & ~1LL);
    const int rowbuf_words = 214 * 1024 / 8;          // the whole arena: short rows get a deeper ring
    const size_t smem = (size_t)rowbuf_words * 8;
    if (n <= 16 || n > 160 || 2LL * slot_max * 8 > 214 * 1024)
        return ldm_jk_s4(h, stream, eri4_d, dm_d, vj_d, vk_d, n);
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long nblk = (npair + JKT_ROWS - 1) / JKT_ROWS;
    // workspace: vj_packed | dd | vj_row | S (n*n) | kpart (npair * 2n) | jpart (nblk * npair)
    const size_t need = ((size_t)npair * 3 + (size_t)n * n + (size_t)npair * 2 * n + (size_t)nblk * npair) * 8;
    if (h->jk_part_bytes < need) {
        if (h->jk_part_d) {
            LDM_CUDA_OK(cudaDeviceSynchronize());
            LDM_CUDA_OK(cudaFree(h->jk_part_d));
            h->jk_part_d = nullptr;
            h->jk_part_bytes = 0;
        }
        LDM_CUDA_OK(cudaMalloc(&h->jk_part_d, need));
        h->jk_part_bytes = need;
    }
    double* vj_packed = static_cast<double*>(h->jk_part_d);
    double* dd = vj_packed + npair;
    double* vj_row = dd + npair;
    double* S = vj_row + npair;
    double* kpart = S + (size_t)n * n;
    double* jpart = kpart + (size_t)npair * 2 * n;
    jk_pack_dm_kernel<<<(unsigned)std::min<long long>((npair + 255) / 256, 1024), 256, 0, st>>>(dm_d, dd, n);
    LDM_CUDA_OK(cudaGetLastError());
    const int wk = vk_d != nullptr;
    static const bool use_tri2 = getenv("LDM_JK_TRI2") ? atoi(getenv("LDM_JK_TRI2")) != 0 : true;
    static bool attr = false;
    if (!attr) {
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_tri_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         214 * 1024));
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_tri_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         214 * 1024));
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_tri_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         214 * 1024));
        LDM_CUDA_OK(cudaFuncSetAttribute((const void*)jk_tri2_kernel<152, 384>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024));
        attr = true;
    }
    if (n <= 64)
        jk_tri_kernel<64><<<(unsigned)nblk, 128, smem, st>>>(eri4_d, dm_d, dd, vj_row, jpart, kpart, n, npair, wk,
                                                            rowbuf_words);
    else if (n <= 128)
        jk_tri_kernel<128><<<(unsigned)nblk, 256, smem, st>>>(eri4_d, dm_d, dd, vj_row, jpart, kpart, n, npair, wk,
                                                             rowbuf_words);
    else if (n <= 152 && use_tri2)      // dd in registers (LDM_JK_TRI2=0: the kernel that fetches it from L2 per row)
        jk_tri2_kernel<152, 384><<<(unsigned)nblk, 384, smem, st>>>(eri4_d, dm_d, dd, vj_row, jpart, kpart, n, npair,
                                                                     wk, rowbuf_words);
    else
        jk_tri_kernel<160><<<(unsigned)nblk, 320, smem, st>>>(eri4_d, dm_d, dd, vj_row, jpart, kpart, n, npair, wk,
                                                             rowbuf_words);
    LDM_CUDA_OK(cudaGetLastError());
    jk_tri_jsum_kernel<<<(unsigned)((npair + 31) / 32), dim3(32, 8), 0, st>>>(vj_row, jpart, vj_packed, npair);
    LDM_CUDA_OK(cudaGetLastError());
    unpack_sym_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(vj_packed, vj_d, n);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches += 4;
    if (vk_d) {
        jk_reduce_kernel<<<n, 256, (size_t)8 * n * 8, st>>>(kpart, S, n);
        LDM_CUDA_OK(cudaGetLastError());
        symmetrise_add_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(S, vk_d, n);
        LDM_CUDA_OK(cudaGetLastError());
        h->launches += 2;
    }
    return 0;
}

int ldm_scale_eri(ldm_handle h, void* stream, double* eri_d, int n, int symmetry, const int32_t* weights_d) {
    LDM_REQUIRE(h && eri_d && weights_d && n > 0 && (symmetry == 1 || symmetry == 4), "arguments");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    const long long npair = (long long)n * (n + 1) / 2;
    if (symmetry == 4)
        scale_s4_kernel<<<(unsigned)npair, 256, 0, (cudaStream_t)stream>>>(eri_d, weights_d, npair);
    else
        scale_s1_kernel<<<n * n, 256, 0, (cudaStream_t)stream>>>(eri_d, weights_d, n);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int ldm_synth_block(ldm_handle h, void* stream, void* out_d, int naux, int nao, int aux_offset, uint32_t key_ij,
                    uint32_t key_ji, uint32_t key_mij, uint32_t key_mji, double scale) {
    LDM_CUDA_OK(cudaSetDevice(h->device));
    LDM_REQUIRE(aux_offset >= 0 && 2.0 * ((double)naux + aux_offset) * nao * (double)nao < 4294967296.0,
                "block too large for the 32-bit counter");
    synth_block_kernel<<<h->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(static_cast<double2*>(out_d), naux, nao,
                                                                          aux_offset, key_ij, key_ji, key_mij,
                                                                          key_mji, scale);
    LDM_CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------------------
// embedding-ERI pipeline
// ----------------------------------------------------------------------------------------------------------
struct PendingBlock {
    int ki, kj, sym;
    int slot;      // slice index in the tensor map of its source
    int source;    // 0 = ring, 1 = store
};

struct EriPlan {
    cudaStream_t st = nullptr, copy_st = nullptr;
    int nk = 0, nao = 0, naux = 0, neo = 0, nspin = 0, G = 1, klg = 1;
    int gso = 0;                 // 1: two spin flavours, one ERI from Lambda_a - Lambda_b
    long long npair = 0, ldx = 0;
    const double2* CT = nullptr;
    double* eri = nullptr;
    const ZConfig* cfg = nullptr;
    // workspaces
    double2* ring = nullptr;     // [2G][naux][nao][nao]
    int ring_slots = 0, ring_next = 0;
    std::vector<cudaEvent_t> ring_free;   // recorded after the launch that last read a slot
    std::vector<char> ring_busy;
    const double2* store = nullptr;
    int store_slots = 0;
    double2* Xt = nullptr;       // [nspin][G][naux][neo][nao]
    double2* S_sym = nullptr;    // [nspin][naux][neo][neo]
    double2* S_pln = nullptr;
    bool sym_init = false, pln_init = false;
    double* XT = nullptr;        // [nspin][npair][ldx]
    long long xt_cols = 0;
    double xt_alpha = 0.0;
    int kl_in_panel = 0;
    int nauxp = 0;               // column stride of one (re or im) part in the panel: naux rounded up to even
    double* imag = nullptr;      // optional (spin_pair, npair, npair): Im(Lambda^dagger Lambda) of weight-0 units
    std::vector<long long> imag_cols;   // panel columns of the weight-0 units waiting in the panel
    CUtensorMap tmRing, tmStore, tmCT, tmXt;
    bool have_store_map = false, have_ring_map = false;
    std::vector<PendingBlock> pending;
    int64_t launches0 = 0, h2d_bytes = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[2];
};

static size_t block_elems(const EriPlan* p) { return (size_t)p->naux * p->nao * p->nao; }

static int plan_timed_begin(EriPlan* p, int kind) {
    cudaEvent_t a, b;
    LDM_CUDA_OK(cudaEventCreate(&a));
    LDM_CUDA_OK(cudaEventCreate(&b));
    LDM_CUDA_OK(cudaEventRecord(a, p->st));
    p->ev[kind].push_back({a, b});
    return 0;
}
static int plan_timed_end(EriPlan* p, int kind) {
    LDM_CUDA_OK(cudaEventRecord(p->ev[kind].back().second, p->st));
    return 0;
}

static int ensure_ring(ldm_handle h) {
    EriPlan* p = h->plan;
    if (p->ring) return 0;
    p->ring_slots = 2 * p->G;
    {
        void* q = nullptr;
        int rc0 = ws_get(h, WS_RING, (size_t)p->ring_slots * block_elems(p) * 16, &q);
        if (rc0) return rc0;
        p->ring = static_cast<double2*>(q);
    }
    p->ring_free.resize(p->ring_slots);
    p->ring_busy.assign(p->ring_slots, 0);
    for (auto& e : p->ring_free) LDM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = encode_tmap_f64_3d(&p->tmRing, p->ring, 2ull * p->nao, (uint64_t)p->naux * p->nao,
                                (uint64_t)p->ring_slots, 16ull * p->nao, 16ull * block_elems(p), 16, p->cfg->BM, true);
    if (rc) return rc;
    p->have_ring_map = true;
    return 0;
}

// stage 1 for the pending group: (a) half transform with C_j, written transposed; (b) second half transform with
// conj(C_i), chained over the blocks of the group, accumulated into S_sym / S_pln.
static int flush_group(ldm_handle h) {
    EriPlan* p = h->plan;
    if (p->pending.empty()) return 0;
    const int nb = (int)p->pending.size();
    const int src = p->pending[0].source;
    const size_t xt_slice = (size_t)p->naux * p->neo * p->nao;
    // ---- (a) Xt[s][g][L][n][p] = sum_q L_g[L][p][q] * CT[s][kj][n][q]
    {
        std::vector<ZSeg> segs;
        std::vector<long long> offs;
        for (int s = 0; s < p->nspin; ++s)
            for (int g = 0; g < nb; ++g) {
                segs.push_back(ZSeg{p->pending[g].slot, s * p->nk + p->pending[g].kj, 0u, 0u});
                offs.push_back((long long)((size_t)(s * p->G + g) * xt_slice));
            }
        int rc = plan_timed_begin(p, 0);
        if (rc) return rc;
        rc = launch_zgemm(h, p->st, *p->cfg, src == 0 ? p->tmRing : p->tmStore, p->tmCT, p->naux * p->nao, p->neo,
                          p->nao, 1, p->nspin * nb, segs.data(), p->Xt, offs.data(), p->nao,
                          (long long)p->neo * p->nao, 1, p->nao, 1.0, 0);
        if (rc) return rc;
        if (src == 0)
            for (int g = 0; g < nb; ++g) {
                LDM_CUDA_OK(cudaEventRecord(p->ring_free[p->pending[g].slot], p->st));
                p->ring_busy[p->pending[g].slot] = 1;
            }
    }
    // ---- (b) S[s][L][n][m] (+)= sum_g sum_p Xt[s][g][L][n][p] * conj(CT[s][ki_g][m][p])
    for (int pass = 0; pass < 2; ++pass) {
        const int want_sym = pass == 0 ? 1 : 0;
        std::vector<int> members;
        for (int g = 0; g < nb; ++g)
            if ((p->pending[g].sym != 0) == (want_sym != 0)) members.push_back(g);
        if (members.empty()) continue;
        std::vector<ZSeg> segs;
        std::vector<long long> offs;
        for (int s = 0; s < p->nspin; ++s) {
            for (int g : members) segs.push_back(ZSeg{s * p->G + g, s * p->nk + p->pending[g].ki, 0u, 0x80000000u});
            offs.push_back((long long)((size_t)s * p->naux * p->neo * p->neo));
        }
        double2* S = want_sym ? p->S_sym : p->S_pln;
        bool& init = want_sym ? p->sym_init : p->pln_init;
        int rc = launch_zgemm(h, p->st, *p->cfg, p->tmXt, p->tmCT, p->naux * p->neo, p->neo, p->nao,
                              (int)members.size(), p->nspin, segs.data(), S, offs.data(), 1, p->neo, 0, 1, 1.0,
                              init ? 1 : 0);
        if (rc) return rc;
        init = true;
    }
    int rc = plan_timed_end(p, 0);
    if (rc) return rc;
    p->pending.clear();
    return 0;
}

static int flush_panel(ldm_handle h) {
    EriPlan* p = h->plan;
    if (p->xt_cols == 0) return 0;
    const int K = (int)p->xt_cols;
    const size_t xt_spin = (size_t)p->npair * p->ldx;
    const size_t eri_blk = (size_t)p->npair * p->npair;
    int rc = plan_timed_begin(p, 1);
    if (rc) return rc;
    if (p->nspin == 1 || p->gso) {
        rc = launch_dgemm(h, p->st, p->XT, p->ldx, p->XT, p->ldx, (int)p->npair, (int)p->npair, K, p->eri, p->npair,
                          p->xt_alpha, 1, 1);
        if (rc) return rc;
    } else {
        // incore order of the reference: aa, ab, bb  (eri_transform.py:463-478)
        rc = launch_dgemm(h, p->st, p->XT, p->ldx, p->XT, p->ldx, (int)p->npair, (int)p->npair, K, p->eri, p->npair,
                          p->xt_alpha, 1, 1);
        if (rc) return rc;
        rc = launch_dgemm(h, p->st, p->XT, p->ldx, p->XT + xt_spin, p->ldx, (int)p->npair, (int)p->npair, K,
                          p->eri + eri_blk, p->npair, p->xt_alpha, 1, 0);
        if (rc) return rc;
        rc = launch_dgemm(h, p->st, p->XT + xt_spin, p->ldx, p->XT + xt_spin, p->ldx, (int)p->npair, (int)p->npair,
                          K, p->eri + 2 * eri_blk, p->npair, p->xt_alpha, 1, 1);
        if (rc) return rc;
    }
    // imaginary part of Lambda^dagger Lambda for the units built without time reversal (diagnostic of
    // eri_transform.py:390-395):  Im[x,y] += Re_x^T Im_y - Im_x^T Re_y  per transfer momentum
    for (long long col : p->imag_cols) {
        const int nsp = (p->nspin == 1 || p->gso) ? 1 : 2;
        int blk = 0;
        for (int x = 0; x < nsp; ++x)
            for (int y = x; y < nsp; ++y, ++blk) {
                const double* RX = p->XT + (size_t)x * xt_spin + col;
                const double* RY = p->XT + (size_t)y * xt_spin + col;
                rc = launch_dgemm(h, p->st, RX, p->ldx, RY + p->nauxp, p->ldx, (int)p->npair, (int)p->npair, p->naux,
                                  p->imag + (size_t)blk * eri_blk, p->npair, 1.0, 1, 0);
                if (rc) return rc;
                rc = launch_dgemm(h, p->st, RX + p->nauxp, p->ldx, RY, p->ldx, (int)p->npair, (int)p->npair, p->naux,
                                  p->imag + (size_t)blk * eri_blk, p->npair, -1.0, 1, 0);
                if (rc) return rc;
            }
    }
    p->imag_cols.clear();
    rc = plan_timed_end(p, 1);
    if (rc) return rc;
    p->xt_cols = 0;
    p->kl_in_panel = 0;
    return 0;
}

extern "C" {

static int eri_begin_body(ldm_handle h, EriPlan* p, int nkpts, int nao, int naux, int neo, int nspin,
                          const void* CT_d) {
    if (!h->copy_st) LDM_CUDA_OK(cudaStreamCreateWithFlags(&h->copy_st, cudaStreamNonBlocking));
    p->copy_st = h->copy_st;            // a small build must not pay for a stream creation
    const size_t xt_slice = (size_t)naux * neo * nao;
    const size_t s_elems = (size_t)nspin * naux * neo * neo;
    void* q = nullptr;
    int rc = ws_get(h, WS_XT, (size_t)nspin * p->G * xt_slice * 16, &q);
    if (rc) return rc;
    p->Xt = static_cast<double2*>(q);
    rc = ws_get(h, WS_SSYM, s_elems * 16, &q);
    if (rc) return rc;
    p->S_sym = static_cast<double2*>(q);
    rc = ws_get(h, WS_SPLN, s_elems * 16, &q);
    if (rc) return rc;
    p->S_pln = static_cast<double2*>(q);
    rc = ws_get(h, WS_PANEL, (size_t)nspin * p->npair * p->ldx * 8, &q);
    if (rc) return rc;
    p->XT = static_cast<double*>(q);
    if (p->cfg->m3)
        rc = make_bforms(h, p->st, WS_CTFORMS, CT_d, nspin * nkpts, neo, nao, p->cfg->BN, &p->tmCT);
    else
        rc = encode_tmap_f64_3d(&p->tmCT, CT_d, 2ull * nao, (uint64_t)neo, (uint64_t)nspin * nkpts, 16ull * nao,
                                16ull * nao * neo, 8, p->cfg->BN, false);
    if (rc) return rc;
    return encode_tmap_f64_3d(&p->tmXt, p->Xt, 2ull * nao, (uint64_t)naux * neo, (uint64_t)nspin * p->G, 16ull * nao,
                              16ull * xt_slice, 16, p->cfg->BM, true);
}

int ldm_eri_begin(ldm_handle h, void* stream, int nkpts, int nao, int naux, int neo, int nspin, const void* CT_d,
                  double* eri_d, int max_group, int kl_group) {
    LDM_REQUIRE(h && CT_d && eri_d, "null pointer");
    LDM_REQUIRE(nkpts > 0 && nao > 0 && naux > 0 && neo > 0 && (nspin == 1 || nspin == 2), "shape");
    LDM_REQUIRE(h->plan == nullptr, "an ERI build is already open on this handle");
    LDM_REQUIRE((double)naux * nao < 2147483647.0 && (double)naux * neo < 2147483647.0, "naux*nao too large");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    EriPlan* p = new EriPlan();
    p->st = (cudaStream_t)stream;
    p->nk = nkpts; p->nao = nao; p->naux = naux; p->neo = neo; p->nspin = nspin;
    p->G = std::max(1, max_group);
    p->klg = std::max(1, kl_group);
    p->npair = (long long)neo * (neo + 1) / 2;
    p->nauxp = naux + (naux & 1);
    p->ldx = ((long long)p->klg * 2 * p->nauxp + 15) / 16 * 16;
    p->CT = static_cast<const double2*>(CT_d);
    p->eri = eri_d;
    p->cfg = &pick_zconfig(neo, h->zgemm_3m);
    p->launches0 = h->launches;
    h->plan = p;
    const int rc = eri_begin_body(h, p, nkpts, nao, naux, neo, nspin, CT_d);
    if (rc) {
        // a failed allocation (typically cudaMalloc out of memory for a workspace) must not leave the handle with a
        // half-built plan attached: every later build would be refused.  The error text is kept.
        const std::string msg = ldm::last_error();
        cudaGetLastError();
        ldm_eri_end(h);
        set_error(msg);
    }
    return rc;
}

int ldm_eri_set_mode(ldm_handle h, int gso) {
    LDM_REQUIRE(h && h->plan, "arguments");
    LDM_REQUIRE(!gso || h->plan->nspin == 2, "GSO mode needs two spin flavours");
    LDM_REQUIRE(h->plan->pending.empty() && h->plan->xt_cols == 0, "mode must be set before the first block");
    h->plan->gso = gso ? 1 : 0;
    return 0;
}

int ldm_eri_set_imag(ldm_handle h, double* imag_d) {
    LDM_REQUIRE(h && h->plan && imag_d, "arguments");
    LDM_REQUIRE(h->plan->xt_cols == 0, "must be set before the first ldm_eri_end_kl");
    h->plan->imag = imag_d;
    return 0;
}

int ldm_max_abs(ldm_handle h, void* stream, const double* x_d, int64_t n, double* out_h) {
    LDM_REQUIRE(h && x_d && out_h && n >= 0, "arguments");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    LDM_CUDA_OK(cudaMemsetAsync(h->imag_d, 0, sizeof(unsigned long long), st));
    if (n > 0) {
        max_abs_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(x_d, (long long)n,
                                                                                               h->imag_d);
        LDM_CUDA_OK(cudaGetLastError());
        h->launches++;
    }
    unsigned long long bits = 0;
    LDM_CUDA_OK(cudaMemcpyAsync(&bits, h->imag_d, sizeof(bits), cudaMemcpyDeviceToHost, st));
    LDM_CUDA_OK(cudaStreamSynchronize(st));
    double v;
    std::memcpy(&v, &bits, sizeof(v));
    *out_h = v;
    return 0;
}

int ldm_eri_set_store(ldm_handle h, const void* store_d, int nslots) {
    LDM_REQUIRE(h && h->plan && store_d && nslots > 0, "arguments");
    EriPlan* p = h->plan;
    p->store = static_cast<const double2*>(store_d);
    p->store_slots = nslots;
    int rc = encode_tmap_f64_3d(&p->tmStore, store_d, 2ull * p->nao, (uint64_t)p->naux * p->nao, (uint64_t)nslots,
                                16ull * p->nao, 16ull * block_elems(p), 16, p->cfg->BM, true);
    if (rc) return rc;
    p->have_store_map = true;
    return 0;
}

static int push_block(ldm_handle h, int ki, int kj, int sym, int slot, int source) {
    EriPlan* p = h->plan;
    if (!p->pending.empty() && p->pending[0].source != source) {
        int rc = flush_group(h);
        if (rc) return rc;
    }
    p->pending.push_back(PendingBlock{ki, kj, sym, slot, source});
    if ((int)p->pending.size() >= p->G) return flush_group(h);
    return 0;
}

static int ring_acquire(ldm_handle h, int* slot) {
    EriPlan* p = h->plan;
    int rc = ensure_ring(h);
    if (rc) return rc;
    *slot = p->ring_next;
    p->ring_next = (p->ring_next + 1) % p->ring_slots;
    // the slot may still be read by the stage-1 launch of two groups ago
    if (p->ring_busy[*slot]) {
        LDM_CUDA_OK(cudaEventSynchronize(p->ring_free[*slot]));
        p->ring_busy[*slot] = 0;
    }
    return 0;
}

int ldm_eri_block_host(ldm_handle h, int ki, int kj, int sym, const void* L_h) {
    LDM_REQUIRE(h && h->plan && L_h, "arguments");
    EriPlan* p = h->plan;
    LDM_REQUIRE(ki >= 0 && ki < p->nk && kj >= 0 && kj < p->nk, "k index");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    // a slot that is part of the not-yet-launched group must not be recycled: the ring has 2G slots and a group
    // holds at most G, so ring_next never laps the pending group.
    int slot;
    int rc = ring_acquire(h, &slot);
    if (rc) return rc;
    const size_t bytes = block_elems(p) * 16;
    LDM_CUDA_OK(cudaMemcpyAsync(p->ring + (size_t)slot * block_elems(p), L_h, bytes, cudaMemcpyHostToDevice,
                                p->copy_st));
    LDM_CUDA_OK(cudaStreamSynchronize(p->copy_st));   // source buffer is consumed when we return
    p->h2d_bytes += (int64_t)bytes;
    return push_block(h, ki, kj, sym, slot, 0);
}

int ldm_eri_block_stored(ldm_handle h, int ki, int kj, int sym, const void* src_h, int rows, int64_t ncols,
                         int flags) {
    LDM_REQUIRE(h && h->plan && (src_h || rows == 0), "arguments");
    EriPlan* p = h->plan;
    LDM_REQUIRE(ki >= 0 && ki < p->nk && kj >= 0 && kj < p->nk, "k index");
    int packed = 0;
    int rc = check_stored_shape(p->naux, rows, p->nao, ncols, flags, &packed);
    if (rc) return rc;
    LDM_CUDA_OK(cudaSetDevice(h->device));
    int slot;
    rc = ring_acquire(h, &slot);
    if (rc) return rc;
    double2* dst = p->ring + (size_t)slot * block_elems(p);
    const size_t bytes = (size_t)rows * (size_t)ncols * ((flags & LDM_STORED_REAL) ? 8 : 16);
    if (flags == 0 && !packed) {
        // already in block layout: straight into the ring slot, absent auxiliary rows zeroed
        if (bytes) LDM_CUDA_OK(cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, p->copy_st));
        if (rows < p->naux)
            LDM_CUDA_OK(cudaMemsetAsync(reinterpret_cast<char*>(dst) + bytes, 0, block_elems(p) * 16 - bytes,
                                        p->copy_st));
    } else {
        void* raw = nullptr;
        rc = ws_get(h, WS_STORED, block_elems(p) * 16, &raw);
        if (rc) return rc;
        if (bytes) LDM_CUDA_OK(cudaMemcpyAsync(raw, src_h, bytes, cudaMemcpyHostToDevice, p->copy_st));
        rc = ldm_unpack_stored(h, p->copy_st, raw, dst, p->naux, rows, p->nao, ncols, flags);
        if (rc) return rc;
    }
    LDM_CUDA_OK(cudaStreamSynchronize(p->copy_st));   // source buffer (and the raw scratch) are consumed on return
    p->h2d_bytes += (int64_t)bytes;
    return push_block(h, ki, kj, sym, slot, 0);
}

int ldm_eri_block_store(ldm_handle h, int ki, int kj, int sym, int slot) {
    LDM_REQUIRE(h && h->plan, "arguments");
    EriPlan* p = h->plan;
    LDM_REQUIRE(p->have_store_map, "no resident store registered (ldm_eri_set_store)");
    LDM_REQUIRE(slot >= 0 && slot < p->store_slots, "store slot");
    LDM_REQUIRE(ki >= 0 && ki < p->nk && kj >= 0 && kj < p->nk, "k index");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    return push_block(h, ki, kj, sym, slot, 1);
}

int ldm_eri_block_synth(ldm_handle h, int ki, int kj, int sym, int aux_offset, uint32_t key_ij, uint32_t key_ji,
                        uint32_t key_mij, uint32_t key_mji, double scale) {
    LDM_REQUIRE(h && h->plan, "arguments");
    EriPlan* p = h->plan;
    LDM_REQUIRE(ki >= 0 && ki < p->nk && kj >= 0 && kj < p->nk, "k index");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    int slot;
    int rc = ring_acquire(h, &slot);
    if (rc) return rc;
    rc = ldm_synth_block(h, p->st, p->ring + (size_t)slot * block_elems(p), p->naux, p->nao, aux_offset, key_ij,
                         key_ji, key_mij, key_mji, scale);
    if (rc) return rc;
    return push_block(h, ki, kj, sym, slot, 0);
}

int ldm_eri_end_kl(ldm_handle h, int weight) {
    LDM_REQUIRE(h && h->plan, "arguments");
    LDM_REQUIRE(weight >= 0 && weight <= 2, "weight must be 0 (no time reversal), 1 or 2");
    EriPlan* p = h->plan;
    LDM_CUDA_OK(cudaSetDevice(h->device));
    int rc = flush_group(h);
    if (rc) return rc;
    LDM_REQUIRE(p->sym_init || p->pln_init, "transfer momentum without blocks");
    const double alpha = weight == 2 ? 2.0 : 1.0;
    const int ncols = (weight == 1 ? 1 : 2) * p->nauxp;
    if (p->xt_cols > 0 && (p->xt_alpha != alpha || p->xt_cols + ncols > p->ldx || p->kl_in_panel >= p->klg)) {
        rc = flush_panel(h);
        if (rc) return rc;
    }
    p->xt_alpha = alpha;
    const long long col_re = p->xt_cols;
    const long long col_im = weight == 1 ? -1 : p->xt_cols + p->nauxp;
    if (p->nauxp != p->naux) {      // odd naux: the pad column of each part must not carry stale data
        for (int s = 0; s < (p->gso ? 1 : p->nspin); ++s)
            for (int part = 0; part < (weight == 1 ? 1 : 2); ++part) {
                const long long c0 = p->xt_cols + (long long)part * p->nauxp + p->naux;
                fill_cols_kernel<<<64, 256, 0, p->st>>>(p->XT + (size_t)s * p->npair * p->ldx, p->npair, p->ldx, c0,
                                                       c0 + 1);
                LDM_CUDA_OK(cudaGetLastError());
                h->launches++;
            }
    }
    if (weight == 0 && p->imag) p->imag_cols.push_back(p->xt_cols);
    const int t = (p->neo + 15) / 16;
    const size_t s_spin = (size_t)p->naux * p->neo * p->neo;
    if (p->gso) {
        pack_sym_kernel<true><<<dim3(t * (t + 1) / 2, (p->naux + 15) / 16), 256, 16 * 16 * 17 * 16, p->st>>>(
            p->sym_init ? p->S_sym : nullptr, p->pln_init ? p->S_pln : nullptr,
            p->sym_init ? p->S_sym + s_spin : nullptr, p->pln_init ? p->S_pln + s_spin : nullptr, p->XT, p->naux,
            p->neo, p->ldx, col_re, col_im);
        LDM_CUDA_OK(cudaGetLastError());
        h->launches++;
    } else {
        for (int s = 0; s < p->nspin; ++s) {
            pack_sym_kernel<false><<<dim3(t * (t + 1) / 2, (p->naux + 15) / 16), 256, 16 * 16 * 17 * 16, p->st>>>(
                p->sym_init ? p->S_sym + s * s_spin : nullptr, p->pln_init ? p->S_pln + s * s_spin : nullptr, nullptr,
                nullptr, p->XT + (size_t)s * p->npair * p->ldx, p->naux, p->neo, p->ldx, col_re, col_im);
            LDM_CUDA_OK(cudaGetLastError());
            h->launches++;
        }
    }
    p->xt_cols += ncols;
    p->kl_in_panel++;
    p->sym_init = p->pln_init = false;
    return 0;
}

int ldm_eri_finish(ldm_handle h) {
    LDM_REQUIRE(h && h->plan, "arguments");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    int rc = flush_group(h);
    if (rc) return rc;
    LDM_REQUIRE(!h->plan->sym_init && !h->plan->pln_init, "ldm_eri_end_kl missing before ldm_eri_finish");
    return flush_panel(h);
}

int ldm_eri_stats(ldm_handle h, int64_t* launches, int64_t* h2d_bytes) {
    LDM_REQUIRE(h && h->plan, "arguments");
    if (launches) *launches = h->launches - h->plan->launches0;
    if (h2d_bytes) *h2d_bytes = h->plan->h2d_bytes;
    return 0;
}

int ldm_eri_kernel_time(ldm_handle h, int kind, double* ms, int64_t* launches) {
    LDM_REQUIRE(h && h->plan && (kind == 0 || kind == 1), "arguments");
    EriPlan* p = h->plan;
    LDM_CUDA_OK(cudaStreamSynchronize(p->st));
    double tot = 0.0;
    for (auto& e : p->ev[kind]) {
        float t = 0.f;
        LDM_CUDA_OK(cudaEventElapsedTime(&t, e.first, e.second));
        tot += t;
    }
    if (ms) *ms = tot;
    if (launches) *launches = (int64_t)p->ev[kind].size();
    return 0;
}

int ldm_eri_end(ldm_handle h) {
    if (!h || !h->plan) return 0;
    EriPlan* p = h->plan;
    cudaSetDevice(h->device);
    // workspaces stay in the handle's pool and later builds queue behind this one on the same stream.  The staging
    // ring is the exception: the next build may fill it through its own copy stream (host blocks) while kernels of
    // this build still read it, so a build that touched the ring -- host OR device-generated blocks -- drains the
    // compute stream before it goes away.  Builds over a resident store never wait here.
    if (p->ring) cudaStreamSynchronize(p->st);
    if (p->copy_st) cudaStreamSynchronize(p->copy_st);
    for (int k = 0; k < 2; ++k)
        for (auto& e : p->ev[k]) {
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
    for (auto& e : p->ring_free)
        if (e) cudaEventDestroy(e);
    delete p;
    h->plan = nullptr;
    return 0;
}

int ldm_release_workspaces(ldm_handle h) {
    LDM_REQUIRE(h && !h->plan, "no build may be open");
    LDM_CUDA_OK(cudaSetDevice(h->device));
    LDM_CUDA_OK(cudaDeviceSynchronize());
    ws_release(h);
    return 0;
}

}  // extern "C"
