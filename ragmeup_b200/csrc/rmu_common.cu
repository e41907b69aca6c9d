#include "rmu_common.h"

#include <mutex>
#include <vector>

namespace rmu {

static thread_local std::string t_err;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_err = msg; }

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    if (!fn) set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                 uint32_t box_cols, uint32_t box_rows, int elem_bytes) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return RMU_ERR_CUDA;
    CUtensorMapDataType dt;
    if (elem_bytes == 4) dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    else if (elem_bytes == 2) dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    else { set_error("make_tmap_2d: bad element size"); return RMU_ERR_ARG; }
    const uint32_t box_bytes = box_cols * elem_bytes;
    if (box_bytes != 128 && box_bytes != 64) { set_error("make_tmap_2d: box must be 128 or 64 bytes wide"); return RMU_ERR_ARG; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return RMU_ERR_CUDA;
    }
    return RMU_OK;
}

int make_tmap_rows_kblocks(CUtensorMap* out, const void* base, uint64_t rows, uint32_t kblocks,
                           uint32_t box_rows, uint32_t box_kb) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return RMU_ERR_CUDA;
    cuuint64_t gdim[3] = {32, rows, kblocks};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(kblocks) * 128, 128};   // bytes: row pitch, k-block pitch
    cuuint32_t box[3] = {32, box_rows, box_kb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed with CUresult " + std::to_string(static_cast<int>(r)));
        return RMU_ERR_CUDA;
    }
    return RMU_OK;
}

// ---------------------------------------------------------------- profiler
std::atomic<int> g_prof_on{0};
namespace {
struct ProfRec { int cls; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_pending;
std::vector<cudaEvent_t> g_prof_pool;
double g_prof_ms[PROF_NCLASS] = {0};
long long g_prof_n[PROF_NCLASS] = {0};
thread_local cudaEvent_t t_prof_open[PROF_NCLASS] = {nullptr};
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace
void prof_begin(int cls, cudaStream_t st) {
    std::lock_guard<std::mutex> g(g_prof_mu);
    cudaEvent_t a = prof_event();
    cudaEventRecord(a, st);
    t_prof_open[cls] = a;
}
void prof_end(int cls, cudaStream_t st) {
    std::lock_guard<std::mutex> g(g_prof_mu);
    cudaEvent_t b = prof_event();
    cudaEventRecord(b, st);
    g_prof_pending.push_back({cls, t_prof_open[cls], b});
}
static void prof_drain() {
    std::lock_guard<std::mutex> g(g_prof_mu);
    for (auto& r : g_prof_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_prof_ms[r.cls] += ms;
            g_prof_n[r.cls] += 1;
        }
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_pending.clear();
}

int device_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

}  // namespace rmu

extern "C" {
const char* rmu_last_error(void) { return rmu::t_err.c_str(); }
int rmu_version(void) { return 100; }
uint64_t rmu_launch_count(void) { return rmu::g_launches.load(); }
void rmu_profile_enable(int on) { rmu::g_prof_on.store(on ? 1 : 0); }
void rmu_profile_reset(void) {
    rmu::prof_drain();
    for (int i = 0; i < rmu::PROF_NCLASS; ++i) { rmu::g_prof_ms[i] = 0; rmu::g_prof_n[i] = 0; }
}
int rmu_profile_read(int cls, double* total_ms, int64_t* launches) {
    if (cls < 0 || cls >= rmu::PROF_NCLASS || !total_ms || !launches) return RMU_ERR_ARG;
    rmu::prof_drain();
    *total_ms = rmu::g_prof_ms[cls];
    *launches = rmu::g_prof_n[cls];
    return RMU_OK;
}
}
