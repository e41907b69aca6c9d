#include "rmu_common.h"

#include <mutex>

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
    if (box_cols * elem_bytes != 128) { set_error("make_tmap_2d: box must be 128 bytes wide"); return RMU_ERR_ARG; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return RMU_ERR_CUDA;
    }
    return RMU_OK;
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
}
