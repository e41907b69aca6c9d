#include <cstdlib>

#include "rmu_gemm.cuh"

namespace rmu {

int make_split_operand(SplitOperand* op, __half* hi, __half* lo, int64_t rows, int cols, bool is_activation) {
    if (cols % kGemmBK != 0) { set_error("make_split_operand: K must be a multiple of 64"); return RMU_ERR_ARG; }
    op->hi = hi; op->lo = lo; op->rows = rows; op->cols = cols;
    const uint64_t pitch = static_cast<uint64_t>(cols) * sizeof(__half);
    int rc = make_tmap_2d(&op->map_hi, hi, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, kGemmBK, 128, 2);
    if (rc == RMU_OK) rc = make_tmap_2d(&op->map_lo, lo, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, kGemmBK, 128, 2);
    op->box64_rows = is_activation ? kWideBM : kWideBN;
    if (rc == RMU_OK) rc = make_tmap_2d(&op->map64_hi, hi, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, 32, op->box64_rows, 2);
    if (rc == RMU_OK) rc = make_tmap_2d(&op->map64_lo, lo, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, 32, op->box64_rows, 2);
    if (rc == RMU_OK) rc = make_tmap_2d(&op->mapw_hi, hi, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, 64, op->box64_rows, 2);
    if (rc == RMU_OK) rc = make_tmap_2d(&op->mapw_lo, lo, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), pitch, 64, op->box64_rows, 2);
    return rc;
}

static int gemm_variant() {
    static const int v = [] { const char* e = getenv("RMU_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
    return v;   // 0: 128x128x64 (SW128, 3 stages)   1: 256x128x32 (SW64, 4 stages)   2: 256x128x64 (SW128, 2 stages)
}

template <int MODE, int BK>
static int launch_wide(const SplitOperand& A, const SplitOperand& W, const GemmParams& p, int sms, cudaStream_t st) {
    auto kern = gemm_f16x3_wide_kernel<MODE, BK>;
    static bool attr_set = false;
    if (!attr_set) {
        RMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWideSmem)));
        attr_set = true;
    }
    const int tiles = ((p.M + kWideBM - 1) / kWideBM) * (p.N / kWideBN);
    const int grid = tiles < sms ? tiles : sms;
    ProfScope _ps(PROF_GEMM, st);
    if (BK == 32) kern<<<grid, kGemmThreads, kWideSmem, st>>>(A.map64_hi, A.map64_lo, W.map64_hi, W.map64_lo, p);
    else kern<<<grid, kGemmThreads, kWideSmem, st>>>(A.mapw_hi, A.mapw_lo, W.mapw_hi, W.mapw_lo, p);
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

template <int MODE>
static int launch_mode(const SplitOperand& A, const SplitOperand& W, const GemmParams& p, int sms, cudaStream_t st) {
    auto kern = gemm_f16x3_kernel<MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        RMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
        attr_set = true;
    }
    const int tiles = ((p.M + kGemmBM - 1) / kGemmBM) * (p.N / kGemmBN);
    const int grid = tiles < sms ? tiles : sms;
    ProfScope _ps(PROF_GEMM, st);
    kern<<<grid, kGemmThreads, kGemmSmem, st>>>(A.map_hi, A.map_lo, W.map_hi, W.map_lo, p);
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

int launch_gemm(int mode, const SplitOperand& A, const SplitOperand& W, const GemmParams& p, int sms, cudaStream_t st) {
    if (p.M <= 0) return RMU_OK;
    if (p.N % kGemmBN != 0 || p.K % kGemmBK != 0 || A.cols != p.K || W.cols != p.K || W.rows < p.N || A.rows < p.M) {
        set_error("launch_gemm: shape not supported (N % 128, K % 64)");
        return RMU_ERR_UNSUPPORTED;
    }
    if (gemm_variant() == 1 && A.box64_rows == kWideBM && W.box64_rows == kWideBN) {
        switch (mode) {
            case GEMM_BIAS_F32: return launch_wide<GEMM_BIAS_F32, 32>(A, W, p, sms, st);
            case GEMM_BIAS_GELU_SPLIT: return launch_wide<GEMM_BIAS_GELU_SPLIT, 32>(A, W, p, sms, st);
            case GEMM_BIAS_RESID_F32: return launch_wide<GEMM_BIAS_RESID_F32, 32>(A, W, p, sms, st);
            case GEMM_BIAS_SPLIT_QSCALE: return launch_wide<GEMM_BIAS_SPLIT_QSCALE, 32>(A, W, p, sms, st);
        }
    }
    if (gemm_variant() == 2 && A.box64_rows == kWideBM && W.box64_rows == kWideBN) {
        switch (mode) {
            case GEMM_BIAS_F32: return launch_wide<GEMM_BIAS_F32, 64>(A, W, p, sms, st);
            case GEMM_BIAS_GELU_SPLIT: return launch_wide<GEMM_BIAS_GELU_SPLIT, 64>(A, W, p, sms, st);
            case GEMM_BIAS_RESID_F32: return launch_wide<GEMM_BIAS_RESID_F32, 64>(A, W, p, sms, st);
            case GEMM_BIAS_SPLIT_QSCALE: return launch_wide<GEMM_BIAS_SPLIT_QSCALE, 64>(A, W, p, sms, st);
        }
    }
    switch (mode) {
        case GEMM_BIAS_F32: return launch_mode<GEMM_BIAS_F32>(A, W, p, sms, st);
        case GEMM_BIAS_GELU_SPLIT: return launch_mode<GEMM_BIAS_GELU_SPLIT>(A, W, p, sms, st);
        case GEMM_BIAS_RESID_F32: return launch_mode<GEMM_BIAS_RESID_F32>(A, W, p, sms, st);
        case GEMM_BIAS_SPLIT_QSCALE: return launch_mode<GEMM_BIAS_SPLIT_QSCALE>(A, W, p, sms, st);
    }
    set_error("launch_gemm: bad mode");
    return RMU_ERR_ARG;
}

}  // namespace rmu
