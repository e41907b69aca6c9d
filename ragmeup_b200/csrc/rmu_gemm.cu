#include "rmu_gemm.cuh"

namespace rmu {

int make_split_operand(SplitOperand* op, __half* hi, __half* lo, int64_t rows, int cols) {
    if (cols % kGemmBK != 0) { set_error("make_split_operand: K must be a multiple of 64"); return RMU_ERR_ARG; }
    op->hi = hi; op->lo = lo; op->rows = rows; op->cols = cols;
    int rc = make_tmap_2d(&op->map_hi, hi, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols),
                          static_cast<uint64_t>(cols) * sizeof(__half), kGemmBK, 128, 2);
    if (rc != RMU_OK) return rc;
    return make_tmap_2d(&op->map_lo, lo, static_cast<uint64_t>(rows), static_cast<uint64_t>(cols),
                        static_cast<uint64_t>(cols) * sizeof(__half), kGemmBK, 128, 2);
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
