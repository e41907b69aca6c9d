// Micro-benchmark (study, not part of the library): how fast can ONE SM pull a DRAM stream into shared memory with
// cp.async (LDGSTS, 16 B per thread, manual 128-byte swizzle) compared with the TMA boxes the scan uses (measured
// there: ~25.6 B/clk/SM = one 128-byte box row per ~5 SM cycles, independent of the SM clock)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_bench stream_bench.cu && ./stream_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows of `dim` floats; a tile = 128 rows; a stage = 128 rows x 64 floats (2 K blocks of 128 B) = 32 KB
template <int STAGES, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) stream_cpasync(const float* __restrict__ x, long long rows, int dim, unsigned* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int ntiles = static_cast<int>(rows / 128);
    const int kst = dim / 64;                        // stages per tile
    int stage = 0;
    unsigned acc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int s = 0; s < kst; ++s) {
            uint8_t* dst = sm + (stage % STAGES) * 32768;
            // 2048 16-byte chunks per stage: chunk c -> K block kb = c / 1024, row r = (c % 1024) / 8, piece j = c % 8
            for (int c = threadIdx.x; c < 2048; c += THREADS) {
                const int kb = c >> 10, r = (c >> 3) & 127, j = c & 7;
                const float* src = x + (static_cast<long long>(t) * 128 + r) * dim + s * 64 + kb * 32 + j * 4;
                cp_async16(dst + kb * 16384 + r * 128 + ((j ^ (r & 7)) << 4), src);
            }
            cp_commit();
            cp_wait<STAGES - 1>();
            ++stage;
        }
    }
    cp_wait<0>();
    __syncthreads();
    acc = reinterpret_cast<unsigned*>(sm)[threadIdx.x];
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const long long rows = 10000000;
    const int dim = 384;
    float* x;
    unsigned* sink;
    cudaMalloc(&x, rows * dim * sizeof(float));
    cudaMalloc(&sink, 4);
    cudaMemset(x, 0, rows * dim * sizeof(float));
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](auto kern, int threads, int stages, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * 32768);
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            kern<<<sms, threads, stages * 32768>>>(x, rows, dim, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("%-28s threads %4d stages %d: %.3f ms  %.0f GB/s  (%s)\n", name, threads, stages, best, rows * dim * 4.0 / best / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
    run(stream_cpasync<4, 128>, 128, 4, "cp.async 16B");
    run(stream_cpasync<4, 256>, 256, 4, "cp.async 16B");
    run(stream_cpasync<6, 256>, 256, 6, "cp.async 16B");
    run(stream_cpasync<4, 512>, 512, 4, "cp.async 16B");
    run(stream_cpasync<6, 512>, 512, 6, "cp.async 16B");
    run(stream_cpasync<6, 1024>, 1024, 6, "cp.async 16B");
    (void)clk;
    return 0;
}
