// Micro-benchmark (study, not part of the library): what does ONE tcgen05.mma cost the tensor pipe of one SM, by shape
// and by operand source?  One CTA, one issuing thread, REP back-to-back MMAs into the same accumulator (operand contents
// are irrelevant: zeros), one commit, clock64 around issue and around completion.  Also the TMEM load / store round trip
// a softmax warp pays (tcgen05.ld 32x32b.x32 + wait, tcgen05.st + wait) with the tensor pipe idle and busy.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../ragmeup_b200/csrc -o mma_cost mma_cost.cu && ./mma_cost
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "rmu_ptx.cuh"

using namespace rmu;

__device__ __forceinline__ uint64_t mk_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}
__device__ __forceinline__ void mma_f16_ts_(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Case { int kind; int n; int ts; int nacc; int bmn; };   // kind 0 = f16 (K = 16), 1 = tf32 (K = 8); bmn: B is MN-major

constexpr int REP = 64;

// ELECT: the issuing thread is chosen with elect.sync (the compiler then knows it is alone and emits the UTCHMMA bare);
// otherwise with `threadIdx.x == 0`, which makes ptxas wrap every tcgen05 instruction in an ELECT / BRA.U.ANY loop
template <bool ELECT>
__global__ void __launch_bounds__(160, 1) mma_cost_kernel(const Case* cases, int ncases, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tslot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    const uint32_t a_smem = smem_u32(sm), b_smem = smem_u32(sm + 32 * 1024);
    uint32_t par = 0;
    for (int ci = 0; ci < ncases; ++ci) {
        const Case c = cases[ci];
        if (ELECT ? (threadIdx.x < 32 && elect_one()) : threadIdx.x == 0) {
            const uint32_t idesc = umma_idesc(c.kind == 0 ? 0 : 2, 128, c.n) | (c.bmn ? (1u << 16) : 0u);
            const uint64_t ad = mk_desc(a_smem, 1024, 2), bd = mk_desc(b_smem, 1024, 2);
            const long long t0 = clock64();
            for (int i = 0; i < REP; ++i) {
                const uint32_t d = tb + 256 + (i % c.nacc) * 64;
                if (c.ts) {
                    if (c.kind == 0) mma_f16_ts_(d, tb + (i & 7) * 8, bd, idesc, 1u);
                    else mma_tf32_ts(d, tb + (i & 7) * 8, bd, idesc, 1u);
                } else {
                    if (c.kind == 0) mma_f16_ss(d, ad, bd, idesc, 1u);
                    else mma_tf32_ss(d, ad, bd, idesc, 1u);
                }
            }
            const long long t1 = clock64();
            tc_commit(&bar);
            mbar_wait(&bar, par);
            const long long t2 = clock64();
            out[ci * 2] = t1 - t0;
            out[ci * 2 + 1] = t2 - t0;
        }
        par ^= 1;
        __syncthreads();
    }
    // TMEM round trips of one warp (warp 1 = lanes 32..63), tensor pipe idle, then busy (thread 0 keeps issuing TS MMAs)
    for (int busy = 0; busy < 2; ++busy) {
        if (threadIdx.x == 0 && busy) {
            const uint32_t idesc = umma_idesc(0, 128, 64) | (1u << 16);
            const uint64_t bd = mk_desc(b_smem, 1024, 2);
            for (int i = 0; i < 4 * REP; ++i) mma_f16_ts_(tb + 256, tb + 448 + (i & 3) * 8, bd, idesc, 1u);
            tc_commit(&bar);
        }
        if (threadIdx.x >= 32 && threadIdx.x < 64) {
            uint32_t v[32];
            const uint32_t ta = tmem_addr(tb, 32, 0);
            long long t0 = clock64();
            unsigned acc = 0;
            for (int i = 0; i < 8; ++i) { tmem_ld32(ta + (i & 3) * 32, v); tmem_ld_wait(); acc += v[0] ^ v[31]; v[1] = acc; }
            long long t1 = clock64();
            for (int i = 0; i < 8; ++i) { tmem_st32(ta + (i & 3) * 32, v); tmem_st_wait(); }
            long long t2 = clock64();
            for (int i = 0; i < 4; ++i) tmem_ld32(ta + i * 32, v);      // four loads in flight, one wait
            tmem_ld_wait();
            long long t3 = clock64();
            if (threadIdx.x == 32) {
                out[ncases * 2 + busy * 3] = (t1 - t0) / 8;
                out[ncases * 2 + busy * 3 + 1] = (t2 - t1) / 8;
                out[ncases * 2 + busy * 3 + 2] = (t3 - t2);
                out[ncases * 2 + 6] = v[0] + acc;
            }
        }
        if (threadIdx.x == 0 && busy) { mbar_wait(&bar, par); par ^= 1; }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

// Straight-line variant: shape and operand source are template parameters, 16 MMAs per unrolled loop body (the issue
// loop above spends ~100 cycles per MMA on its own integer work, which hides everything below that).
template <int KIND, int N, bool TS, bool BMN>
__global__ void __launch_bounds__(160, 1) mma_clean_kernel(long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tslot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (threadIdx.x < 32 && elect_one()) {
        constexpr uint32_t idesc = umma_idesc(KIND == 0 ? 0 : 2, 128, N) | (BMN ? (1u << 16) : 0u);
        const uint32_t a_smem = smem_u32(sm), b_smem = smem_u32(sm + 64 * 1024);
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                // operands walk through distinct 16 KB A tiles / 32 KB B tiles like a K loop over 64-wide blocks does
                const uint64_t ad = mk_desc(a_smem + (j >> 2) * 16384 + (j & 3) * 32, 1024, 2);
                const uint64_t bd = mk_desc(b_smem + (j >> 3) * 32768 + (j & 3) * 32, 1024, 2);
                if (TS) {
                    if (KIND == 0) mma_f16_ts_(tb + 256, tb + (j & 7) * 8, bd, idesc, 1u);
                    else mma_tf32_ts(tb + 256, tb + (j & 7) * 8, bd, idesc, 1u);
                } else {
                    if (KIND == 0) mma_f16_ss(tb + 256, ad, bd, idesc, 1u);
                    else mma_tf32_ss(tb + 256, ad, bd, idesc, 1u);
                }
            }
        }
        const long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

template <int KIND, int N, bool TS, bool BMN>
static void run_clean(long long* dout) {
    cudaFuncSetAttribute(mma_clean_kernel<KIND, N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) mma_clean_kernel<KIND, N, TS, BMN><<<1, 160, 160 * 1024>>>(dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-6s %-4d %-3s %-5s | %8.1f      | %8.1f      | floor %d\n", KIND ? "tf32" : "f16", N, TS ? "TS" : "SS", BMN ? "MN" : "K",
           double(h[0]) / 128, double(h[1]) / 128, N / 2);
}

int main() {
    const Case cases[] = {
        {0, 160, 0, 1, 0},   // S = Q K^T: f16 SS N=160
        {0, 64, 0, 1, 1},    // P V from shared memory: f16 SS N=64, B MN-major
        {0, 32, 0, 1, 1},    // f16 SS N=32
        {0, 64, 1, 1, 1},    // P V from TMEM: f16 TS N=64
        {0, 32, 1, 1, 1},    // f16 TS N=32
        {0, 64, 1, 2, 1},    // f16 TS N=64, two accumulators alternating
        {0, 256, 0, 1, 0},   // f16 SS N=256 (GEMM shape)
        {0, 256, 1, 1, 0},   // f16 TS N=256
        {1, 128, 0, 1, 0},   // scan: tf32 SS N=128
        {1, 128, 1, 1, 0},   // tf32 TS N=128 (round-1 scan form)
    };
    const int n = sizeof(cases) / sizeof(cases[0]);
    Case* dc; long long* dout;
    cudaMalloc(&dc, sizeof(cases)); cudaMalloc(&dout, (2 * n + 8) * sizeof(long long));
    cudaMemcpy(dc, cases, sizeof(cases), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(mma_cost_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(mma_cost_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    long long h[2 * 16 + 8];
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
        if (mode) mma_cost_kernel<true><<<1, 160, 96 * 1024>>>(dc, n, dout);
        else mma_cost_kernel<false><<<1, 160, 96 * 1024>>>(dc, n, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, dout, (2 * n + 8) * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("---- issuing thread chosen by %s\n", mode ? "elect.sync" : "threadIdx.x == 0");
    printf("%-6s %-4s %-3s %-5s %-5s | issue cyc/MMA | complete cyc/MMA (REP=%d)\n", "kind", "N", "A", "nacc", "B", REP);
    for (int i = 0; i < n; ++i)
        printf("%-6s %-4d %-3s %-5d %-5s | %8.1f      | %8.1f\n", cases[i].kind ? "tf32" : "f16", cases[i].n, cases[i].ts ? "TS" : "SS",
               cases[i].nacc, cases[i].bmn ? "MN" : "K", double(h[2 * i]) / REP, double(h[2 * i + 1]) / REP);
    printf("TMEM round trip of one warp (cycles): idle  ld.x32+wait %lld  st.x32+wait %lld  4 x ld.x32 then one wait %lld\n", h[2 * n], h[2 * n + 1], h[2 * n + 2]);
    printf("                                      busy  ld.x32+wait %lld  st.x32+wait %lld  4 x ld.x32 then one wait %lld\n", h[2 * n + 3], h[2 * n + 4], h[2 * n + 5]);
  }
    printf("---- straight-line issue, elect.sync, 128 MMAs: issue cyc/MMA | complete cyc/MMA\n");
    run_clean<0, 32, false, true>(dout);
    run_clean<0, 64, false, true>(dout);
    run_clean<0, 160, false, false>(dout);
    run_clean<0, 192, false, false>(dout);
    run_clean<0, 256, false, false>(dout);
    run_clean<0, 32, true, true>(dout);
    run_clean<0, 64, true, true>(dout);
    run_clean<0, 192, true, false>(dout);
    run_clean<0, 256, true, false>(dout);
    run_clean<1, 64, false, false>(dout);
    run_clean<1, 128, false, false>(dout);
    run_clean<1, 128, true, false>(dout);
    return 0;
}
