// Flat (brute-force) vector index for sm_100a.
//
// Stands behind the reference's vector store (Milvus-lite FLAT / pgvector seq-scan, CPU) as it is
// driven from server/RAGHelper.py:388-404 (ctor), :431,:525 (add_documents) and :497-499
// (retriever -> col.search).  SURVEY.md §8 rows a6, a7 (+ gather for a8).
//
// Search = (1) tcgen05 coarse scan: the fp32 corpus streams HBM -> smem through TMA exactly once,
//              is consumed in place as TF32 by tcgen05.mma (queries resident in TMEM as the A
//              operand, scores accumulate in TMEM), and the epilogue keeps, per query, the best
//              KEEP rows of this CTA's row range (running threshold + warp-shuffle bitonic
//              compaction);
//          (2) finalize: radix-select the global best KEEP coarse candidates of every query,
//              re-score them exactly in fp32, sort, and CERTIFY that no excluded row can belong to
//              the top-k (coarse bound + TF32 error bound < exact k-th score);
//          (3) queries that fail the certificate (and corpora too small / shapes the tensor scan
//              does not take) run the exact fp32 CUDA-core scan.
// The ids and scores returned are therefore those of an exact fp32 brute-force search.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "rmu_common.h"
#include "rmu_ptx.cuh"

namespace rmu {

// =====================================================================================================
// exact fp32 metric — ONE definition, used by the finalize re-score and by the exact scan, so both
// paths return bit-identical scores.  Fixed summation order: four interleaved fmaf chains.
// =====================================================================================================
__device__ __forceinline__ float exact_metric(const float* __restrict__ q, const float* __restrict__ x, int D,
                                              int metric, float qnorm) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
    const int D4 = D & ~3;
    if (metric == RMU_METRIC_L2) {
        for (int d = 0; d < D4; d += 4) {
            float e0 = q[d] - x[d], e1 = q[d + 1] - x[d + 1], e2 = q[d + 2] - x[d + 2], e3 = q[d + 3] - x[d + 3];
            a0 = fmaf(e0, e0, a0); a1 = fmaf(e1, e1, a1); a2 = fmaf(e2, e2, a2); a3 = fmaf(e3, e3, a3);
        }
        for (int d = D4; d < D; ++d) { float e = q[d] - x[d]; a0 = fmaf(e, e, a0); }
        return (a0 + a1) + (a2 + a3);
    }
    if (metric == RMU_METRIC_IP) {
        for (int d = 0; d < D4; d += 4) {
            a0 = fmaf(q[d], x[d], a0); a1 = fmaf(q[d + 1], x[d + 1], a1);
            a2 = fmaf(q[d + 2], x[d + 2], a2); a3 = fmaf(q[d + 3], x[d + 3], a3);
        }
        for (int d = D4; d < D; ++d) a0 = fmaf(q[d], x[d], a0);
        return (a0 + a1) + (a2 + a3);
    }
    // cosine
    for (int d = 0; d < D4; d += 4) {
        float x0 = x[d], x1 = x[d + 1], x2 = x[d + 2], x3 = x[d + 3];
        a0 = fmaf(q[d], x0, a0); a1 = fmaf(q[d + 1], x1, a1); a2 = fmaf(q[d + 2], x2, a2); a3 = fmaf(q[d + 3], x3, a3);
        n0 = fmaf(x0, x0, n0); n1 = fmaf(x1, x1, n1); n2 = fmaf(x2, x2, n2); n3 = fmaf(x3, x3, n3);
    }
    for (int d = D4; d < D; ++d) { a0 = fmaf(q[d], x[d], a0); n0 = fmaf(x[d], x[d], n0); }
    float ip = (a0 + a1) + (a2 + a3);
    float xn = sqrtf((n0 + n1) + (n2 + n3));
    float den = qnorm * xn;
    return den > 0.f ? ip / den : 0.f;
}

// The same arithmetic as exact_metric (same four fmaf chains per (query,row), so bit-identical results)
// for QT queries at once: the row is read once for all of them.
template <int QT>
__device__ __forceinline__ void exact_metric_multi(const float* __restrict__ qs /*[QT][D]*/, const float* __restrict__ x,
                                                   int D, int metric, const float* __restrict__ qnorm, float (&out)[QT]) {
    float a[QT][4];
#pragma unroll
    for (int i = 0; i < QT; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
    const int D4 = D & ~3;
    if (metric == RMU_METRIC_L2) {
        for (int d = 0; d < D4; d += 4) {
            const float x0 = x[d], x1 = x[d + 1], x2 = x[d + 2], x3 = x[d + 3];
#pragma unroll
            for (int i = 0; i < QT; ++i) {
                const float* q = qs + i * D;
                const float e0 = q[d] - x0, e1 = q[d + 1] - x1, e2 = q[d + 2] - x2, e3 = q[d + 3] - x3;
                a[i][0] = fmaf(e0, e0, a[i][0]); a[i][1] = fmaf(e1, e1, a[i][1]);
                a[i][2] = fmaf(e2, e2, a[i][2]); a[i][3] = fmaf(e3, e3, a[i][3]);
            }
        }
        for (int d = D4; d < D; ++d) {
#pragma unroll
            for (int i = 0; i < QT; ++i) { const float e = qs[i * D + d] - x[d]; a[i][0] = fmaf(e, e, a[i][0]); }
        }
#pragma unroll
        for (int i = 0; i < QT; ++i) out[i] = (a[i][0] + a[i][1]) + (a[i][2] + a[i][3]);
        return;
    }
    for (int d = 0; d < D4; d += 4) {
        const float x0 = x[d], x1 = x[d + 1], x2 = x[d + 2], x3 = x[d + 3];
        if (metric == RMU_METRIC_COSINE) { n0 = fmaf(x0, x0, n0); n1 = fmaf(x1, x1, n1); n2 = fmaf(x2, x2, n2); n3 = fmaf(x3, x3, n3); }
#pragma unroll
        for (int i = 0; i < QT; ++i) {
            const float* q = qs + i * D;
            a[i][0] = fmaf(q[d], x0, a[i][0]); a[i][1] = fmaf(q[d + 1], x1, a[i][1]);
            a[i][2] = fmaf(q[d + 2], x2, a[i][2]); a[i][3] = fmaf(q[d + 3], x3, a[i][3]);
        }
    }
    for (int d = D4; d < D; ++d) {
        if (metric == RMU_METRIC_COSINE) n0 = fmaf(x[d], x[d], n0);
#pragma unroll
        for (int i = 0; i < QT; ++i) a[i][0] = fmaf(qs[i * D + d], x[d], a[i][0]);
    }
    const float xn = sqrtf((n0 + n1) + (n2 + n3));
#pragma unroll
    for (int i = 0; i < QT; ++i) {
        const float ip = (a[i][0] + a[i][1]) + (a[i][2] + a[i][3]);
        if (metric == RMU_METRIC_IP) out[i] = ip;
        else { const float den = qnorm[i] * xn; out[i] = den > 0.f ? ip / den : 0.f; }
    }
}

// "larger is better" key of a metric value
__device__ __forceinline__ float metric_to_rank(float v, int metric) { return metric == RMU_METRIC_L2 ? -v : v; }

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// =====================================================================================================
// per-row statistics at insert time: cosine scale 1/||x||, L2 bias -0.5||x||^2, running max ||x||
// =====================================================================================================
__global__ void row_stats_kernel(const float* __restrict__ x, long long row0, long long n, int D, int metric,
                                 float* __restrict__ rscale, float* __restrict__ rbias, unsigned* __restrict__ max_norm_bits,
                                 unsigned* __restrict__ max_dev_bits) {
    int warps_per_block = blockDim.x >> 5;
    long long r = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* p = x + (row0 + r) * D;
    float s = 0.f;
    for (int d = lane_id(); d < D; d += 32) s = fmaf(p[d], p[d], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane_id() == 0) {
        float nrm = sqrtf(s);
        if (metric == RMU_METRIC_COSINE) rscale[row0 + r] = nrm > 0.f ? 1.f / nrm : 0.f;
        if (metric == RMU_METRIC_L2) rbias[row0 + r] = -0.5f * s;
        atomicMax(max_norm_bits, __float_as_uint(nrm));  // non-negative floats order like uints
        atomicMax(max_dev_bits, __float_as_uint(fabsf(s - 1.0f)));   // how far from unit norm the corpus gets
    }
}

// =====================================================================================================
// warp-shuffle bitonic sort, descending, of 32*E u64 keys (element i = e*32 + lane)
// =====================================================================================================
template <int E>
__device__ __forceinline__ void warp_bitonic_desc(unsigned long long (&v)[E]) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & je) == 0) {
                        const int e2 = e | je;
                        const bool desc = (((e << 5) & k) == 0);
                        unsigned long long a = v[e], b = v[e2];
                        const bool sw = desc ? (a < b) : (a > b);
                        v[e] = sw ? b : a;
                        v[e2] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned i = (static_cast<unsigned>(e) << 5) | lane;
                    unsigned long long o = __shfl_xor_sync(0xffffffffu, v[e], j);
                    const bool lower = (lane & j) == 0;
                    const bool desc = ((i & k) == 0);
                    const bool keepmax = (lower == desc);
                    v[e] = keepmax ? (v[e] > o ? v[e] : o) : (v[e] < o ? v[e] : o);
                }
            }
        }
    }
}

// Warp-cooperative compaction of the per-thread (= per-query) candidate lists of the lanes that ask
// for it: sort the list descending, keep the best KEEP, raise that lane's threshold.
template <int KEEP, int CAP>
__device__ __forceinline__ void warp_compact(unsigned long long* mybuf, int& cnt, float& tau, bool need) {
    constexpr int E = CAP / 32;
    unsigned mask = __ballot_sync(0xffffffffu, need);
    if (mask == 0) return;
    __syncwarp();
    const unsigned lane = lane_id();
    while (mask) {
        const int L = __ffs(mask) - 1;
        mask &= mask - 1;
        unsigned long long* b = reinterpret_cast<unsigned long long*>(
            __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(mybuf), L));
        const int n = __shfl_sync(0xffffffffu, cnt, L);
        unsigned long long v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = e * 32 + lane;
            v[e] = (i < n) ? b[i] : 0ull;
        }
        warp_bitonic_desc<E>(v);
#pragma unroll
        for (int e = 0; e < KEEP / 32; ++e) b[e * 32 + lane] = v[e];
        const unsigned long long kth = __shfl_sync(0xffffffffu, v[KEEP / 32 - 1], 31);
        if (lane == static_cast<unsigned>(L)) {
            if (n >= KEEP) { cnt = KEEP; tau = key_score(kth); }
            // n < KEEP: list is now sorted, count unchanged, threshold unchanged
        }
    }
    __syncwarp();
}

// =====================================================================================================
// (1) tcgen05 coarse scan
// =====================================================================================================
// list slots per (CTA, query): a compaction leaves KEEP entries and is triggered above CAP - 32, so the
// capacity must leave real headroom (KEEP = 32 with 64 slots would compact after every single insert)
__host__ __device__ constexpr int scan_cap(int keep) { return keep < 64 ? 128 : 2 * keep; }

constexpr int kScanQ = 128;        // queries per launch = MMA M = TMEM lanes
constexpr int kScanACols = 384;    // TMEM columns reserved for the query block (max dim of this path)
constexpr int kScanThreads = 192;  // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 epilogue

struct ScanParams {
    const float* q;        // [nq_total, dim]
    int q0, nq;            // this launch: queries q0 .. q0+nq-1, nq <= 128
    int dim;
    long long n;           // rows in the index
    int ntiles;            // ceil(n / BN)
    const float* rscale;   // nullable (cosine)
    const float* rbias;    // nullable (L2)
    unsigned long long* lists;  // [gridDim.x][128][CAP]: unsorted candidates of (CTA, query), counts[] of them valid
    int* counts;                // [gridDim.x][128]
    float* dbg;            // diagnostics: CTA 0 dumps the raw accumulators of its first tile [128][BN]
    int ablate;            // profiling only: bit0 skip MMA issue, bit1 skip epilogue work, bit2 skip TMEM loads
    // threshold exchange: phase 0 = whole range, thresholds start at -inf; phase 1 = only the first `lead`
    // tiles of every CTA, run with a small KEEP purely to estimate thresholds; phase 2 = whole range again,
    // every query starting from tau0 = the KEEP-th best key over the union of all CTAs' phase-1 lists (a
    // subset of the rows, hence a valid lower bound of the final KEEP-th best key).
    int phase, lead;
    const float* tau0;     // [nq_total]
    // K split for 384 < dim <= 768 (the query block only has 384 TMEM columns): pass 1 multiplies columns
    // [0, 384) and stores raw partial scores, pass 2 multiplies [384, dim) and adds them before thresholding.
    int kcol0, kdim;       // first column and number of columns of this launch
    int kpass;             // 0 single pass, 1 write partials only, 2 add partials then continue as usual
    float* partial;        // [128][npad] raw partial scores of the current query block
    long long npad;        // row pitch of `partial` (multiple of 32)
};

// BN rows per tile (= MMA N), NBUF TMEM accumulators, NSLAB pipeline stages, each stage = KD K-blocks
// (one TMA op; KD > 1 uses the 3-D (32, rows, kblock) tensor map and needs dim % 32 == 0)
template <int BN, int NBUF, int NSLAB, int KD, bool TMA3D, int KEEP>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_tf32_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
    constexpr int CAP = scan_cap(KEEP);
    constexpr int SLAB_BYTES = BN * 128 * KD;
    constexpr uint32_t IDESC = umma_idesc(2 /*tf32*/, 128, BN);
    static_assert(BN * NBUF <= 512 - kScanACols, "accumulators must fit beside the query block in TMEM");

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by OFFSET into the shared array: a pointer round-trip through an integer would lose the
    // shared address space and turn every staging access into a generic LD/ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* slabs = smem;
    float* stage = reinterpret_cast<float*>(smem + NSLAB * SLAB_BYTES);          // [32][128] epilogue staging
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSLAB * SLAB_BYTES + 32 * 128 * sizeof(float));
    uint64_t* empty = full + NSLAB;
    uint64_t* acc_full = empty + NSLAB;
    uint64_t* acc_empty = acc_full + NBUF;
    uint64_t* a_ready = acc_empty + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const int KB = (p.kdim + 31) / 32;  // 128-byte K blocks of this launch's column range

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLAB; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        mbar_init(a_ready, 128);
        fence_mbar_init();
        prefetch_tmap(&tmap);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // contiguous tile range of this CTA
    int t0 = static_cast<int>((static_cast<long long>(blockIdx.x) * p.ntiles) / gridDim.x);
    int t1 = static_cast<int>((static_cast<long long>(blockIdx.x + 1) * p.ntiles) / gridDim.x);
    if (p.phase == 1) t1 = min(t1, t0 + p.lead);

    // ---- query block -> TMEM (A operand): lane = query, column = dimension, zero padded
    const int quad = warp & 3;  // TMEM lane quadrant this warp may touch
    if (warp >= 2) {
        const int qi = quad * 32 + lane;
        const float* qrow = (qi < p.nq) ? p.q + static_cast<long long>(p.q0 + qi) * p.dim : nullptr;
        // dim % 4 == 0 and 16-byte aligned rows (checked by the host): two float4 loads per 8 columns
        for (int c = 0; c < KB * 4; ++c) {
            uint32_t r[8];
            const int d = c * 8;
            float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
            if (qrow != nullptr && d < p.kdim) v0 = *reinterpret_cast<const float4*>(qrow + p.kcol0 + d);
            if (qrow != nullptr && d + 4 < p.kdim) v1 = *reinterpret_cast<const float4*>(qrow + p.kcol0 + d + 4);
            r[0] = __float_as_uint(v0.x); r[1] = __float_as_uint(v0.y); r[2] = __float_as_uint(v0.z); r[3] = __float_as_uint(v0.w);
            r[4] = __float_as_uint(v1.x); r[5] = __float_as_uint(v1.y); r[6] = __float_as_uint(v1.z); r[7] = __float_as_uint(v1.w);
            tmem_st8(tmem_addr(tmem_base, quad * 32, c * 8), r);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(a_ready);                      // the MMA issuer waits for all 128 query rows; TMA streams meanwhile
    }

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                for (int kb = 0; kb < KB; kb += KD) {
                    mbar_wait(&empty[slot], phase ^ 1);
                    mbar_arrive_expect_tx(&full[slot], SLAB_BYTES);
                    if (TMA3D) {
                        tma_load_3d(slabs + slot * SLAB_BYTES, &tmap, 0, t * BN, kb, &full[slot], kEvictFirst);
                    } else {
                        // KD boxes of {128 B, BN rows} credited to one barrier (K-blocks past the row end are
                        // out of bounds and arrive as zeros, still counting their bytes)
#pragma unroll
                        for (int kk = 0; kk < KD; ++kk)
                            tma_load_2d(slabs + slot * SLAB_BYTES + kk * (BN * 128), &tmap, p.kcol0 + (kb + kk) * 32, t * BN,
                                        &full[slot], kEvictFirst);
                    }
                    if (++slot == NSLAB) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            mbar_wait(a_ready, 0);                 // query block is in TMEM
            tc_fence_after();
            for (int t = t0; t < t1; ++t) {
                const int i = t - t0;
                const int buf = i % NBUF;
                const uint32_t use = static_cast<uint32_t>(i / NBUF);
                mbar_wait(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + kScanACols + buf * BN;
                for (int kb0 = 0; kb0 < KB; kb0 += KD) {
                    mbar_wait(&full[slot], phase);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KD; ++kk) {
                        const int kb = kb0 + kk;
                        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(slabs + slot * SLAB_BYTES + kk * (BN * 128)));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (kb * 32 + k * 8 < p.kdim && !(p.ablate & 1)) {
                                mma_tf32_ts(d_addr, tmem_base + kb * 32 + k * 8, bdesc + static_cast<uint64_t>(k * 2),
                                            IDESC, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    tc_commit(&empty[slot]);
                    if (++slot == NSLAB) { slot = 0; phase ^= 1; }
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else {
        // =========================== epilogue: thread = query ===========================
        const int qi = quad * 32 + lane;
        const int te = qi;                       // this thread's column in the staging buffer
        const bool live = qi < p.nq;
        unsigned long long* mybuf = p.lists + (static_cast<long long>(blockIdx.x) * kScanQ + qi) * CAP;
        int cnt = 0;
        float tau = live ? -INFINITY : INFINITY;
        if (p.phase == 2 && live) tau = p.tau0[p.q0 + qi];
        const bool has_sb = (p.rscale != nullptr) || (p.rbias != nullptr);
        for (int t = t0; t < t1; ++t) {
            const int i = t - t0;
            const int buf = i % NBUF;
            const uint32_t use = static_cast<uint32_t>(i / NBUF);
            mbar_wait(&acc_full[buf], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                if (p.ablate & 4) break;
                uint32_t r[32];
                tmem_ld32(tmem_addr(tmem_base, quad * 32, kScanACols + buf * BN + c * 32), r);
                tmem_ld_wait();
                if (c == BN / 32 - 1) {          // last TMEM read of this tile: hand the accumulator back early
                    tc_fence_before();
                    mbar_arrive(&acc_empty[buf]);
                }
                if (p.ablate & 2) continue;
                const long long row0 = static_cast<long long>(t) * BN + c * 32;
                if (p.kpass == 1) {                       // first K half: park the raw partial scores
                    if (live && row0 < p.n) {
                        float4* dst = reinterpret_cast<float4*>(p.partial + static_cast<long long>(qi) * p.npad + row0);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            dst[g4] = make_float4(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]),
                                                  __uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3]));
                    }
                    continue;
                }
                if (p.kpass == 2 && live && row0 < p.n) { // second K half: add what the first half parked
                    const float4* src = reinterpret_cast<const float4*>(p.partial + static_cast<long long>(qi) * p.npad + row0);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 v = src[g4];
                        r[4 * g4] = __float_as_uint(__uint_as_float(r[4 * g4]) + v.x);
                        r[4 * g4 + 1] = __float_as_uint(__uint_as_float(r[4 * g4 + 1]) + v.y);
                        r[4 * g4 + 2] = __float_as_uint(__uint_as_float(r[4 * g4 + 2]) + v.z);
                        r[4 * g4 + 3] = __float_as_uint(__uint_as_float(r[4 * g4 + 3]) + v.w);
                    }
                }
                if (p.dbg != nullptr && blockIdx.x == 0 && t == t0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) p.dbg[qi * BN + c * 32 + j] = __uint_as_float(r[j]);
                }
                if (row0 < p.n) {
                    const int valid = static_cast<int>(min(static_cast<long long>(32), p.n - row0));
                    float key[32];
                    if (has_sb) {
                        // per-row scale / bias of this 32-row chunk: one coalesced load, shuffled out
                        const float sc = (p.rscale != nullptr && static_cast<int>(lane) < valid) ? __ldg(p.rscale + row0 + lane) : 1.f;
                        const float bi = (p.rbias != nullptr && static_cast<int>(lane) < valid) ? __ldg(p.rbias + row0 + lane) : 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            key[j] = fmaf(__uint_as_float(r[j]), __shfl_sync(0xffffffffu, sc, j), __shfl_sync(0xffffffffu, bi, j));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) key[j] = __uint_as_float(r[j]);
                    }
                    // branch-free pass mask (bit j = row j beats this query's running threshold); the
                    // insert path is entered warp-uniformly and walks only the set bits, reading the
                    // scores back from a per-thread smem column (no dynamic register indexing).
                    unsigned m = 0u;
#pragma unroll
                    for (int j = 0; j < 32; ++j) m |= (key[j] > tau) ? (1u << j) : 0u;
                    if (valid < 32) m &= (1u << valid) - 1u;
                    if (__any_sync(0xffffffffu, m != 0u)) {
                        float* col = stage + te;
#pragma unroll
                        for (int j = 0; j < 32; ++j) col[j * 128] = key[j];
                        while (m) {
                            const int j = __ffs(m) - 1;
                            m &= m - 1;
                            mybuf[cnt++] = make_key(col[j * 128], static_cast<uint32_t>(row0 + j));
                        }
                    }
                }
                warp_compact<KEEP, CAP>(mybuf, cnt, tau, cnt > CAP - 32);
            }
            if (p.ablate & 4) {
                tc_fence_before();
                mbar_arrive(&acc_empty[buf]);
            }
        }
        // final: the list stays as it is (unsorted, <= CAP entries, a superset of this CTA's KEEP best above the
        // threshold); the selection kernels read `counts` entries.  (Sorting 128 lists per CTA here, one lane at a
        // time, was a fixed ~0.1 ms tail on every launch: it dominated 1M-row shards.)
        // Long lists (KEEP >= 128) are still sorted and cut to KEEP here: reading 2 * KEEP entries of every CTA in the
        // 8-pass radix select of finalize would cost more than it saves.
        if (p.kpass != 1) {
            if (KEEP >= 128) {
                warp_compact<KEEP, CAP>(mybuf, cnt, tau, true);
                for (int e = cnt; e < KEEP; ++e) mybuf[e] = 0ull;
            } else {
                p.counts[blockIdx.x * kScanQ + qi] = cnt;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================
// (3) exact fp32 scan: one CTA = one chunk of rows x one (flagged) query -> sorted top list
// =====================================================================================================
constexpr int kChunk = 2048;

__device__ __forceinline__ void block_bitonic_desc(unsigned long long* s, int n /*pow2*/) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
                const int l = i | j;
                const bool desc = ((i & k) == 0);
                unsigned long long a = s[i], b = s[l];
                const bool sw = desc ? (a < b) : (a > b);
                if (sw) { s[i] = b; s[l] = a; }
            }
        }
    }
    __syncthreads();
}

struct ExactParams {
    const float* x; long long n; int dim; int metric;
    const float* q;            // [nq_total, dim]
    const int* qmap;           // [*nsel] query indices to process (nullable = identity over nq_total)
    const int* nsel;           // device count of selected queries (nullable -> nq_total)
    int nq_total;
    unsigned long long* lists; // [nchunks][nq_total][keep]
    int keep;                  // pow2 <= kChunk
};

constexpr int kExactQT = 8;    // queries scored per pass over a chunk of rows

__global__ void __launch_bounds__(256) exact_scan_kernel(const ExactParams p) {
    __shared__ unsigned long long keys[kChunk];
    __shared__ float red[32];
    __shared__ float qn[kExactQT];
    extern __shared__ float esm[];                 // qs [QT][dim], then sc [QT][kChunk]
    float* qs = esm;
    float* sc = esm + kExactQT * p.dim;
    const int nsel = p.nsel ? *p.nsel : p.nq_total;
    const long long row0 = static_cast<long long>(blockIdx.x) * kChunk;
    const int ngroups = (nsel + kExactQT - 1) / kExactQT;
    for (int grp = blockIdx.y; grp < ngroups; grp += gridDim.y) {
        const int f0 = grp * kExactQT;
        const int nq = min(kExactQT, nsel - f0);
        __syncthreads();
        for (int i = 0; i < kExactQT; ++i) {
            float part = 0.f;
            if (i < nq) {
                const int qg = p.qmap ? p.qmap[f0 + i] : f0 + i;
                for (int d = threadIdx.x; d < p.dim; d += blockDim.x) {
                    const float v = p.q[static_cast<long long>(qg) * p.dim + d];
                    qs[i * p.dim + d] = v;
                    part = fmaf(v, v, part);
                }
            } else {
                for (int d = threadIdx.x; d < p.dim; d += blockDim.x) qs[i * p.dim + d] = 0.f;
            }
            const float nrm = sqrtf(block_sum(part, red));
            if (threadIdx.x == 0) qn[i] = nrm;
        }
        __syncthreads();
        for (int r = threadIdx.x; r < kChunk; r += blockDim.x) {
            const long long row = row0 + r;
            if (row < p.n) {
                float v[kExactQT];
                exact_metric_multi<kExactQT>(qs, p.x + row * p.dim, p.dim, p.metric, qn, v);
#pragma unroll
                for (int i = 0; i < kExactQT; ++i) sc[i * kChunk + r] = v[i];
            }
        }
        __syncthreads();
        for (int i = 0; i < nq; ++i) {
            for (int r = threadIdx.x; r < kChunk; r += blockDim.x) {
                const long long row = row0 + r;
                keys[r] = row < p.n ? make_key(metric_to_rank(sc[i * kChunk + r], p.metric), static_cast<uint32_t>(row)) : 0ull;
            }
            block_bitonic_desc(keys, kChunk);
            unsigned long long* out = p.lists + (static_cast<long long>(blockIdx.x) * p.nq_total + f0 + i) * p.keep;
            for (int e = threadIdx.x; e < p.keep; e += blockDim.x) out[e] = keys[e];
            __syncthreads();
        }
    }
}

// threads of the selection kernels (select_tau, finalize): their 8 radix passes over nlists * len keys are latency-bound,
// one CTA per query
constexpr int kSelThreads = 512;

// Block-wide radix select (8 bits per pass, most significant first) over passes [pass0, pass1): after pass 8 the
// state holds the ksel-th largest non-zero key among load_key(0..total); after pass 4 its 32 score bits (the low
// word still zero) and, in *s_ties, how many keys share those score bits.  *s_remaining < 0 means fewer than ksel
// keys exist ("keep everything").  pass0 == 0 initialises the state.  All threads call.
template <typename LoadKey>
__device__ __forceinline__ void block_radix_passes(LoadKey load_key, long long total, int ksel, int pass0, int pass1, int* hist,
                                                   unsigned long long* s_prefix, int* s_remaining, int* s_ties) {
    const int tid = threadIdx.x;
    if (pass0 == 0) {
        if (tid == 0) { *s_prefix = 0ull; *s_remaining = ksel; *s_ties = 0; }
        __syncthreads();
    }
    for (int pass = pass0; pass < pass1; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        const int remaining = *s_remaining;         // < 0: fewer than ksel keys exist, keep all
        const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        if (remaining >= 0) {
            for (long long idx = tid; idx < total; idx += blockDim.x) {
                const unsigned long long key = load_key(idx);
                if (key != 0ull && (key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
            }
        }
        __syncthreads();
        if (tid == 0 && remaining >= 0) {
            int rem = remaining;
            int b = 255;
            for (; b >= 0; --b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            if (b < 0) { *s_remaining = -1; }
            else { *s_prefix = prefix | (static_cast<unsigned long long>(b) << shift); *s_remaining = rem; *s_ties = hist[b]; }
        }
        __syncthreads();
    }
}

// the ksel-th largest non-zero key, or 1 ("keep everything") when fewer than ksel exist.  Ties on the 32 score bits
// are rare: when the keys that share the selected score are all needed (s_ties == s_remaining after four passes) the
// threshold is that score with a zero row word and the four passes over the row bits are skipped.
template <typename LoadKey>
__device__ __forceinline__ unsigned long long block_radix_select(LoadKey load_key, long long total, int ksel, int* hist,
                                                                 unsigned long long* s_prefix, int* s_remaining, int* s_ties) {
    block_radix_passes(load_key, total, ksel, 0, 4, hist, s_prefix, s_remaining, s_ties);
    if (*s_remaining >= 0 && *s_ties != *s_remaining)        // uniform: shared state, read after the barrier
        block_radix_passes(load_key, total, ksel, 4, 8, hist, s_prefix, s_remaining, s_ties);
    return *s_remaining < 0 ? 1ull : *s_prefix;
}

// per-query global threshold for phase 2 of the scan: score of the keep-th best key over all CTAs'
// phase-1 lists (-inf when fewer than keep candidates exist yet)
__global__ void __launch_bounds__(kSelThreads) select_tau_kernel(const unsigned long long* __restrict__ lists,
                                                         const int* __restrict__ counts, int nlists, int lstride,
                                                         int len, int keep, int q0, float* __restrict__ tau0) {
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_ties;
    const int f = blockIdx.x;
    const long long total = static_cast<long long>(nlists) * len;
    const int len_shift = __ffs(len) - 1;
    auto load_key = [&](long long idx) -> unsigned long long {
        const long long l = idx >> len_shift;
        const int e = static_cast<int>(idx & (len - 1));
        return e < counts[l * kScanQ + f] ? lists[(l * kScanQ + f) * lstride + e] : 0ull;
    };
    block_radix_passes(load_key, total, keep, 0, 4, hist, &s_prefix, &s_remaining, &s_ties);   // the score bits suffice
    if (threadIdx.x == 0) tau0[q0 + f] = s_remaining < 0 ? -INFINITY : key_score(s_prefix);
}

// =====================================================================================================
// (2) finalize: global selection of the best KSEL candidates, exact re-score, sort, certificate
// =====================================================================================================
struct FinalizeParams {
    const unsigned long long* lists;
    int nlists;        // lists per query
    int qstride;       // queries per list block (128 for the tensor scan, nq_total for the exact scan)
    int lstride;       // entries between consecutive queries' lists
    int len;           // entries read per list, pow2 (zero padded, or bounded by counts[])
    const int* counts; // nullable: valid entries of list (l, query slot) = counts[l * qstride + slot]
    int ksel;          // candidates to select (<= 1024, <= len for a valid certificate)
    const float* x; long long n; int dim; int metric;
    const float* q;    // [nq_total, dim]
    int q0;            // blockIdx.x + q0 = query (when qmap == nullptr)
    const int* qmap;   // exact mode: blockIdx.x -> query through qmap
    const int* nsel;   // exact mode: number of valid blockIdx.x
    int exact;         // 1: lists hold exact keys, no certificate
    int unit_rows;     // 1: every row has ||x||^2 = 1 +- 1e-6 and the coarse keys are plain inner products
    int k;
    long long id_offset;
    const unsigned* max_norm_bits;
    float eps_rel;
    float* out_scores; long long* out_ids;   // [nq_total, k]
    int* flags;        // [nq_total] 1 = certificate failed
};

__global__ void __launch_bounds__(kSelThreads) finalize_kernel(const FinalizeParams p) {
    __shared__ unsigned long long cand[1024];
    __shared__ float cval[1024];
    __shared__ int hist[256];
    __shared__ float red[32];
    __shared__ int s_ncand;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_ties;
    extern __shared__ float qs[];  // [dim]

    const int f = blockIdx.x;
    if (p.nsel != nullptr && f >= *p.nsel) return;
    const int qg = p.qmap ? p.qmap[f] : p.q0 + f;
    const int qslot = p.exact ? f : f;  // position of this query inside a list block
    const int tid = threadIdx.x;

    float part = 0.f;
    for (int d = tid; d < p.dim; d += blockDim.x) {
        float v = p.q[static_cast<long long>(qg) * p.dim + d];
        qs[d] = v;
        part = fmaf(v, v, part);
    }
    const float qnorm = sqrtf(block_sum(part, red));

    const long long total = static_cast<long long>(p.nlists) * p.len;
    const int len_shift = __ffs(p.len) - 1;
    auto load_key = [&](long long idx) -> unsigned long long {
        const long long l = idx >> len_shift;
        const int e = static_cast<int>(idx & (p.len - 1));
        if (p.counts != nullptr && e >= p.counts[l * p.qstride + qslot]) return 0ull;
        return p.lists[(l * p.qstride + qslot) * p.lstride + e];
    };

    const unsigned long long T = block_radix_select(load_key, total, p.ksel, hist, &s_prefix, &s_remaining, &s_ties);
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    // ---- collect keys >= T
    long long nonzero_local = 0;
    for (long long idx = tid; idx < total; idx += blockDim.x) {
        const unsigned long long key = load_key(idx);
        if (key != 0ull) {
            ++nonzero_local;
            if (key >= T) {
                const int pos = atomicAdd(&s_ncand, 1);
                if (pos < 1024) cand[pos] = key;
            }
        }
    }
    const float nonzero = block_sum(static_cast<float>(nonzero_local), red);
    __syncthreads();
    const int ncand = min(s_ncand, p.ksel);

    // ---- exact re-score
    for (int c = tid; c < 1024; c += blockDim.x) {
        if (c < ncand) {
            const uint32_t row = key_row(cand[c]);
            const float v = exact_metric(qs, p.x + static_cast<long long>(row) * p.dim, p.dim, p.metric, qnorm);
            cval[c] = v;
            cand[c] = make_key(metric_to_rank(v, p.metric), row);
        } else {
            cand[c] = 0ull;
        }
    }
    int n2 = 32;
    while (n2 < ncand) n2 <<= 1;
    block_bitonic_desc(cand, n2);

    // ---- outputs
    const float missing = p.metric == RMU_METRIC_L2 ? INFINITY : -INFINITY;
    for (int j = tid; j < p.k; j += blockDim.x) {
        float s = missing;
        long long id = -1;
        if (j < ncand) {
            const float rk = key_score(cand[j]);
            s = p.metric == RMU_METRIC_L2 ? -rk : rk;
            id = p.id_offset + key_row(cand[j]);
        }
        p.out_scores[static_cast<long long>(qg) * p.k + j] = s;
        p.out_ids[static_cast<long long>(qg) * p.k + j] = id;
    }
    // ---- certificate: every row outside the candidate set has coarse key <= score(T); its exact key is
    //      at most eps above.  The k-th exact candidate must beat that bound strictly.
    if (tid == 0 && p.flags != nullptr) {
        int flag = 0;
        if (!p.exact && nonzero >= static_cast<float>(p.ksel)) {
            const int kk = min(p.k, ncand);
            const float rk = key_score(cand[kk - 1]);  // rank value of the k-th exact result
            float kth_key;   // in the units of the coarse key
            float scale;
            const float xmax = __uint_as_float(*p.max_norm_bits);
            if (p.metric == RMU_METRIC_IP) { kth_key = rk; scale = qnorm * xmax; }
            else if (p.metric == RMU_METRIC_COSINE) { kth_key = rk * qnorm; scale = qnorm; }
            else { kth_key = 0.5f * (qnorm * qnorm + rk) + (p.unit_rows ? 0.5f : 0.f); scale = qnorm * xmax; }  // rk = -dist
            const float bound = key_score(T);
            // unit_rows: cosine / L2 keys were scanned as inner products, exact to 1e-6 (||x||^2 = 1 +- 1e-6)
            const float eps = p.eps_rel * scale + 1e-6f * (1.f + fabsf(kth_key)) + (p.unit_rows ? 4e-6f * (1.f + qnorm) : 0.f);
            if (!(bound + eps < kth_key)) flag = 1;
            if (ncand < p.k) flag = 1;
        }
        p.flags[qg] = flag;
    }
}

// flags[nq] -> qmap (ordered, compacted indices of flagged queries) + nsel; launched with one warp
__global__ void compact_flags_kernel(const int* __restrict__ flags, int nq, int* __restrict__ qmap, int* __restrict__ nsel) {
    int count = 0;
    for (int base = 0; base < nq; base += 32) {
        const int i = base + static_cast<int>(lane_id());
        const bool f = i < nq && flags[i] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) qmap[count + __popc(m & ((1u << lane_id()) - 1u))] = i;
        count += __popc(m);
    }
    if (lane_id() == 0) *nsel = count;
}

// =====================================================================================================
// shard merge (after the all-gather of per-shard results): [R, nq, k] -> [nq, k]
// =====================================================================================================
__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ scores, const long long* __restrict__ ids,
                                                    int R, int nq, int k, int metric, float* __restrict__ out_s,
                                                    long long* __restrict__ out_i) {
    extern __shared__ unsigned char sm[];
    // entries sorted by (rank desc, id asc); ids are 64-bit so keep them beside a (rank, slot) key
    const int n = R * k;
    int n2 = 32;
    while (n2 < n) n2 <<= 1;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm);   // [n2] (ordered rank << 32 | ~slot)
    long long* sid = reinterpret_cast<long long*>(keys + n2);               // [n]
    float* sval = reinterpret_cast<float*>(sid + n);                        // [n]
    const int qi = blockIdx.x;
    // stable rule shared with the oracle: rank desc, then global id asc.  Slots are ordered by id
    // inside equal ranks by a final fix-up pass below (ties across shards are rare: duplicates).
    for (int t = threadIdx.x; t < n2; t += blockDim.x) {
        unsigned long long key = 0ull;
        if (t < n) {
            const int r = t / k, j = t % k;
            const long long src = (static_cast<long long>(r) * nq + qi) * k + j;
            const long long id = ids[src];
            const float v = scores[src];
            sid[t] = id;
            sval[t] = v;
            if (id >= 0) key = (static_cast<unsigned long long>(f32_to_ordered(metric_to_rank(v, metric))) << 32) |
                               static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(t));
        }
        keys[t] = key;
    }
    block_bitonic_desc(keys, n2);
    // fix-up: within runs of equal rank order by id ascending (insertion sort by one thread; runs are tiny)
    if (threadIdx.x == 0) {
        int a = 0;
        while (a < n && keys[a] != 0ull) {
            int b = a + 1;
            while (b < n && keys[b] != 0ull && (keys[b] >> 32) == (keys[a] >> 32)) ++b;
            for (int i = a + 1; i < b; ++i) {
                unsigned long long ki = keys[i];
                long long idi = sid[0xFFFFFFFFu - static_cast<unsigned>(ki)];
                int j = i - 1;
                while (j >= a && sid[0xFFFFFFFFu - static_cast<unsigned>(keys[j])] > idi) { keys[j + 1] = keys[j]; --j; }
                keys[j + 1] = ki;
            }
            a = b;
        }
    }
    __syncthreads();
    const float missing = metric == RMU_METRIC_L2 ? INFINITY : -INFINITY;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        float s = missing;
        long long id = -1;
        if (j < n && keys[j] != 0ull) {
            const unsigned slot = 0xFFFFFFFFu - static_cast<unsigned>(keys[j]);
            s = sval[slot];
            id = sid[slot];
        }
        out_s[static_cast<long long>(qi) * k + j] = s;
        out_i[static_cast<long long>(qi) * k + j] = id;
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ x, long long n, int dim, const long long* __restrict__ rows,
                                   int nrows, float* __restrict__ out) {
    const int r = blockIdx.x;
    if (r >= nrows) return;
    const long long row = rows[r];
    for (int d = threadIdx.x; d < dim; d += blockDim.x)
        out[static_cast<long long>(r) * dim + d] = (row >= 0 && row < n) ? x[row * dim + d] : 0.f;
}

// =====================================================================================================
// greedy MMR (langchain_core.vectorstores.utils.maximal_marginal_relevance), fp64 like numpy on
// python-float embeddings.  One CTA per query; fetch_k is small (20 by default).
// =====================================================================================================
__global__ void __launch_bounds__(128) mmr_kernel(const float* __restrict__ q, const float* __restrict__ cand,
                                                  const int* __restrict__ n_cand, int fetch_k, int dim, int k,
                                                  double lambda_mult, int* __restrict__ out_sel) {
    extern __shared__ double sh[];
    double* sim_q = sh;                    // [fetch_k]
    double* norms = sim_q + fetch_k;       // [fetch_k]
    double* simmat = norms + fetch_k;      // [fetch_k * fetch_k]
    double* best_sel = simmat + fetch_k * fetch_k;  // [fetch_k] running max similarity to the selected set
    __shared__ int s_sel[1];
    const int qi = blockIdx.x;
    const int n = n_cand ? min(n_cand[qi], fetch_k) : fetch_k;
    const float* qv = q + static_cast<long long>(qi) * dim;
    const float* E = cand + static_cast<long long>(qi) * fetch_k * dim;
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = lane_id();
    // norms, <q, e_i>, <e_i, e_j>
    double qn = 0.0;
    for (int d = lane; d < dim; d += 32) qn += static_cast<double>(qv[d]) * qv[d];
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    qn = sqrt(qn);
    for (int i = warp; i < n; i += nw) {
        double s = 0.0, nn = 0.0;
        for (int d = lane; d < dim; d += 32) {
            const double e = E[static_cast<long long>(i) * dim + d];
            s += e * qv[d];
            nn += e * e;
        }
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); nn += __shfl_xor_sync(0xffffffffu, nn, o); }
        if (lane == 0) { norms[i] = sqrt(nn); sim_q[i] = s; }
    }
    __syncthreads();
    for (int pq = warp; pq < n * n; pq += nw) {
        const int i = pq / n, j = pq % n;
        double s = 0.0;
        for (int d = lane; d < dim; d += 32) s += static_cast<double>(E[static_cast<long long>(i) * dim + d]) * E[static_cast<long long>(j) * dim + d];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const double den = norms[i] * norms[j];
            double c = den > 0.0 ? s / den : 0.0;
            if (isnan(c) || isinf(c)) c = 0.0;
            simmat[i * fetch_k + j] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) {
            const double den = qn * norms[i];
            double c = den > 0.0 ? sim_q[i] / den : 0.0;
            if (isnan(c) || isinf(c)) c = 0.0;
            sim_q[i] = c;
            best_sel[i] = -1e300;
        }
        int nsel = 0;
        const int kk = min(k, n);
        unsigned long long chosen = 0ull;  // fetch_k <= 64
        if (kk > 0) {
            int first = 0;
            for (int i = 1; i < n; ++i) if (sim_q[i] > sim_q[first]) first = i;   // np.argmax: first maximum
            out_sel[static_cast<long long>(qi) * k + nsel++] = first;
            chosen |= 1ull << first;
            for (int i = 0; i < n; ++i) best_sel[i] = simmat[i * fetch_k + first];
        }
        while (nsel < kk) {
            double best = -INFINITY;
            int add = -1;
            for (int i = 0; i < n; ++i) {
                if (chosen & (1ull << i)) continue;
                const double sc = lambda_mult * sim_q[i] - (1.0 - lambda_mult) * best_sel[i];
                if (sc > best) { best = sc; add = i; }
            }
            if (add < 0) break;
            out_sel[static_cast<long long>(qi) * k + nsel++] = add;
            chosen |= 1ull << add;
            for (int i = 0; i < n; ++i) best_sel[i] = fmax(best_sel[i], simmat[i * fetch_k + add]);
        }
        for (int j = nsel; j < k; ++j) out_sel[static_cast<long long>(qi) * k + j] = -1;
        s_sel[0] = nsel;
    }
}

}  // namespace rmu

// =====================================================================================================
// host side
// =====================================================================================================
using namespace rmu;

struct rmu_index {
    int dim = 0, metric = 0, device = 0, sms = 0;
    int64_t n = 0, cap = 0;
    float* x = nullptr;
    float* rscale = nullptr;
    float* rbias = nullptr;
    unsigned* max_norm_bits = nullptr;   // [0] max ||x||, [1] max | ||x||^2 - 1 |
    float* dev_h = nullptr;              // pinned host copy of [1], refreshed after every insert
    cudaEvent_t stats_ev = nullptr;
    CUtensorMap tmap{};
    int64_t tmap_rows = -1;
    int tmap_bn = 0;
    // scratch
    void* ws = nullptr;
    size_t ws_bytes = 0;
    void* hbuf = nullptr;          // device staging of the *_host entry points (queries, scores, ids)
    size_t hbuf_bytes = 0;
    cudaEvent_t ws_done = nullptr;
    std::mutex mu;
    std::mutex host_mu;            // serialises the *_host entry points (they share hbuf and synchronise anyway)
};

static int ensure_ws(rmu_index* idx, size_t bytes) {
    if (bytes <= idx->ws_bytes) return RMU_OK;
    if (idx->ws) {
        RMU_CUDA(cudaDeviceSynchronize());
        RMU_CUDA(cudaFree(idx->ws));
        idx->ws = nullptr;
        idx->ws_bytes = 0;
    }
    size_t want = bytes + bytes / 4;
    RMU_CUDA(cudaMalloc(&idx->ws, want));
    idx->ws_bytes = want;
    return RMU_OK;
}

struct ScanCfg { int bn, nbuf, nslab, keep; };

static int keep_for_k(int k) {
    // candidates kept per query by the coarse pass: >= 3k where that fits, never above 256
    if (3 * k <= 64) return 64;
    if (3 * k <= 128) return 128;
    return 256;
}

template <int BN, int NBUF, int NSLAB, int KD, bool TMA3D, int KEEP>
static int launch_scan(const CUtensorMap& tmap, const ScanParams& p, int grid, cudaStream_t st) {
    auto kern = scan_tf32_kernel<BN, NBUF, NSLAB, KD, TMA3D, KEEP>;
    const size_t smem = static_cast<size_t>(NSLAB) * BN * 128 * KD + 32 * 128 * sizeof(float) + (2 * NSLAB + 2 * NBUF + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        RMU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr_set = true;
    }
    ProfScope _ps(p.phase == 1 ? PROF_SCAN_LEAD : PROF_SCAN, st);
    kern<<<grid, kScanThreads, smem, st>>>(tmap, p);
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

// scan geometry variants {rows per tile, TMEM accumulators, stages, K-blocks per stage, 3-D map}:
//   0: 64 x 2, 24 x  8 KB (2-D)           1: 128 x 1, 12 x 16 KB (2-D)
//   2: 64 x 2,  6 x 32 KB (3-D, 4 kb/op)  3: 64 x 2, 12 x 16 KB (3-D, 2 kb/op)   4: 64 x 2, 4 x 48 KB (3-D, 6 kb/op)
//   5: 128 x 1, 4 stages x 3 boxes of 16 KB   6: 128 x 1, 3 stages x 4 boxes   7: 128 x 1, 2 stages x 6 boxes
static int scan_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RMU_SCAN_VARIANT");
        v = e ? atoi(e) : 5;
        if (v < 0 || v > 7) v = 5;
    }
    return v;
}
static bool scan_3d() { const int v = scan_variant(); return v >= 2 && v <= 4; }
static int scan_bn() { const int v = scan_variant(); return (v == 1 || v >= 5) ? 128 : 64; }
static int scan_kd() { const int v = scan_variant(); return v == 2 ? 4 : v == 3 ? 2 : v == 4 ? 6 : 1; }   // 3-D box depth

template <int KEEP>
static int dispatch_variant(const CUtensorMap& tmap, const ScanParams& p, int grid, cudaStream_t st) {
    switch (scan_variant()) {
        case 0: return launch_scan<64, 2, 24, 1, false, KEEP>(tmap, p, grid, st);
        case 1: return launch_scan<128, 1, 12, 1, false, KEEP>(tmap, p, grid, st);
        case 2: return launch_scan<64, 2, 6, 4, true, KEEP>(tmap, p, grid, st);
        case 3: return launch_scan<64, 2, 12, 2, true, KEEP>(tmap, p, grid, st);
        case 4: return launch_scan<64, 2, 4, 6, true, KEEP>(tmap, p, grid, st);
        case 6: return launch_scan<128, 1, 3, 4, false, KEEP>(tmap, p, grid, st);
        case 7: return launch_scan<128, 1, 2, 6, false, KEEP>(tmap, p, grid, st);
        default: return launch_scan<128, 1, 4, 3, false, KEEP>(tmap, p, grid, st);
    }
}

static int dispatch_scan(int keep, const CUtensorMap& tmap, const ScanParams& p, int grid, cudaStream_t st) {
    switch (keep) {
        case 32: return dispatch_variant<32>(tmap, p, grid, st);
        case 64: return dispatch_variant<64>(tmap, p, grid, st);
        case 128: return dispatch_variant<128>(tmap, p, grid, st);
        case 256: return dispatch_variant<256>(tmap, p, grid, st);
        default: set_error("scan: unsupported KEEP"); return RMU_ERR_UNSUPPORTED;
    }
}

extern "C" {

int rmu_index_create(int dim, int metric, rmu_index** out) {
    if (!out || dim <= 0 || metric < 0 || metric > 2) { set_error("rmu_index_create: bad argument"); return RMU_ERR_ARG; }
    rmu_index* idx = new rmu_index();
    idx->dim = dim;
    idx->metric = metric;
    if (cudaGetDevice(&idx->device) != cudaSuccess || (idx->sms = device_sm_count()) <= 0) {
        set_error("rmu_index_create: no CUDA device (this library has no CPU path)");
        delete idx;
        return RMU_ERR_CUDA;
    }
    cudaError_t e = cudaMalloc(&idx->max_norm_bits, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(idx->max_norm_bits, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMallocHost(&idx->dev_h, sizeof(float));
    if (e == cudaSuccess) { *idx->dev_h = 0.f; e = cudaEventCreateWithFlags(&idx->stats_ev, cudaEventDisableTiming); }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&idx->ws_done, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error(std::string("rmu_index_create: ") + cudaGetErrorString(e)); delete idx; return RMU_ERR_CUDA; }
    *out = idx;
    return RMU_OK;
}

void rmu_index_destroy(rmu_index* idx) {
    if (!idx) return;
    cudaDeviceSynchronize();
    cudaFree(idx->x); cudaFree(idx->rscale); cudaFree(idx->rbias); cudaFree(idx->max_norm_bits); cudaFree(idx->ws); cudaFree(idx->hbuf);
    if (idx->ws_done) cudaEventDestroy(idx->ws_done);
    if (idx->stats_ev) cudaEventDestroy(idx->stats_ev);
    if (idx->dev_h) cudaFreeHost(idx->dev_h);
    delete idx;
}

}  // extern "C"

// grow the corpus allocation to `rows` rows; idx->mu must be held
static int reserve_locked(rmu_index* idx, int64_t rows) {
    if (rows <= idx->cap) return RMU_OK;
    RMU_CUDA(cudaDeviceSynchronize());
    float* nx = nullptr; float* ns = nullptr; float* nb = nullptr;
    cudaError_t e = cudaMalloc(&nx, static_cast<size_t>(rows) * idx->dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&ns, static_cast<size_t>(rows) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&nb, static_cast<size_t>(rows) * sizeof(float));
    if (e == cudaSuccess && idx->n > 0) {
        e = cudaMemcpy(nx, idx->x, static_cast<size_t>(idx->n) * idx->dim * sizeof(float), cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ns, idx->rscale, static_cast<size_t>(idx->n) * sizeof(float), cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(nb, idx->rbias, static_cast<size_t>(idx->n) * sizeof(float), cudaMemcpyDeviceToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(nx); cudaFree(ns); cudaFree(nb);            // the index keeps its old storage
        set_error(std::string("rmu_index_reserve: ") + cudaGetErrorString(e));
        return RMU_ERR_CUDA;
    }
    cudaFree(idx->x); cudaFree(idx->rscale); cudaFree(idx->rbias);
    idx->x = nx; idx->rscale = ns; idx->rbias = nb;
    idx->cap = rows;
    idx->tmap_rows = -1;
    return RMU_OK;
}

extern "C" {

int rmu_index_reserve(rmu_index* idx, int64_t rows) {
    if (!idx || rows < 0) { set_error("rmu_index_reserve: bad argument"); return RMU_ERR_ARG; }
    std::lock_guard<std::mutex> g(idx->mu);
    return reserve_locked(idx, rows);
}

int rmu_index_add(rmu_index* idx, const float* vecs, int64_t n, int src_is_host, void* stream) {
    if (!idx || (n > 0 && !vecs) || n < 0) { set_error("rmu_index_add: bad argument"); return RMU_ERR_ARG; }
    if (n == 0) return RMU_OK;
    std::lock_guard<std::mutex> g(idx->mu);      // size check, growth and append are one critical section
    if (idx->n + n > static_cast<int64_t>(0xFFFFFFF0u)) { set_error("rmu_index_add: more than 2^32 rows per shard"); return RMU_ERR_UNSUPPORTED; }
    if (idx->n + n > idx->cap) {
        int64_t want = std::max<int64_t>(idx->n + n, idx->cap + idx->cap / 2);
        int rc = reserve_locked(idx, std::max<int64_t>(want, 1024));
        if (rc != RMU_OK) return rc;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RMU_CUDA(cudaMemcpyAsync(idx->x + idx->n * idx->dim, vecs, static_cast<size_t>(n) * idx->dim * sizeof(float),
                             src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    const int wpb = 8;
    const long long blocks = (n + wpb - 1) / wpb;
    row_stats_kernel<<<static_cast<unsigned>(blocks), wpb * 32, 0, st>>>(idx->x, idx->n, n, idx->dim, idx->metric,
                                                                          idx->rscale, idx->rbias, idx->max_norm_bits,
                                                                          idx->max_norm_bits + 1);
    count_launch();
    RMU_CHECK_LAUNCH();
    RMU_CUDA(cudaMemcpyAsync(idx->dev_h, idx->max_norm_bits + 1, sizeof(float), cudaMemcpyDeviceToHost, st));
    RMU_CUDA(cudaEventRecord(idx->stats_ev, st));
    idx->n += n;
    idx->tmap_rows = -1;
    return RMU_OK;
}

int64_t rmu_index_size(const rmu_index* idx) { return idx ? idx->n : -1; }
int rmu_index_dim(const rmu_index* idx) { return idx ? idx->dim : -1; }
int rmu_index_metric(const rmu_index* idx) { return idx ? idx->metric : -1; }
const float* rmu_index_data(const rmu_index* idx) { return idx ? idx->x : nullptr; }

int rmu_index_clear(rmu_index* idx) {
    if (!idx) return RMU_ERR_ARG;
    std::lock_guard<std::mutex> g(idx->mu);
    RMU_CUDA(cudaDeviceSynchronize());
    idx->n = 0;
    idx->tmap_rows = -1;
    RMU_CUDA(cudaMemset(idx->max_norm_bits, 0, 2 * sizeof(unsigned)));
    *idx->dev_h = 0.f;
    return RMU_OK;
}

int rmu_index_search(rmu_index* idx, const float* queries, int nq, int k, int64_t id_offset, int mode,
                     float* out_scores, int64_t* out_ids, int32_t* stats_h, void* stream) {
    if (!idx || nq < 0 || k <= 0 || (nq > 0 && (!queries || !out_scores || !out_ids))) {
        set_error("rmu_index_search: bad argument");
        return RMU_ERR_ARG;
    }
    if (k > 1024) { set_error("rmu_index_search: k > 1024 is not supported"); return RMU_ERR_UNSUPPORTED; }
    if (stats_h) { stats_h[0] = stats_h[1] = stats_h[2] = stats_h[3] = 0; }
    if (nq == 0) return RMU_OK;
    std::lock_guard<std::mutex> g(idx->mu);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RMU_CUDA(cudaStreamWaitEvent(st, idx->ws_done, 0));   // scratch is shared by all callers of this handle

    const int D = idx->dim;
    const long long N = idx->n;
    // tensor scan eligibility: TMA row pitch multiple of 16 B, query block fits TMEM, enough rows, k small
    // (queries must be 16-byte aligned: the scan loads them as float4)
    const bool tensor_ok = (D % 4 == 0) && D <= 2 * kScanACols && N >= 16384 && k <= 128 && mode != RMU_SEARCH_EXACT &&
                           (reinterpret_cast<uintptr_t>(queries) & 15) == 0 &&
                           (D <= kScanACols || !scan_3d());   // the K-split passes use the 2-D tensor map
    const bool ksplit = tensor_ok && D > kScanACols;
    const int keep = keep_for_k(k);                       // tensor: candidates kept per query (>= 3k)
    int keepx = 32; while (keepx < k) keepx <<= 1;        // exact: per-chunk list length (>= k)
    const int nchunks = static_cast<int>((N + kChunk - 1) / kChunk);

    // ---- scratch layout
    const int grid_scan = idx->sms;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_flags = carve(sizeof(int) * nq);
    const size_t o_qmap = carve(sizeof(int) * nq);
    const size_t o_nsel = carve(sizeof(int) * 4);
    const size_t o_scan = tensor_ok ? carve(sizeof(unsigned long long) * grid_scan * kScanQ * 2 * keep) : 0;
    const size_t o_tau = carve(sizeof(float) * nq);
    const size_t o_cnt = tensor_ok ? carve(sizeof(int) * grid_scan * kScanQ) : 0;
    const long long npad = (N + 31) / 32 * 32;
    const size_t o_part = ksplit ? carve(sizeof(float) * kScanQ * static_cast<size_t>(npad)) : 0;
    const size_t o_exact = carve(sizeof(unsigned long long) * std::max(nchunks, 1) * static_cast<size_t>(nq) * keepx);
    int rc = ensure_ws(idx, off);
    if (rc != RMU_OK) return rc;
    uint8_t* ws = static_cast<uint8_t*>(idx->ws);
    int* d_flags = reinterpret_cast<int*>(ws + o_flags);
    int* d_qmap = reinterpret_cast<int*>(ws + o_qmap);
    int* d_nsel = reinterpret_cast<int*>(ws + o_nsel);
    unsigned long long* d_scan = reinterpret_cast<unsigned long long*>(ws + o_scan);
    unsigned long long* d_exact = reinterpret_cast<unsigned long long*>(ws + o_exact);
    float* d_tau0 = reinterpret_cast<float*>(ws + o_tau);
    int* d_cnt = reinterpret_cast<int*>(ws + o_cnt);
    float* d_part = reinterpret_cast<float*>(ws + o_part);

    const size_t qsmem = static_cast<size_t>(D) * sizeof(float);
    int scan_launches = 0;

    if (N == 0) {
        // nothing to search: all results missing.  Reuse finalize with zero lists.
        FinalizeParams fp{};
        fp.lists = d_exact; fp.nlists = 0; fp.qstride = nq; fp.lstride = keepx; fp.len = keepx; fp.ksel = keepx;
        fp.x = idx->x; fp.n = 0; fp.dim = D; fp.metric = idx->metric; fp.q = queries; fp.q0 = 0; fp.exact = 1;
        fp.k = k; fp.id_offset = id_offset; fp.max_norm_bits = idx->max_norm_bits; fp.eps_rel = 0.f;
        fp.out_scores = out_scores; fp.out_ids = reinterpret_cast<long long*>(out_ids); fp.flags = nullptr;
        finalize_kernel<<<nq, kSelThreads, qsmem, st>>>(fp);
        count_launch();
        RMU_CHECK_LAUNCH();
        RMU_CUDA(cudaEventRecord(idx->ws_done, st));
        return RMU_OK;
    }

    // unit-norm corpora (sentence-transformers' Normalize output): cosine and L2 rank exactly like the inner
    // product, so the scan skips the per-row scale / bias and the certificate absorbs the 1e-6 slack
    bool unit_rows = false;
    if (tensor_ok && idx->metric != RMU_METRIC_IP) {
        RMU_CUDA(cudaEventSynchronize(idx->stats_ev));
        unit_rows = *idx->dev_h < 1e-6f;
    }
    if (tensor_ok) {
        if (idx->tmap_rows != N || idx->tmap_bn != scan_bn()) {
            if (scan_3d()) {
                if (D % 32 != 0) { set_error("scan variant needs dim % 32 == 0"); return RMU_ERR_UNSUPPORTED; }
                rc = make_tmap_rows_kblocks(&idx->tmap, idx->x, static_cast<uint64_t>(N), D / 32, scan_bn(), scan_kd());
            } else {
                rc = make_tmap_2d(&idx->tmap, idx->x, static_cast<uint64_t>(N), static_cast<uint64_t>(D),
                                  static_cast<uint64_t>(D) * sizeof(float), 32, scan_bn(), 4);
            }
            if (rc != RMU_OK) return rc;
            idx->tmap_rows = N;
            idx->tmap_bn = scan_bn();
        }
        const int ntiles = static_cast<int>((N + scan_bn() - 1) / scan_bn());
        // one logical scan = one launch, or two when dim > 384 (K split: park partial scores, then finish)
        auto scan_launch = [&](int keep_x, ScanParams sp, int grid) -> int {
            sp.partial = d_part; sp.npad = npad;
            if (!ksplit) {
                sp.kcol0 = 0; sp.kdim = D; sp.kpass = 0;
                ++scan_launches;
                return dispatch_scan(keep_x, idx->tmap, sp, grid, st);
            }
            sp.kcol0 = 0; sp.kdim = kScanACols; sp.kpass = 1;
            int r = dispatch_scan(keep_x, idx->tmap, sp, grid, st);
            if (r != RMU_OK) return r;
            sp.kcol0 = kScanACols; sp.kdim = D - kScanACols; sp.kpass = 2;
            scan_launches += 2;
            return dispatch_scan(keep_x, idx->tmap, sp, grid, st);
        };
        for (int q0 = 0; q0 < nq; q0 += kScanQ) {
            ScanParams sp{};
            sp.q = queries; sp.q0 = q0; sp.nq = std::min(kScanQ, nq - q0); sp.dim = D; sp.n = N; sp.ntiles = ntiles;
            sp.rscale = (idx->metric == RMU_METRIC_COSINE && !unit_rows) ? idx->rscale : nullptr;
            sp.rbias = (idx->metric == RMU_METRIC_L2 && !unit_rows) ? idx->rbias : nullptr;
            sp.lists = d_scan; sp.counts = d_cnt;
            { static const char* ab = getenv("RMU_SCAN_ABLATE"); sp.ablate = ab ? atoi(ab) : 0; }
            const int grid = std::min(grid_scan, ntiles);
            // threshold exchange (big corpora): a cheap lead pass (first ~1 % of every CTA's tiles, KEEP = 32)
            // estimates per-query thresholds, then the full pass starts from them, so its epilogue almost
            // never takes the insert path.  The lead rows are read twice (+2 % traffic).
            static const int lead_env = [] { const char* e = getenv("RMU_SCAN_LEAD_PCT"); return e ? atoi(e) : -1; }();
            const int lead_pct = lead_env >= 0 ? lead_env : 1;
            const int tiles_per_cta = ntiles / grid;
            constexpr int kLeadKeep = 32;
            const bool exchange = lead_pct > 0 && tiles_per_cta >= 16 && grid * kLeadKeep >= keep;
            sp.tau0 = d_tau0;
            if (exchange) {
                sp.phase = 1;
                sp.lead = std::max(2, (tiles_per_cta * lead_pct + 99) / 100);
                rc = scan_launch(kLeadKeep, sp, grid);
                if (rc != RMU_OK) return rc;
                { ProfScope _ps(PROF_FINALIZE, st);
                select_tau_kernel<<<sp.nq, kSelThreads, 0, st>>>(d_scan, d_cnt, grid, scan_cap(kLeadKeep), scan_cap(kLeadKeep), keep, q0, d_tau0); }
                count_launch();
                RMU_CHECK_LAUNCH();
                sp.phase = 2;
                rc = scan_launch(keep, sp, grid);
                if (rc != RMU_OK) return rc;
            } else {
                sp.phase = 0;
                rc = scan_launch(keep, sp, grid);
                if (rc != RMU_OK) return rc;
            }
            FinalizeParams fp{};
            fp.lists = d_scan; fp.nlists = grid; fp.qstride = kScanQ; fp.lstride = scan_cap(keep); fp.ksel = keep;
            if (keep >= 128) { fp.counts = nullptr; fp.len = keep; }            // sorted, cut to keep, zero padded
            else { fp.counts = d_cnt; fp.len = scan_cap(keep); }                // raw lists + counts
            fp.x = idx->x; fp.n = N; fp.dim = D; fp.metric = idx->metric; fp.q = queries; fp.q0 = q0; fp.exact = 0;
            fp.unit_rows = unit_rows ? 1 : 0;
            fp.k = k; fp.id_offset = id_offset; fp.max_norm_bits = idx->max_norm_bits;
            fp.eps_rel = 2.2e-3f;   // > 2^-9: both TF32 operands truncated to 10 mantissa bits
            fp.out_scores = out_scores; fp.out_ids = reinterpret_cast<long long*>(out_ids); fp.flags = d_flags;
            { ProfScope _ps(PROF_FINALIZE, st);
            finalize_kernel<<<sp.nq, kSelThreads, qsmem, st>>>(fp); }
            count_launch();
            RMU_CHECK_LAUNCH();
        }
        compact_flags_kernel<<<1, 32, 0, st>>>(d_flags, nq, d_qmap, d_nsel);
        count_launch();
        RMU_CHECK_LAUNCH();
    }

    if (!tensor_ok || mode == RMU_SEARCH_AUTO) {
        ExactParams ep{};
        ep.x = idx->x; ep.n = N; ep.dim = D; ep.metric = idx->metric; ep.q = queries;
        ep.qmap = tensor_ok ? d_qmap : nullptr; ep.nsel = tensor_ok ? d_nsel : nullptr; ep.nq_total = nq;
        ep.lists = d_exact; ep.keep = keepx;
        const int ngroups = (nq + kExactQT - 1) / kExactQT;
        int gy = tensor_ok ? 1 : std::min(ngroups, std::max(1, (2 * idx->sms + nchunks - 1) / nchunks));
        gy = std::min(gy, 65535);
        dim3 eg(static_cast<unsigned>(nchunks), static_cast<unsigned>(gy));
        const size_t esmem = sizeof(float) * kExactQT * (static_cast<size_t>(D) + kChunk);
        if (esmem > 200 * 1024) { set_error("rmu_index_search: dim too large for the exact scan"); return RMU_ERR_UNSUPPORTED; }
        static size_t esmem_set = 0;
        if (esmem > esmem_set) {
            RMU_CUDA(cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(esmem)));
            esmem_set = esmem;
        }
        { ProfScope _ps(PROF_EXACT, st);
        exact_scan_kernel<<<eg, 256, esmem, st>>>(ep); }
        count_launch();
        RMU_CHECK_LAUNCH();
        FinalizeParams fp{};
        fp.lists = d_exact; fp.nlists = nchunks; fp.qstride = nq; fp.lstride = keepx; fp.len = keepx; fp.ksel = keepx;
        fp.x = idx->x; fp.n = N; fp.dim = D; fp.metric = idx->metric; fp.q = queries; fp.q0 = 0;
        fp.qmap = ep.qmap; fp.nsel = ep.nsel; fp.exact = 1;
        fp.k = k; fp.id_offset = id_offset; fp.max_norm_bits = idx->max_norm_bits; fp.eps_rel = 0.f;
        fp.out_scores = out_scores; fp.out_ids = reinterpret_cast<long long*>(out_ids); fp.flags = nullptr;
        { ProfScope _ps(PROF_EXACT, st);
        finalize_kernel<<<nq, kSelThreads, qsmem, st>>>(fp); }
        count_launch();
        RMU_CHECK_LAUNCH();
    }

    if (stats_h) {
        int nsel = tensor_ok ? 0 : nq;
        if (tensor_ok) {
            RMU_CUDA(cudaMemcpyAsync(&nsel, d_nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
            RMU_CUDA(cudaStreamSynchronize(st));
        }
        stats_h[0] = nsel;
        stats_h[1] = scan_launches;
    }
    RMU_CUDA(cudaEventRecord(idx->ws_done, st));
    return RMU_OK;
}

// diagnostics (tests only): raw TF32 accumulators of the first 64-row tile, out [128, 64] device fp32
int rmu_debug_scan_tile(rmu_index* idx, const float* queries, int nq, float* out, void* stream) {
    if (!idx || !queries || !out || nq <= 0 || nq > kScanQ || idx->n <= 0 || idx->dim % 4 != 0 || idx->dim > kScanACols) {
        set_error("rmu_debug_scan_tile: bad argument");
        return RMU_ERR_ARG;
    }
    std::lock_guard<std::mutex> g(idx->mu);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t dbg_lists = sizeof(unsigned long long) * kScanQ * 2 * 64;
    int rc = ensure_ws(idx, dbg_lists + sizeof(int) * kScanQ + 1024);
    if (rc != RMU_OK) return rc;
    if (scan_3d()) rc = make_tmap_rows_kblocks(&idx->tmap, idx->x, static_cast<uint64_t>(idx->n), idx->dim / 32, scan_bn(), scan_kd());
    else rc = make_tmap_2d(&idx->tmap, idx->x, static_cast<uint64_t>(idx->n), static_cast<uint64_t>(idx->dim),
                           static_cast<uint64_t>(idx->dim) * sizeof(float), 32, scan_bn(), 4);
    if (rc != RMU_OK) return rc;
    idx->tmap_rows = idx->n;
    idx->tmap_bn = scan_bn();
    ScanParams sp{};
    sp.q = queries; sp.q0 = 0; sp.nq = nq; sp.dim = idx->dim; sp.n = idx->n; sp.ntiles = 1;
    sp.lists = static_cast<unsigned long long*>(idx->ws);
    sp.counts = reinterpret_cast<int*>(static_cast<uint8_t*>(idx->ws) + dbg_lists);
    sp.dbg = out; sp.kcol0 = 0; sp.kdim = idx->dim; sp.kpass = 0;
    return dispatch_scan(64, idx->tmap, sp, 1, st);
}

int rmu_index_search_host(rmu_index* idx, const float* queries_h, int nq, int k, int64_t id_offset, int mode,
                          float* out_scores_h, int64_t* out_ids_h, void* stream) {
    if (!idx || nq <= 0 || k <= 0 || !queries_h || !out_scores_h || !out_ids_h) {
        set_error("rmu_index_search_host: bad argument");
        return RMU_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::lock_guard<std::mutex> hg(idx->host_mu);
    const size_t qb = (static_cast<size_t>(nq) * idx->dim * sizeof(float) + 255) & ~size_t(255);
    const size_t sb = (static_cast<size_t>(nq) * k * sizeof(float) + 255) & ~size_t(255);
    const size_t ib = static_cast<size_t>(nq) * k * sizeof(int64_t);
    {
        // handle-owned staging (a cudaMallocAsync/free pair per call costs milliseconds once the pool is trimmed)
        std::lock_guard<std::mutex> g(idx->mu);
        if (qb + sb + ib > idx->hbuf_bytes) {
            RMU_CUDA(cudaStreamSynchronize(st));
            if (idx->hbuf) RMU_CUDA(cudaFree(idx->hbuf));
            idx->hbuf = nullptr;
            idx->hbuf_bytes = 0;
            RMU_CUDA(cudaMalloc(&idx->hbuf, 2 * (qb + sb + ib)));
            idx->hbuf_bytes = 2 * (qb + sb + ib);
        }
    }
    uint8_t* hb = static_cast<uint8_t*>(idx->hbuf);
    float* dq = reinterpret_cast<float*>(hb);
    float* ds = reinterpret_cast<float*>(hb + qb);
    int64_t* di = reinterpret_cast<int64_t*>(hb + qb + sb);
    RMU_CUDA(cudaMemcpyAsync(dq, queries_h, static_cast<size_t>(nq) * idx->dim * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = rmu_index_search(idx, dq, nq, k, id_offset, mode, ds, di, nullptr, st);
    if (rc == RMU_OK) {
        RMU_CUDA(cudaMemcpyAsync(out_scores_h, ds, static_cast<size_t>(nq) * k * sizeof(float), cudaMemcpyDeviceToHost, st));
        RMU_CUDA(cudaMemcpyAsync(out_ids_h, di, ib, cudaMemcpyDeviceToHost, st));
    }
    RMU_CUDA(cudaStreamSynchronize(st));
    return rc;
}

int rmu_index_gather(rmu_index* idx, const int64_t* rows, int n, float* out, void* stream) {
    if (!idx || n < 0 || (n > 0 && (!rows || !out))) { set_error("rmu_index_gather: bad argument"); return RMU_ERR_ARG; }
    if (n == 0) return RMU_OK;
    gather_rows_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(idx->x, idx->n, idx->dim,
                                                                          reinterpret_cast<const long long*>(rows), n, out);
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

int rmu_topk_merge(const float* scores, const int64_t* ids, int R, int nq, int k, int metric, float* out_scores,
                   int64_t* out_ids, void* stream) {
    if (R <= 0 || nq < 0 || k <= 0 || !scores || !ids || !out_scores || !out_ids || metric < 0 || metric > 2) {
        set_error("rmu_topk_merge: bad argument");
        return RMU_ERR_ARG;
    }
    if (nq == 0) return RMU_OK;
    const int n = R * k;
    if (n > 4096) { set_error("rmu_topk_merge: R*k > 4096"); return RMU_ERR_UNSUPPORTED; }
    int n2 = 32; while (n2 < n) n2 <<= 1;
    const size_t smem = static_cast<size_t>(n2) * 8 + static_cast<size_t>(n) * 8 + static_cast<size_t>(n) * 4;
    static bool attr_set = false;
    if (!attr_set) {
        RMU_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    ProfScope _ps(PROF_MERGE, static_cast<cudaStream_t>(stream));
    merge_kernel<<<nq, 256, smem, static_cast<cudaStream_t>(stream)>>>(scores, reinterpret_cast<const long long*>(ids), R, nq, k,
                                                                       metric, out_scores, reinterpret_cast<long long*>(out_ids));
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

int rmu_mmr_select(const float* q, const float* cand, const int32_t* n_cand, int nq, int fetch_k, int dim, int k,
                   float lambda_mult, int32_t* out_sel, void* stream) {
    if (!q || !cand || !out_sel || nq < 0 || fetch_k <= 0 || fetch_k > 64 || dim <= 0 || k <= 0) {
        set_error("rmu_mmr_select: bad argument (fetch_k must be in 1..64)");
        return RMU_ERR_ARG;
    }
    if (nq == 0) return RMU_OK;
    const size_t smem = sizeof(double) * (3 * static_cast<size_t>(fetch_k) + static_cast<size_t>(fetch_k) * fetch_k);
    mmr_kernel<<<nq, 128, smem, static_cast<cudaStream_t>(stream)>>>(q, cand, n_cand, fetch_k, dim, k,
                                                                     static_cast<double>(lambda_mult), out_sel);
    count_launch();
    RMU_CHECK_LAUNCH();
    return RMU_OK;
}

}  // extern "C"
