// Tensor-core engine of the fused distance + top-KC candidate search: tcgen05.mma (kind::tf32)
// with accumulators in TMEM, operands staged by 1-D TMA bulk copies, selection fused into the
// TMEM epilogue so the score matrix never leaves the SM.
//
// Same role as search_simt.cu (it replaces the dgemm + heap-test hot loop of scikit-learn's
// EuclideanArgKmin64, $SP/sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:401-510)
// but the contraction runs on the 5th-generation tensor cores.  TF32 scores carry ~1e-3
// relative error, so this kernel is only ever a FILTER: it returns the KC best references by
// approximate score plus the KC-th score, refine.cu re-evaluates the survivors in float64 and
// proves (error bound eps_s = 2^-10) that nothing outside the list can belong to the k nearest;
// rows it cannot certify are re-searched by the FP32 SIMT engine and, failing that, by the
// exhaustive float64 kernel.
//
// One CTA = MT x 128 queries (MT = 3, or 2 when shared memory is short) = MT M=128 MMA tiles that
// share every reference tile (N=64):
//   last warp, lane 0  driver: TMA bulk copies (cp.async.bulk + mbarrier) of the query image once and
//                  of reference tiles through an NSTAGE ring, then K/8 x MT tcgen05.mma per tile,
//                  tcgen05.commit onto the "smem slot free" and "accumulator ready" mbarriers;
//   warps 0..4MT-1 epilogue: thread <-> query (TMEM lane), tcgen05.ld 32 columns at a time, min3
//                  tree against the query's threshold, survivors appended to a private shared-
//                  memory buffer that the warp compacts cooperatively (bitonic sort) when it
//                  fills; two TMEM accumulator stages let the MMAs of tile t+1 overlap the
//                  epilogue of tile t.
// The |r|^2 term is folded into the contraction: each operand gets one extra K block holding
// (1,1,1,0,..) on the query side and a 3-way TF32 split of |r|^2 on the reference side, so the
// accumulator is directly s = |r|^2 - 2 q.r.
//
// Operand layout (no swizzle, K-major "interleaved" canonical layout): 16-byte K chunks of 4
// TF32 values, [chunk][row][4]; core matrix = 8 rows x 16 B contiguous, SBO = 128 B between row
// groups, LBO = rows * 16 B between K chunks.  The images are pre-arranged in HBM in exactly
// this order, so a plain bulk copy stages them.
#include "common.cuh"
#include "kernels.h"

namespace sk {

constexpr int TC_CAP = 32;             // private buffer slots per query (KC kept + lazy appends)
constexpr uint32_t TC_ROWB = 16;       // bytes of one row of one K chunk (4 TF32)

// ---- tcgen05 wrappers -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                 uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread (asynchronous: the
// registers are valid after tmem_ld_wait)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2,
// B=TF32 [10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_M >> 4) << 24);

template <int MT> struct TcShape {
    static constexpr int kQueries = MT * TC_M;          // queries per CTA
    static constexpr int kEpiWarps = MT * 4;
    static constexpr int kThreads = (kEpiWarps + 1) * 32;
    static constexpr int kBufLd = kQueries + 1;          // odd stride: conflict-free compaction reads
};

// warp-cooperative compaction of the private buffers of the lanes in `need`
template <int KC, int LD>
__device__ __forceinline__ void tc_compact(unsigned need, float *buf_s, int *buf_i, int warp,
                                           int lane, float &thr, int &cnt) {
    __syncwarp();
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        const int cL = __shfl_sync(SK_FULL, cnt, L);
        const int T = warp * 32 + L;
        float s = SK_INF_F;
        int id = 0x7fffffff;
        if (lane < cL) {
            s = buf_s[lane * LD + T];
            id = buf_i[lane * LD + T];
        }
        warp_sort_pairs<float>(s, id, lane);
        if (lane < KC) {
            buf_s[lane * LD + T] = s;
            buf_i[lane * LD + T] = id;
        }
        const float nthr = __shfl_sync(SK_FULL, s, KC - 1);
        if (lane == L) {
            thr = nthr;
            cnt = min(cL, KC);
        }
    }
    __syncwarp();
}

// append the values of one 8-column sub-group that beat the threshold
template <int LD>
__device__ __forceinline__ void tc_append8(const float *w, float thr, int idb, float *buf_s, int *buf_i,
                                           int tid, int &cnt) {
    float *ps = buf_s + cnt * LD + tid;
    int *pi = buf_i + cnt * LD + tid;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (w[j] < thr) {
            *ps = w[j];
            *pi = idb + j;
            ps += LD;
            pi += LD;
            ++cnt;
        }
    }
}

// selection over 32 accumulator columns of this thread's query.  Hits are rare (about one per
// warp per call), so every level of the test is a warp-uniform vote + branch: the common path
// is the min3 tree and one vote.
template <int KC, int LD>
__device__ __forceinline__ void tc_select32(const uint32_t (&r)[32], int idb, float *buf_s, int *buf_i,
                                            int tid, int warp, int lane, float &thr, int &cnt) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    float ms[4];
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
        const float *w = v + sub * 8;
        ms[sub] = fminf(fminf(fminf(w[0], w[1]), fminf(w[2], w[3])),
                        fminf(fminf(w[4], w[5]), fminf(w[6], w[7])));
    }
    const float m = fminf(fminf(ms[0], ms[1]), fminf(ms[2], ms[3]));
    if (!__any_sync(SK_FULL, m < thr)) return;
    unsigned pend = 0;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
        const bool hit = ms[sub] < thr;
        if (__any_sync(SK_FULL, hit)) {
            if (hit) {
                if (cnt > TC_CAP - 8)
                    pend |= 1u << sub;   // no room for 8 more: compact first (warp-cooperative)
                else
                    tc_append8<LD>(v + sub * 8, thr, idb + sub * 8, buf_s, buf_i, tid, cnt);
            }
        }
    }
    // rare: some lane ran out of buffer slots -> compact those lanes, then they retry
    unsigned need = __ballot_sync(SK_FULL, pend != 0);
    while (need) {
        tc_compact<KC, LD>(need, buf_s, buf_i, warp, lane, thr, cnt);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
            const bool retry = (pend >> sub) & 1u;
            if (__any_sync(SK_FULL, retry)) {
                if (retry) {
                    if (!(ms[sub] < thr)) {
                        pend &= ~(1u << sub);
                    } else if (cnt <= TC_CAP - 8) {
                        tc_append8<LD>(v + sub * 8, thr, idb + sub * 8, buf_s, buf_i, tid, cnt);
                        pend &= ~(1u << sub);
                    }
                }
            }
        }
        need = __ballot_sync(SK_FULL, pend != 0);
    }
}

template <int KC, int MT>
__global__ void __launch_bounds__(TcShape<MT>::kThreads, 1)
search_tc_kernel(const float *__restrict__ qimg, const float *__restrict__ rimg, int kc_tot,
                 int n_rtiles, int nstage, long long n_q, int *__restrict__ cand_idx,
                 float *__restrict__ cand_thr) {
    using Shape = TcShape<MT>;
    constexpr int LD = Shape::kBufLd;
    constexpr int EPI_WARPS = Shape::kEpiWarps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t a_bytes = (uint32_t)kc_tot * TC_M * TC_ROWB;  // one 128-query operand image
    const uint32_t b_bytes = (uint32_t)kc_tot * TC_N * TC_ROWB;  // one 64-plot operand image
    unsigned char *Qs = smem_raw;                                // MT operand images
    unsigned char *Rs = Qs + MT * a_bytes;                       // nstage operand images
    float *buf_s = reinterpret_cast<float *>(Rs + (size_t)nstage * b_bytes);
    int *buf_i = reinterpret_cast<int *>(buf_s + TC_CAP * LD);
    // (2 * TC_CAP * LD * 4 bytes is a multiple of 8, so the barriers stay 8-byte aligned)
    uint64_t *full = reinterpret_cast<uint64_t *>(buf_i + TC_CAP * LD);
    uint64_t *empty = full + nstage;
    uint64_t *tfull = empty + nstage;   // [2] accumulator stage ready
    uint64_t *tempty = tfull + 2;       // [2] accumulator stage drained
    uint64_t *qbar = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EPI_WARPS);
        }
        mbar_init(qbar, 1);
        fence_mbar_init();
    }
    if (warp == EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long qtile = blockIdx.x;
    const int ksteps = kc_tot >> 1;  // MMA K = 8 TF32 = two 16-byte chunks

    if (warp == EPI_WARPS) {
        // ======================= driver: TMA producer + MMA issuer =======================
        if (lane == 0) {
            auto issue_tile = [&](int tn) {
                const int sn = tn % nstage;
                if (tn >= nstage) mbar_wait(&empty[sn], ((tn / nstage) - 1) & 1);
                mbar_expect_tx(&full[sn], b_bytes);
                bulk_g2s(Rs + (size_t)sn * b_bytes, (const unsigned char *)rimg + (size_t)tn * b_bytes,
                         b_bytes, &full[sn]);
            };
            mbar_expect_tx(qbar, MT * a_bytes);
            bulk_g2s(Qs, (const unsigned char *)qimg + (size_t)qtile * MT * a_bytes, MT * a_bytes, qbar);
            for (int tn = 0; tn < nstage - 1 && tn < n_rtiles; ++tn) issue_tile(tn);
            mbar_wait(qbar, 0);
            const uint32_t q_addr = smem_u32(Qs), r_addr = smem_u32(Rs);
            const uint32_t a_lbo = TC_M * TC_ROWB, b_lbo = TC_N * TC_ROWB;
            for (int t = 0; t < n_rtiles; ++t) {
                const int s = t % nstage, a = t & 1;
                mbar_wait(&full[s], (t / nstage) & 1);
                if (t >= 2) mbar_wait(&tempty[a], ((t >> 1) - 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < MT; ++h) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)((a * MT + h) * TC_N);
                    for (int j = 0; j < ksteps; ++j) {
                        const uint64_t adesc = tc_smem_desc(q_addr + h * a_bytes + j * 2 * a_lbo, a_lbo, 128);
                        const uint64_t bdesc = tc_smem_desc(r_addr + s * b_bytes + j * 2 * b_lbo, b_lbo, 128);
                        tc_mma_tf32(d_tmem, adesc, bdesc, TC_IDESC, j > 0 ? 1u : 0u);
                    }
                }
                tc_commit(&empty[s]);   // smem slot reusable once these MMAs have read it
                tc_commit(&tfull[a]);   // accumulators of tile t complete
                if (t + nstage - 1 < n_rtiles) issue_tile(t + nstage - 1);
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue: thread <-> query =======================
        const int tid = threadIdx.x;          // query slot; TMEM lane (tid & 127) of M tile (tid >> 7)
        const int h = warp >> 2;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float thr = SK_INF_F;
        int cnt = 0;
        static_assert(TC_N == 64, "epilogue assumes two 32-column groups per tile");
        for (int t = 0; t < n_rtiles; ++t) {
            const int a = t & 1;
            mbar_wait(&tfull[a], (t >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_base + (uint32_t)((a * MT + h) * TC_N);
            uint32_t r0[32], r1[32];
            tmem_ld32_issue(taddr, r0);
            tmem_ld32_issue(taddr + 32, r1);
            tmem_ld_wait();
            // all TMEM reads of this stage are complete: release it before selecting
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
            tc_select32<KC, LD>(r0, t * TC_N, buf_s, buf_i, tid, warp, lane, thr, cnt);
            tc_select32<KC, LD>(r1, t * TC_N + 32, buf_s, buf_i, tid, warp, lane, thr, cnt);
        }
        // final compaction of every lane, then write the candidates
        __syncwarp();
        for (int L = 0; L < 32; ++L) {
            const int cL = __shfl_sync(SK_FULL, cnt, L);
            const int T = warp * 32 + L;
            float s = SK_INF_F;
            int id = 0x7fffffff;
            if (lane < cL) {
                s = buf_s[lane * LD + T];
                id = buf_i[lane * LD + T];
            }
            warp_sort_pairs<float>(s, id, lane);
            const long long q = qtile * Shape::kQueries + T;
            if (q < n_q && lane < KC) {
                cand_idx[q * KC + lane] = (id == 0x7fffffff) ? -1 : id;
                if (lane == KC - 1) cand_thr[q] = s;  // +inf when fewer than KC candidates exist
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

size_t search_tc_smem_bytes(int kc_tot, int nstage, int mt) {
    const size_t a = (size_t)kc_tot * TC_M * TC_ROWB, b = (size_t)kc_tot * TC_N * TC_ROWB;
    const size_t ld = (size_t)mt * TC_M + 1;
    return mt * a + nstage * b + 2 * TC_CAP * ld * 4 + (size_t)(2 * nstage + 5) * 8 + 16;
}

// (M tiles per CTA, ring stages) that fit the 227 KB of shared memory; mt = 0 if nothing fits
void search_tc_pick_shape(int kc_tot, int *mt, int *nstage) {
    for (int m = 3; m >= 2; --m)
        for (int s = 4; s >= 3; --s)
            if (search_tc_smem_bytes(kc_tot, s, m) <= 227 * 1024) {
                *mt = m;
                *nstage = s;
                return;
            }
    for (int m = 3; m >= 2; --m)
        if (search_tc_smem_bytes(kc_tot, 2, m) <= 227 * 1024) {
            *mt = m;
            *nstage = 2;
            return;
        }
    *mt = 0;
    *nstage = 0;
}

template <int KC, int MT>
static cudaError_t launch_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles, int nstage,
                             long long n_q, int *cand_idx, float *cand_thr, cudaStream_t st) {
    const size_t smem = search_tc_smem_bytes(kc_tot, nstage, MT);
    cudaError_t e = cudaFuncSetAttribute(search_tc_kernel<KC, MT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const long long per = TcShape<MT>::kQueries;
    const long long n_qtiles = (n_q + per - 1) / per;
    search_tc_kernel<KC, MT><<<(unsigned)n_qtiles, TcShape<MT>::kThreads, smem, st>>>(
        qimg, rimg, kc_tot, n_rtiles, nstage, n_q, cand_idx, cand_thr);
    return cudaGetLastError();
}

cudaError_t launch_search_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles,
                             long long n_q, int kc, int mt, int nstage, int *cand_idx,
                             float *cand_thr, cudaStream_t st) {
    if (n_q <= 0) return cudaSuccess;
    if (kc == 8 && mt == 3) return launch_tc<8, 3>(qimg, rimg, kc_tot, n_rtiles, nstage, n_q, cand_idx, cand_thr, st);
    if (kc == 8 && mt == 2) return launch_tc<8, 2>(qimg, rimg, kc_tot, n_rtiles, nstage, n_q, cand_idx, cand_thr, st);
    if (kc == 16 && mt == 3) return launch_tc<16, 3>(qimg, rimg, kc_tot, n_rtiles, nstage, n_q, cand_idx, cand_thr, st);
    if (kc == 16 && mt == 2) return launch_tc<16, 2>(qimg, rimg, kc_tot, n_rtiles, nstage, n_q, cand_idx, cand_thr, st);
    return cudaErrorInvalidValue;
}

}  // namespace sk
