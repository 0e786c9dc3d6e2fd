// Tensor-core engine of the fused distance + top-KC candidate search: tcgen05.mma (kind::tf32)
// with accumulators in TMEM, operands staged by 1-D TMA bulk copies, selection fused into the
// TMEM epilogue so the score matrix never leaves the SM.
//
// Same role as search_simt.cu (it replaces the dgemm + heap-test hot loop of scikit-learn's
// EuclideanArgKmin64, $SP/sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:401-510)
// but the contraction runs on the 5th-generation tensor cores.  TF32 scores carry ~1e-3
// relative error, so this kernel is only ever a FILTER: it returns (at most) the KC best
// references by approximate score plus a threshold that every reference outside the list is
// known to reach; refine.cu re-evaluates the survivors in float64 and proves (error bound
// eps_s = 2^-10) that nothing outside the list can belong to the k nearest; rows it cannot
// certify are re-searched by the FP32 SIMT engine and, failing that, by the exhaustive float64
// kernel.
//
// One CTA = 256 queries = two M=128 MMA tiles that share every reference tile (N=128):
//   warp 8, lane 0   driver: TMA bulk copies (cp.async.bulk + mbarrier) of the query image once
//                    and of reference tiles through an NSTAGE ring, then K/8 x 2 tcgen05.mma
//                    per tile, tcgen05.commit onto the "smem slot free" and "accumulator ready"
//                    mbarriers;
//   warps 0..7       epilogue: thread <-> query (TMEM lane).  The 128 accumulator columns of a
//                    tile are read 32 at a time with tcgen05.ld into two alternating register
//                    buffers (the load of chunk c+1 is in flight while chunk c is reduced), a
//                    min3 tree gives the chunk minimum, one vote tells whether any lane of the
//                    warp beat its threshold.  A lane that did publishes its 32 scores to a
//                    per-warp scratch line, the warp re-tests them one score per lane and the
//                    survivors are appended to the lane's candidate buffer (32 slots per query
//                    in shared memory).  When a buffer overflows the whole warp compacts: every
//                    thread sorts the scores of its own buffer in registers (bitonic network),
//                    keeps the KC smallest and lowers its threshold to the KC-th.
//   two TMEM accumulator stages (2 stages x 2 M tiles x 128 columns = all 512 columns) let the
//   MMAs of tile t+1 overlap the epilogue of tile t.
//
// Threshold seeding.  A streaming top-KC pays KC*ln(n_ref/KC) threshold hits per query, almost
// all of them while the threshold is still loose.  The kernel therefore first runs every
// `seed_stride`-th reference tile in a min-only mode that keeps 32 group minima per query in
// registers; the KC-th smallest group minimum is an upper bound of the query's KC-th best
// score (KC distinct references reach it), so the main pass starts with a threshold close to
// its final value and sees ~KC*(1+ln(1.4*seed_stride)) hits instead.
//
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

#include <type_traits>

namespace sk {

constexpr int TC_MT = 2;                        // M tiles (of 128 queries) per CTA
constexpr int TC_QT = TC_MT * TC_M;             // queries per CTA
constexpr int TC_EPI_WARPS = TC_MT * 4;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1) * 32;
constexpr int TC_CAP = 32;                      // candidate buffer slots per query
constexpr int TC_LD = TC_QT + 1;                // slot stride (odd: a query's slots hit 32 banks)
constexpr int TC_GROUPS = 32;                   // seeding: group minima per query
constexpr uint32_t TC_ROWB = 16;                // bytes of one row of one K chunk (4 TF32)
static_assert(TC_N == 128, "epilogue assumes four 32-column chunks per tile");

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
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread.  Asynchronous: the
// registers are valid only after tmem_ld_wait on the same buffer.
// `dep` (a register of the buffer about to be reduced) is a fake in/out operand: it pins the
// issue ABOVE the reduction of the other buffer so the load overlaps it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32], uint32_t &dep) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "+r"(dep)
        : "r"(taddr)
        : "memory");
}
// Waits for every outstanding tcgen05.ld of this thread.  The buffer is an in/out operand so
// that no use of its registers can be scheduled above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                   "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31])
                 :
                 : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2,
// B=TF32 [10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_M >> 4) << 24);

// ---- thread-parallel register networks ---------------------------------------------------
// ascending bitonic sort of N registers (compile-time network: N/2 * log2(N)*(log2(N)+1)/2
// compare-exchanges of one FMNMX pair each); every lane sorts its own values
template <int N>
__device__ __forceinline__ void sort_regs(float (&s)[N]) {
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool asc = (i & size) == 0;
                    const float lo = fminf(s[i], s[j]), hi = fmaxf(s[i], s[j]);
                    s[i] = asc ? lo : hi;
                    s[j] = asc ? hi : lo;
                }
            }
        }
    }
}

// minimum of the 32 accumulator columns of one chunk (min3 tree: 16 FMNMX3/FMNMX)
__device__ __forceinline__ float tc_min32(const uint32_t (&r)[32]) {
    float a[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        a[i] = fminf(fminf(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1])),
                     __uint_as_float(r[3 * i + 2]));
    a[10] = fminf(__uint_as_float(r[30]), __uint_as_float(r[31]));
    const float b0 = fminf(fminf(a[0], a[1]), a[2]);
    const float b1 = fminf(fminf(a[3], a[4]), a[5]);
    const float b2 = fminf(fminf(a[6], a[7]), a[8]);
    const float b3 = fminf(a[9], a[10]);
    return fminf(fminf(b0, b1), fminf(b2, b3));
}

struct ThrCnt {
    float thr;
    int cnt;
};

// Warp-wide compaction, thread-parallel: every thread reduces its OWN candidate buffer (column
// col_s/col_i, slot stride TC_LD) to the KC smallest scores and lowers its threshold to the
// KC-th smallest.  Entries equal to the new threshold are kept only up to KC entries in total;
// the dropped ones are >= the threshold, which is all the certificate needs.  Called by all 32
// lanes (data-independent network, no divergence).
template <int KC>
__device__ __noinline__ ThrCnt tc_compact_all(float *col_s, int *col_i, float thr, int cnt) {
    __syncwarp();
    float s[TC_CAP];
#pragma unroll
    for (int j = 0; j < TC_CAP; ++j) s[j] = (j < cnt) ? col_s[j * TC_LD] : SK_INF_F;
    sort_regs<TC_CAP>(s);
    const float t = s[KC - 1];  // +inf while the buffer holds fewer than KC entries
    int n_less = 0;
#pragma unroll
    for (int j = 0; j < KC - 1; ++j) n_less += (s[j] < t) ? 1 : 0;
    int quota = KC - n_less;    // entries equal to t that may stay
    int w = 0;
#pragma unroll 4
    for (int j = 0; j < TC_CAP; ++j) {
        const float v = col_s[j * TC_LD];
        const int id = col_i[j * TC_LD];
        const bool valid = j < cnt;
        const bool lt = valid && (v < t);
        const bool eq = valid && (v == t) && quota > 0;
        if (eq) --quota;
        if (lt || eq) {
            col_s[w * TC_LD] = v;
            col_i[w * TC_LD] = id;
            ++w;
        }
    }
    __syncwarp();
    ThrCnt out;
    out.thr = fminf(thr, t);
    out.cnt = w;
    return out;
}

// One 32-column chunk of the main pass.  `r` holds this thread's scores against references
// idb .. idb+31; `inflight` is the other register buffer, whose tcgen05.ld may still be in
// flight (it must land before a function call may spill it).
template <int KC>
__device__ __forceinline__ void tc_process(const uint32_t (&r)[32], uint32_t (&inflight)[32], int idb,
                                           float *buf_s, int *buf_i, float *scratch, int tid,
                                           int lane, float &thr, int &cnt) {
    const float m = tc_min32(r);
    unsigned hits = __ballot_sync(SK_FULL, m < thr);
    while (hits) {  // warp-uniform: one iteration per lane whose chunk minimum beat its threshold
        const int L = __ffs(hits) - 1;
        hits &= hits - 1;
        __syncwarp();
        if (lane == L) {
            float4 *sc = reinterpret_cast<float4 *>(scratch);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                sc[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                    __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
        __syncwarp();
        const float x = scratch[lane];                 // score of lane L's query vs reference idb+lane
        float thrL = __shfl_sync(SK_FULL, thr, L);
        int cntL = __shfl_sync(SK_FULL, cnt, L);
        unsigned pending = __ballot_sync(SK_FULL, x < thrL);
        float *cs = buf_s + (tid - lane + L);
        int *ci = buf_i + (tid - lane + L);
        while (pending) {
            unsigned take = pending;
            if (cntL + __popc(pending) > TC_CAP) {
                if (cntL > KC) {
                    if (lane == L) cnt = cntL;  // entries appended earlier in this loop
                    tmem_ld_wait(inflight);
                    const ThrCnt tc = tc_compact_all<KC>(buf_s + tid, buf_i + tid, thr, cnt);
                    thr = tc.thr;
                    cnt = tc.cnt;
                    thrL = __shfl_sync(SK_FULL, thr, L);
                    cntL = __shfl_sync(SK_FULL, cnt, L);
                    pending &= __ballot_sync(SK_FULL, x < thrL);
                    continue;
                }
                // at most KC entries held, more than CAP - KC = 16 hits: half a warp at a time
                take = pending & 0xffffu;
                if (take == 0) take = pending;
            }
            if ((take >> lane) & 1u) {
                const int slot = cntL + __popc(take & ((1u << lane) - 1u));
                cs[slot * TC_LD] = x;
                ci[slot * TC_LD] = idb + lane;
            }
            cntL += __popc(take);
            pending &= ~take;
        }
        if (lane == L) cnt = cntL;
    }
}

// Epilogue of one reference tile: 4 chunks of 32 columns through the alternating register
// buffers A / B.  On entry the load of chunk 0 into A has been issued; on exit the load of the
// next tile's chunk 0 into A has been issued (if there is a next tile).
template <class F>
__device__ __forceinline__ void tc_epi_tile(uint32_t (&A)[32], uint32_t (&B)[32], uint32_t tbase,
                                            int t, int n_seq, uint64_t *tfull, uint64_t *tempty,
                                            int lane, F &&proc) {
    const int a = t & 1;
    const uint32_t tcol = tbase + (uint32_t)(a * TC_MT * TC_N);
    tmem_ld_wait(A);
    tmem_ld32_issue(tcol + 32, B, A[0]);
    proc(A, B, std::integral_constant<int, 0>{});
    tmem_ld_wait(B);
    tmem_ld32_issue(tcol + 64, A, B[0]);
    proc(B, A, std::integral_constant<int, 1>{});
    tmem_ld_wait(A);
    tmem_ld32_issue(tcol + 96, B, A[0]);
    proc(A, B, std::integral_constant<int, 2>{});
    tmem_ld_wait(B);
    // every TMEM read of this accumulator stage has landed: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[a]);
    if (t + 1 < n_seq) {
        mbar_wait(&tfull[a ^ 1], ((t + 1) >> 1) & 1);
        tc_fence_after();
        tmem_ld32_issue(tbase + (uint32_t)((a ^ 1) * TC_MT * TC_N), A, B[0]);
    }
    proc(B, A, std::integral_constant<int, 3>{});
}

template <int KC>
__global__ void __launch_bounds__(TC_THREADS, 1)
search_tc_kernel(const float *__restrict__ qimg, const float *__restrict__ rimg, int kc_tot,
                 int n_rtiles, int nstage, int n_seed, int seed_stride, long long n_q,
                 int *__restrict__ cand_idx, float *__restrict__ cand_thr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t a_bytes = (uint32_t)kc_tot * TC_M * TC_ROWB;  // one 128-query operand image
    const uint32_t b_bytes = (uint32_t)kc_tot * TC_N * TC_ROWB;  // one 128-plot operand image
    unsigned char *Qs = smem_raw;                                // TC_MT operand images
    unsigned char *Rs = Qs + TC_MT * a_bytes;                    // nstage operand images
    float *buf_s = reinterpret_cast<float *>(Rs + (size_t)nstage * b_bytes);
    int *buf_i = reinterpret_cast<int *>(buf_s + TC_CAP * TC_LD);
    float *scratch_all = reinterpret_cast<float *>(buf_i + TC_CAP * TC_LD);  // [warps][32], 16-B aligned
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch_all + TC_EPI_WARPS * 32);
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
            mbar_init(&tempty[a], TC_EPI_WARPS);
        }
        mbar_init(qbar, 1);
        fence_mbar_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long qtile = blockIdx.x;
    const int ksteps = kc_tot >> 1;  // MMA K = 8 TF32 = two 16-byte chunks
    const int n_seq = n_seed + n_rtiles;  // sampled tiles (seeding), then every tile

    if (warp == TC_EPI_WARPS) {
        // ======================= driver: TMA producer + MMA issuer =======================
        if (lane == 0) {
            auto issue_tile = [&](int tn) {
                const int sn = tn % nstage;
                if (tn >= nstage) mbar_wait(&empty[sn], ((tn / nstage) - 1) & 1);
                const int tile = tn < n_seed ? tn * seed_stride : tn - n_seed;
                mbar_expect_tx(&full[sn], b_bytes);
                bulk_g2s(Rs + (size_t)sn * b_bytes, (const unsigned char *)rimg + (size_t)tile * b_bytes,
                         b_bytes, &full[sn]);
            };
            mbar_expect_tx(qbar, TC_MT * a_bytes);
            bulk_g2s(Qs, (const unsigned char *)qimg + (size_t)qtile * TC_MT * a_bytes, TC_MT * a_bytes, qbar);
            for (int tn = 0; tn < nstage - 1 && tn < n_seq; ++tn) issue_tile(tn);
            mbar_wait(qbar, 0);
            const uint32_t q_addr = smem_u32(Qs), r_addr = smem_u32(Rs);
            const uint32_t a_lbo = TC_M * TC_ROWB, b_lbo = TC_N * TC_ROWB;
            for (int t = 0; t < n_seq; ++t) {
                const int s = t % nstage, a = t & 1;
                mbar_wait(&full[s], (t / nstage) & 1);
                if (t >= 2) mbar_wait(&tempty[a], ((t >> 1) - 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < TC_MT; ++h) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)((a * TC_MT + h) * TC_N);
                    for (int j = 0; j < ksteps; ++j) {
                        const uint64_t adesc = tc_smem_desc(q_addr + h * a_bytes + j * 2 * a_lbo, a_lbo, 128);
                        const uint64_t bdesc = tc_smem_desc(r_addr + s * b_bytes + j * 2 * b_lbo, b_lbo, 128);
                        tc_mma_tf32(d_tmem, adesc, bdesc, TC_IDESC, j > 0 ? 1u : 0u);
                    }
                }
                tc_commit(&empty[s]);   // smem slot reusable once these MMAs have read it
                tc_commit(&tfull[a]);   // accumulators of tile t complete
                if (t + nstage - 1 < n_seq) issue_tile(t + nstage - 1);
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue: thread <-> query =======================
        const int tid = threadIdx.x;          // query slot; TMEM lane (tid & 127) of M tile (tid >> 7)
        const int h = warp >> 2;
        const uint32_t tbase = tmem_base + (((uint32_t)((warp & 3) * 32)) << 16) + (uint32_t)(h * TC_N);
        float *scratch = scratch_all + warp * 32;
        float thr = SK_INF_F;
        int cnt = 0;
        uint32_t A[32], B[32];
        int t = 0;
        mbar_wait(&tfull[0], 0);
        tc_fence_after();
        uint32_t dep0 = 0;
        tmem_ld32_issue(tbase, A, dep0);

        // ---- seeding pass: group minima over the sampled tiles ----
        if (n_seed > 0) {
            float gm[TC_GROUPS];
#pragma unroll
            for (int g = 0; g < TC_GROUPS; ++g) gm[g] = SK_INF_F;
            for (int tb = 0; tb < n_seed; tb += TC_GROUPS / 4) {
#pragma unroll
                for (int j = 0; j < TC_GROUPS / 4; ++j) {
                    if (tb + j < n_seed) {
                        tc_epi_tile(A, B, tbase, t, n_seq, tfull, tempty, lane,
                                    [&](const uint32_t (&r)[32], uint32_t (&)[32], auto ic) {
                                        constexpr int c = decltype(ic)::value;
                                        gm[j * 4 + c] = fminf(gm[j * 4 + c], tc_min32(r));
                                    });
                        ++t;
                    }
                }
            }
            sort_regs<TC_GROUPS>(gm);
            thr = gm[KC - 1];
        }

        // ---- main pass ----
        for (int tile = 0; tile < n_rtiles; ++tile, ++t) {
            const int idb = tile * TC_N;
            tc_epi_tile(A, B, tbase, t, n_seq, tfull, tempty, lane,
                        [&](const uint32_t (&r)[32], uint32_t (&inflight)[32], auto ic) {
                            constexpr int c = decltype(ic)::value;
                            tc_process<KC>(r, inflight, idb + c * 32, buf_s, buf_i, scratch, tid, lane,
                                           thr, cnt);
                        });
        }

        // ---- final compaction, then every thread writes its own candidates ----
        {
            const ThrCnt tc = tc_compact_all<KC>(buf_s + tid, buf_i + tid, thr, cnt);
            thr = tc.thr;
            cnt = tc.cnt;
        }
        const long long q = qtile * TC_QT + tid;
        if (q < n_q) {
            int4 *dst = reinterpret_cast<int4 *>(cand_idx + q * KC);
#pragma unroll
            for (int j = 0; j < KC; j += 4) {
                int4 v;
                v.x = (j + 0 < cnt) ? buf_i[(j + 0) * TC_LD + tid] : -1;
                v.y = (j + 1 < cnt) ? buf_i[(j + 1) * TC_LD + tid] : -1;
                v.z = (j + 2 < cnt) ? buf_i[(j + 2) * TC_LD + tid] : -1;
                v.w = (j + 3 < cnt) ? buf_i[(j + 3) * TC_LD + tid] : -1;
                dst[j / 4] = v;
            }
            cand_thr[q] = thr;  // +inf only when the list holds every reference
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TC_EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

size_t search_tc_smem_bytes(int kc_tot, int nstage) {
    const size_t a = (size_t)kc_tot * TC_M * TC_ROWB, b = (size_t)kc_tot * TC_N * TC_ROWB;
    return TC_MT * a + nstage * b + 2 * (size_t)TC_CAP * TC_LD * 4 + (size_t)TC_EPI_WARPS * 32 * 4 +
           (size_t)(2 * nstage + 5) * 8 + 16;
}

// ring stages that fit the 227 KB of shared memory (0 = the shape does not fit the engine)
int search_tc_pick_stages(int kc_tot) {
    for (int s = 4; s >= 2; --s)
        if (search_tc_smem_bytes(kc_tot, s) <= 227 * 1024) return s;
    return 0;
}

// sampled tiles of the seeding pass (0 = no seeding: too few references for it to pay)
int search_tc_seed_tiles(int n_rtiles, int seed_stride) {
    if (seed_stride <= 0 || n_rtiles < 64) return 0;
    return (n_rtiles + seed_stride - 1) / seed_stride;
}

template <int KC>
static cudaError_t launch_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles, int nstage,
                             int seed_stride, long long n_q, int *cand_idx, float *cand_thr,
                             cudaStream_t st) {
    const size_t smem = search_tc_smem_bytes(kc_tot, nstage);
    cudaError_t e = cudaFuncSetAttribute(search_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    const long long n_qtiles = (n_q + TC_QT - 1) / TC_QT;
    const int n_seed = search_tc_seed_tiles(n_rtiles, seed_stride);
    search_tc_kernel<KC><<<(unsigned)n_qtiles, TC_THREADS, smem, st>>>(
        qimg, rimg, kc_tot, n_rtiles, nstage, n_seed, seed_stride, n_q, cand_idx, cand_thr);
    return cudaGetLastError();
}

cudaError_t launch_search_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles,
                             long long n_q, int kc, int nstage, int seed_stride, int *cand_idx,
                             float *cand_thr, cudaStream_t st) {
    if (n_q <= 0) return cudaSuccess;
    if (kc == 8) return launch_tc<8>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q, cand_idx, cand_thr, st);
    if (kc == 16) return launch_tc<16>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q, cand_idx, cand_thr, st);
    return cudaErrorInvalidValue;
}

}  // namespace sk
