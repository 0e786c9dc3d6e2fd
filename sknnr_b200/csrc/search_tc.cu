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
// One CTA = 256 queries = two M=128 MMA tiles that share every 128-plot reference tile (N=128).  A
// "job" is one (reference tile, M tile) pair = K/8 tcgen05.mma into one of four 128-column TMEM
// accumulator slots; jobs run in (tile, M tile) order and job j uses slot j % 4, i.e. every M tile
// owns two slots and its MMAs run at most one job ahead of its scanners.  19 warps (ns = 2):
//   TMA producer (1 thread)   bulk copies (cp.async.bulk + mbarrier) of the query image once and of
//                    the reference tiles through an NSTAGE ring;
//   MMA issuers (one warp per M tile, warp-uniform loop, one elected lane)  wait "tile staged" and
//                    "slot drained", issue the job's tcgen05.mma, tcgen05.commit onto "slot ready"
//                    and (both issuers) "stage free";
//   16 scanner warps = (column stream p, M tile h, TMEM lane quarter): thread <-> (query, stream).
//                    Per job a warp waits "slot ready", reads its stream's 64 accumulator columns
//                    with one tcgen05.ld.x64, hands the slot back BEFORE reducing anything, then per
//                    32-column chunk: min3 tree -> one vote "did any lane beat its threshold".  If
//                    so the hit lanes descend the tree in-lane (group of 9 -> triple -> values) and
//                    append (score, index) to their own candidate column in shared memory with
//                    branch-free stores; the count of a column lives in the offset of its next free
//                    slot (power-of-two slot stride).  When a column is nearly full the whole warp
//                    compacts: every thread sorts its own column in registers (bitonic network),
//                    keeps the KC smallest and lowers its threshold to the KC-th.  A lane that runs
//                    out of slots inside one chunk is redone through a cooperative path (publish
//                    the 32 scores to the warp's scratch line, one score per lane).
//   All role / slot / barrier values derive from a shuffled (provably warp-uniform) warp index and
//   live in uniform registers: the scanners have 96 vector registers, 64 of them hold a job.
//   (ns = 1: one stream of 16 candidates, 8 scanner warps of four chunks, for k (+1) <= 14.)
//
// Threshold seeding.  A streaming top-KC pays KC*ln(n_ref/KC) threshold hits per query, almost
// all of them while the threshold is still loose.  The kernel therefore first runs every
// `seed_stride`-th reference tile in a min-only mode that keeps 32 group minima per (query, stream)
// in the still unused candidate column; the KC-th smallest group minimum is an upper bound of the
// stream's KC-th best score (KC distinct references reach it), so the main pass starts with a
// threshold close to its final value: ~20 hits per stream instead of ~65.
//
// The |r|^2 term is folded into the contraction: each operand gets one extra K block holding
// (1,1,1,0,..) on the query side and a 3-way TF32 split of |r|^2 on the reference side, so the
// accumulator is directly s = |r|^2 - 2 q.r.
//
// Operand layout (no swizzle, K-major "interleaved" canonical layout): 16-byte K chunks of 4
// TF32 values, [chunk][row][4]; core matrix = 8 rows x 16 B contiguous, SBO = 128 B between row
// groups, LBO = rows * 16 B between K chunks.  The images are pre-arranged in HBM in exactly
// this order, so a plain bulk copy stages them.
#include <cstdio>
#include "common.cuh"
#include "kernels.h"

#include <type_traits>

namespace sk {

constexpr int TC_SLOTS = 4;                     // TMEM accumulator slots of TC_N columns
constexpr int TC_GROUPS = 32;                   // seeding: group minima per query
// MT M tiles (of 128 queries) per CTA; NS independent candidate streams per query: stream p owns
// columns [p * 128 / NS, (p + 1) * 128 / NS) of every reference tile and has its own scanner
// warps, candidate buffer and threshold; KCS candidates kept per stream, CAP buffer slots.
template <int MT_, int NS_, int CAP_> struct TcCfg {
    static constexpr int MT = MT_, NS = NS_, CAP = CAP_;
    static constexpr int QT = MT * TC_M;                  // queries per CTA
    static constexpr int EPI_WARPS = MT * 4 * NS;         // scanner warps: (stream, M tile, lane quarter)
    static constexpr int THREADS = (EPI_WARPS + MT + 1) * 32;  // + MT MMA issuer warps + TMA producer warp
    // slot stride of the candidate buffers: a power of two, so that a thread's slot count sits in the
    // high bits of its buffer offset (4 * column + count * LD * 4).  Threads work on their own
    // columns (bank = column mod 32, no conflicts); only the rare cooperative path, where a warp
    // writes into ONE column, serialises on a bank.
    static constexpr int LD = QT * NS;
    static constexpr int SORT = CAP <= 16 ? 16 : 32;      // width of the register sorting network
    static_assert(CAP <= SORT && CAP % 2 == 0, "candidate buffer shape");
    static constexpr int CH = 4 / NS;                     // 32-column chunks a scanner warp reads per job
};
constexpr uint32_t TC_ROWB = 16;                // bytes of one row of one K chunk (4 TF32)
static_assert(TC_N == 128, "epilogue assumes four 32-column chunks per tile");
static_assert(TC_SLOTS * TC_N == 512, "the accumulator slots fill TMEM");

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
// true in exactly one (converged) lane of the warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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
// Both 32-column chunks of a job in ONE instruction (64 consecutive columns): issued as two .x32
// loads, ptxas sinks the second one below the first chunk's min tree and the slot release waits for it.
__device__ __forceinline__ void tmem_ld64_issue(uint32_t taddr, uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31]), "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]), "=r"(b[16]), "=r"(b[17]), "=r"(b[18]), "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]), "=r"(b[24]), "=r"(b[25]), "=r"(b[26]), "=r"(b[27]), "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
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

// diagnostics of the debug instantiation (tc_debug bit 3): chunk visits, events, hit lanes,
// proactive compactions, cooperative fallbacks - printed by the last CTA
__device__ unsigned long long g_tc_counters[8];

struct ThrCnt {
    float thr;
    int cnt;
};

// Warp-wide compaction, thread-parallel: every thread reduces its OWN candidate buffer (column
// cs0, slot stride LD) to the KC smallest scores and lowers its threshold to the
// KC-th smallest.  Entries equal to the new threshold are kept only up to KC entries in total;
// the dropped ones are >= the threshold, which is all the certificate needs.  Called by all 32
// lanes (data-independent network, no divergence).
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// (the slow paths take 32-bit shared-memory addresses, not generic pointers: `cs0` = address of
// slot 0 of the calling thread's score column; slot j lies j * LD * 4 bytes further and the index
// column CAP * LD * 4 bytes behind the score column - nothing else has to stay live in the scanner
// loop on their behalf)
template <int KC, int CAP, int LD>
__device__ __noinline__ ThrCnt tc_compact_all(uint32_t cs0, float thr, int cnt) {
    static_assert(KC < CAP, "the buffer needs slack above KC");
    constexpr int SORT = CAP <= 16 ? 16 : 32;
    constexpr uint32_t IOFF = CAP * LD * 4, STEP = LD * 4;
    __syncwarp();
    float s[SORT];
#pragma unroll
    for (int j = 0; j < SORT; ++j) s[j] = (j < CAP && j < cnt) ? lds_f32(cs0 + j * STEP) : SK_INF_F;
    sort_regs<SORT>(s);
    const float t = s[KC - 1];  // +inf while the buffer holds fewer than KC entries
    int n_less = 0;
#pragma unroll
    for (int j = 0; j < KC - 1; ++j) n_less += (s[j] < t) ? 1 : 0;
    int quota = KC - n_less;    // entries equal to t that may stay
    uint32_t w = cs0;
    int nw = 0;
#pragma unroll 4
    for (int j = 0; j < CAP; ++j) {
        const float v = lds_f32(cs0 + j * STEP);
        const int id = lds_s32(cs0 + IOFF + j * STEP);
        const bool valid = j < cnt;
        const bool lt = valid && (v < t);
        const bool eq = valid && (v == t) && quota > 0;
        if (eq) --quota;
        if (lt || eq) {
            st_shared_b32(w, __float_as_uint(v));
            st_shared_b32(w + IOFF, (uint32_t)id);
            w += STEP;
            ++nw;
        }
    }
    __syncwarp();
    ThrCnt out;
    out.thr = fminf(thr, t);
    out.cnt = nw;
    return out;
}

// Cooperative (slow, always safe) hit path for ONE lane L whose 32 scores (references idb ..
// idb+31) have been published to the warp's scratch line: the warp re-tests them one score per
// lane and appends the survivors to lane L's candidate buffer, compacting the warp's buffers
// whenever it would overflow.  Used when the in-lane path below ran out of buffer slots.
template <int KC, int CAP, int LD>
__device__ __noinline__ ThrCnt tc_process_coop(int L, int idb, uint32_t cs0, uint32_t scratch_a, int lane,
                                               float thr, int cnt) {
    constexpr uint32_t IOFF = CAP * LD * 4, STEP = LD * 4;
    const float x = lds_f32(scratch_a + 4u * lane);   // score of lane L's query vs reference idb+lane
    float thrL = __shfl_sync(SK_FULL, thr, L);
    int cntL = __shfl_sync(SK_FULL, cnt, L);
    unsigned pending = __ballot_sync(SK_FULL, x < thrL);
    const uint32_t csL = cs0 + 4u * (uint32_t)(L - lane);   // lane L's column
    while (pending) {
        unsigned take = pending;
        if (cntL + __popc(pending) > CAP) {
            if (cntL > KC) {
                if (lane == L) cnt = cntL;  // entries appended earlier in this loop
                const ThrCnt tc = tc_compact_all<KC, CAP, LD>(cs0, thr, cnt);
                thr = tc.thr;
                cnt = tc.cnt;
                thrL = __shfl_sync(SK_FULL, thr, L);
                cntL = __shfl_sync(SK_FULL, cnt, L);
                pending &= __ballot_sync(SK_FULL, x < thrL);
                continue;
            }
            // at most KC entries held but more hits than free slots: lowest hits first
            const int room = CAP - cntL;
            while (__popc(take) > room) take &= ~(0x80000000u >> __clz(take));
        }
        if ((take >> lane) & 1u) {
            const int slot = cntL + __popc(take & ((1u << lane) - 1u));
            st_shared_b32(csL + slot * STEP, __float_as_uint(x));
            st_shared_b32(csL + IOFF + slot * STEP, (uint32_t)(idb + lane));
        }
        cntL += __popc(take);
        pending &= ~take;
    }
    if (lane == L) cnt = cntL;
    ThrCnt out;
    out.thr = thr;
    out.cnt = cnt;
    return out;
}

// in-lane append of the values of v[0..N) below thr (references id0 ..), N <= 3: every value is
// stored at the lane's next free slot and the slot only advances for the values that qualify, so
// there is no branch per value.  The caller guarantees N free slots.
// `ps` = shared-memory address of the lane's next free score slot (the index slot lies IOFF bytes
// further); see the comment above for the protocol
template <int N, int LD, int IOFF>
__device__ __forceinline__ void tc_leaf(const float *v, int id0, float thr, uint32_t bs0, uint32_t &pr) {
    // (three independent store groups: a "minimum first, others only if the median qualifies"
    // variant has fewer instructions but a longer dependent chain and measured 8 % slower)
#pragma unroll
    for (int j = 0; j < N; ++j) {
        st_shared_b32(bs0 + pr, __float_as_uint(v[j]));
        st_shared_b32(bs0 + pr + IOFF, (uint32_t)(id0 + j));
        if (v[j] < thr) pr += LD * 4;
    }
}

// One 32-column chunk of the main pass.  `r` holds this thread's scores against references
// idb .. idb+31.  Fast path: min3 tree + one vote.  Hit path, in the hit lanes only and without
// any cross-lane traffic: the tree's intermediate minima (four groups of <= 9 values) locate the
// values below the threshold, which the lane appends to its own candidate buffer column.
// `cs0` = shared-memory address of slot 0 of this thread's candidate column.
// The thread's candidate count lives in `pr`, the offset of its next free score slot from the
// start `bs0` of the score buffer: pr = 4 * column + count * LD * 4 with LD * 4 a power of two above
// 4 * column, so "fewer than three free slots" is one compare of pr with a constant and the hit
// path only ever bumps that offset.
template <int KC, int CAP, int LD, int EPI_WARPS, bool DBG>
__device__ __forceinline__ void tc_process(const uint32_t (&r)[32], int idb, uint32_t bs0, int lane, float &thr,
                                           uint32_t &pr, int dbg) {
    constexpr uint32_t STEP = LD * 4, IOFF = CAP * LD * 4, FULL = (CAP - 2) * STEP;
    static_assert((STEP & (STEP - 1)) == 0, "slot stride must be a power of two");
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    // group minima b0..b3 over v[0..9), v[9..18), v[18..27), v[27..32).  The fast path reduces each
    // group through STRIDED triples (v[g], v[g+3], v[g+6]); the hit path below re-derives the
    // minima of the CONTIGUOUS triples it descends into, so that those eleven values are not kept
    // live (and spilled) across the fast path.
    float b[4];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const float *w = v + 9 * g;
        b[g] = fminf(fminf(fminf(fminf(w[0], w[3]), w[6]), fminf(fminf(w[1], w[4]), w[7])),
                     fminf(fminf(w[2], w[5]), w[8]));
    }
    b[3] = fminf(fminf(fminf(v[27], v[29]), v[31]), fminf(v[28], v[30]));
    const float b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];
    const float m = fminf(fminf(b0, b1), fminf(b2, b3));
    bool hit = m < thr;
    unsigned hits = 1u;
    if constexpr (DBG) {
        hits = __ballot_sync(SK_FULL, hit);
        if ((dbg & 8) && lane == 0) {
            atomicAdd(&g_tc_counters[0], 1ull);
            if (hits) atomicAdd(&g_tc_counters[1], 1ull);
            atomicAdd(&g_tc_counters[2], (unsigned long long)__popc(hits));
        }
        if (hits == 0u) return;
    } else {
        if (!__any_sync(SK_FULL, hit)) return;   // vote straight into a predicate
    }
    if (DBG && (dbg & 1)) {  // timing experiment: count the hits, skip the hit path (results are wrong)
        pr = (pr & (STEP - 1)) + (((pr / STEP) + __popc(hits)) & 7) * STEP;
        return;
    }
    // a triple is appended only while three slots are free (pr < FULL); keep room for the usual one
    // or two appends - compaction also refreshes the thresholds
    if (__any_sync(SK_FULL, hit && pr >= FULL)) {
        if (DBG && (dbg & 8) && lane == 0) atomicAdd(&g_tc_counters[3], 1ull);
        const uint32_t c4 = pr & (STEP - 1);   // 4 * column
        const ThrCnt tc = tc_compact_all<KC, CAP, LD>(bs0 + c4, thr, (int)(pr / STEP));
        thr = tc.thr;
        pr = c4 + (uint32_t)tc.cnt * STEP;
        hit = m < thr;
    }
    const uint32_t pr0 = pr;
    if (hit) {
        // descend the tree: group of <= 9 values -> triple -> values.  Out of room: bit 0 of pr is set
        // (offsets are multiples of 4) and pr >= FULL stays true
#define SK_TC_TRIPLE(I, N)                                                      \
        if (fminf(fminf(v[3 * (I)], v[3 * (I) + 1]), v[3 * (I) + ((N) == 3 ? 2 : 1)]) < thr) { \
            if (pr >= FULL) pr |= 1u;                                           \
            else tc_leaf<N, LD, IOFF>(v + 3 * (I), idb + 3 * (I), thr, bs0, pr); \
        }
        if (b0 < thr) { SK_TC_TRIPLE(0, 3) SK_TC_TRIPLE(1, 3) SK_TC_TRIPLE(2, 3) }
        if (b1 < thr) { SK_TC_TRIPLE(3, 3) SK_TC_TRIPLE(4, 3) SK_TC_TRIPLE(5, 3) }
        if (b2 < thr) { SK_TC_TRIPLE(6, 3) SK_TC_TRIPLE(7, 3) SK_TC_TRIPLE(8, 3) }
        if (b3 < thr) { SK_TC_TRIPLE(9, 3) SK_TC_TRIPLE(10, 2) }
#undef SK_TC_TRIPLE
    }
    // rare: a lane found more values than it had free slots -> undo its appends and redo the
    // chunk for it through the cooperative path (which compacts as often as needed)
    unsigned ovf = __ballot_sync(SK_FULL, (pr & 1u) != 0u);
    if (ovf) {
        if (DBG && (dbg & 8) && lane == 0) atomicAdd(&g_tc_counters[4], (unsigned long long)__popc(ovf));
        if (pr & 1u) pr = pr0;
        float thr2 = thr;
        int cnt = (int)(pr / STEP);
        const uint32_t c4 = pr & (STEP - 1), cs0 = bs0 + c4;
        // the warp's scratch line lies in front of the candidate buffers: [EPI_WARPS][32] floats
        const uint32_t scratch_a = bs0 - (uint32_t)EPI_WARPS * 128u + (c4 >> 7) * 128u;
        while (ovf) {  // warp-uniform
            const int L = __ffs(ovf) - 1;
            ovf &= ovf - 1;
            __syncwarp();
            if (lane == L) {
#pragma unroll
                for (int i = 0; i < 32; ++i) st_shared_b32(scratch_a + 4u * i, __float_as_uint(v[i]));
            }
            __syncwarp();
            const ThrCnt tc = tc_process_coop<KC, CAP, LD>(L, idb, cs0, scratch_a, lane, thr2, cnt);
            thr2 = tc.thr;
            cnt = tc.cnt;
        }
        thr = thr2;
        pr = c4 + (uint32_t)cnt * STEP;
    }
}

// Epilogue of one job (reference tile x this warp's M tile): the warp's CH chunks of 32 columns
// are all read into registers first and the accumulator slot is handed back to the MMA issuer
// BEFORE any of them is reduced, so the MMAs of the slot's next job overlap the reduction and a
// slow hit path never holds TMEM.
template <int CH, class F>
__device__ __forceinline__ void tc_epi_job(uint32_t (&R)[CH][32], uint32_t tcol, uint32_t afull_addr,
                                           uint32_t parity, uint32_t aempty_addr, int lane, F &&proc) {
    mbar_wait_addr(afull_addr, parity);
    tc_fence_after();
    if constexpr (CH == 2) {
        tmem_ld64_issue(tcol, R[0], R[1]);
    } else if constexpr (CH == 4) {
        tmem_ld64_issue(tcol, R[0], R[1]);
        tmem_ld64_issue(tcol + 64, R[2], R[3]);
    } else {
        uint32_t dep = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) tmem_ld32_issue(tcol + 32 * c, R[c], dep);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) tmem_ld_wait(R[c]);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_addr(aempty_addr);
    if constexpr (CH >= 1) proc(R[0], std::integral_constant<int, 0>{});
    if constexpr (CH >= 2) proc(R[1], std::integral_constant<int, 1>{});
    if constexpr (CH >= 3) proc(R[2], std::integral_constant<int, 2>{});
    if constexpr (CH >= 4) proc(R[3], std::integral_constant<int, 3>{});
}

// (register budget: the register file is allocated per 4 warps, so the 18-warp dual-stream CTA
// gets 65536 / (20 * 32) = 102 -> 96 registers per thread and the 10-warp one 168)
template <int KC, int MT, int NS, int CAP, bool DBG>
__global__ void __launch_bounds__(TcCfg<MT, NS, CAP>::THREADS, 1)
search_tc_kernel(const float *__restrict__ qimg, const float *__restrict__ rimg, int kc_tot,
                 int n_rtiles, int nstage, int n_seed, int seed_stride, long long n_q,
                 int *__restrict__ cand_idx, float *__restrict__ cand_thr, int dbg) {
    using Cfg = TcCfg<MT, NS, CAP>;
    constexpr int LD = Cfg::LD, EPI_WARPS = Cfg::EPI_WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t a_bytes = (uint32_t)kc_tot * TC_M * TC_ROWB;  // one 128-query operand image
    const uint32_t b_bytes = (uint32_t)kc_tot * TC_N * TC_ROWB;  // one 128-plot operand image
    unsigned char *Qs = smem_raw;                                // MT operand images
    unsigned char *Rs = Qs + MT * a_bytes;                       // nstage operand images
    float *scratch_all = reinterpret_cast<float *>(Rs + (size_t)nstage * b_bytes);  // [warps][32]
    float *buf_s = scratch_all + EPI_WARPS * 32;
    int *buf_i = reinterpret_cast<int *>(buf_s + CAP * LD);
    uint64_t *full = reinterpret_cast<uint64_t *>(buf_i + CAP * LD);  // (CAP * LD * 4) % 8 == 0
    uint64_t *empty = full + nstage;
    uint64_t *afull = empty + nstage;       // [4] accumulator slot ready
    uint64_t *aempty = afull + TC_SLOTS;    // [4] accumulator slot drained
    uint64_t *qbar = aempty + TC_SLOTS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qbar + 1);

    // the shuffle tells the compiler that `warp` (and every role / slot / address value derived from
    // it) is warp-uniform: those live in uniform registers, not in the scanners' 96 vector registers
    const int warp = __shfl_sync(SK_FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], MT);   // one tcgen05.commit per MMA issuer
        }
        for (int a = 0; a < TC_SLOTS; ++a) {
            mbar_init(&afull[a], 1);
            mbar_init(&aempty[a], 4 * NS);  // the scanner warps of the M tile that used the slot
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
    const int n_seq = n_seed + n_rtiles;  // sampled tiles (seeding), then every tile

    if (warp == EPI_WARPS + MT) {
        // ======================= TMA producer (one thread) =======================
        // No divisions, no descriptor rebuilds in these two single-thread loops: each of their
        // instructions is on the critical path of the tensor pipe.
        if (lane == 0) {
            mbar_expect_tx(qbar, MT * a_bytes);
            bulk_g2s(Qs, (const unsigned char *)qimg + (size_t)qtile * MT * a_bytes, MT * a_bytes, qbar);
            int sn = 0;
            uint32_t wrap_par = 1;   // parity of (uses of the slot so far - 1); first pass: no wait
            bool wrapped = false;
            const unsigned char *rbase = (const unsigned char *)rimg;
            for (int tn = 0; tn < n_seq; ++tn) {
                if (wrapped) mbar_wait(&empty[sn], wrap_par);
                const int tile = tn < n_seed ? tn * seed_stride : tn - n_seed;
                if (DBG && (dbg & 4) && wrapped) {
                    mbar_arrive(&full[sn]);   // timing experiment: reuse the stale tile, no L2 traffic
                } else {
                    mbar_expect_tx(&full[sn], b_bytes);
                    bulk_g2s(Rs + (size_t)sn * b_bytes, rbase + (size_t)tile * b_bytes, b_bytes, &full[sn]);
                }
                if (++sn == nstage) {
                    sn = 0;
                    wrap_par = wrapped ? (wrap_par ^ 1u) : 0u;
                    wrapped = true;
                }
            }
        }
        __syncwarp();
    } else if (warp >= EPI_WARPS) {
        // ======================= MMA issuers: one warp per M tile =======================
        // Issuing is effectively synchronous (the thread gets the next tcgen05.mma out at the
        // rate the tensor pipe retires them), and barrier waits + bookkeeping cost a lone warp
        // another ~300 cycles per job during which ITS MMAs are not being queued; with one issuer
        // per M tile the other issuer's MMAs fill that gap.
        // The whole warp runs the loop in lock step (uniform control flow keeps descriptors and
        // counters in uniform registers); one elected lane issues the tcgen05 instructions.
        const int h = warp - EPI_WARPS;
        mbar_wait(qbar, 0);
        const uint32_t a_lbo = TC_M * TC_ROWB, b_lbo = TC_N * TC_ROWB;
        // descriptor = {hi: SBO = 128 B, version 1; lo: start address >> 4 | LBO >> 4 << 16}; moving
        // to the next K step (two 16-byte chunks) or operand image only adds to the address field
        const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
        const uint32_t a_lo0 = (((smem_u32(Qs) + h * a_bytes) >> 4) & 0x3fffu) | (((a_lbo >> 4) & 0x3fffu) << 16);
        const uint32_t b_lo0 = ((smem_u32(Rs) >> 4) & 0x3fffu) | (((b_lbo >> 4) & 0x3fffu) << 16);
        const uint32_t a_kstep = (2 * a_lbo) >> 4, b_kstep = (2 * b_lbo) >> 4;
        const uint32_t b_img = b_bytes >> 4;
        int s = 0;
        uint32_t full_par = 0, b_lo_s = b_lo0;
        for (int t = 0; t < n_seq; ++t) {
            const int j = t * MT + h;                    // this issuer's job on tile t
            const uint32_t sl = (uint32_t)j & (TC_SLOTS - 1);
            mbar_wait(&full[s], full_par);
            if (j >= TC_SLOTS) mbar_wait(&aempty[sl], (uint32_t)(((j >> 2) - 1) & 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + sl * TC_N;
            if (elect_one()) {
                if (DBG && (dbg & 16)) {   // timing experiment: the job's MMAs issued twice (same result)
                    uint32_t a2 = a_lo0, b2 = b_lo_s;
                    tc_mma_tf32(d_tmem, desc_hi | a2, desc_hi | b2, TC_IDESC, 0u);
                    for (int ks = 1; ks < ksteps; ++ks) {
                        a2 += a_kstep;
                        b2 += b_kstep;
                        tc_mma_tf32(d_tmem, desc_hi | a2, desc_hi | b2, TC_IDESC, 1u);
                    }
                }
                uint32_t a_lo = a_lo0, b_lo = b_lo_s;
                tc_mma_tf32(d_tmem, desc_hi | a_lo, desc_hi | b_lo, TC_IDESC, 0u);
#pragma unroll 4
                for (int ks = 1; ks < ksteps; ++ks) {
                    a_lo += a_kstep;
                    b_lo += b_kstep;
                    tc_mma_tf32(d_tmem, desc_hi | a_lo, desc_hi | b_lo, TC_IDESC, 1u);
                }
                tc_commit(&afull[sl]);   // accumulators of job j complete
                tc_commit(&empty[s]);    // one of the MT arrivals that free the smem slot
            }
            __syncwarp();
            b_lo_s += b_img;
            if (++s == nstage) {
                s = 0;
                full_par ^= 1u;
                b_lo_s = b_lo0;
            }
        }
    } else {
        // ======================= epilogue: thread <-> (query, stream) =======================
        // warp = (stream p, M tile h, lane quarter); every warp takes part in every job of its M
        // tile and reads the stream's CH chunks of the 128 accumulator columns
        constexpr int CH = Cfg::CH;
        const int col = threadIdx.x;          // candidate buffer column of this (query, stream)
        const int p = warp / (MT * 4);
        const int h = (warp >> 2) % MT;
        const int qslot = h * TC_M + (warp & 3) * 32 + lane;   // query within the CTA = TMEM lane of M tile h
        const uint32_t tlane = tmem_base + (((uint32_t)((warp & 3) * 32)) << 16) + (uint32_t)(p * CH * 32);
        float thr = SK_INF_F;
        uint32_t R[CH][32];
        int t = 0;                            // position in the tile sequence; this warp's job = t * MT + h
        const uint32_t afull_a0 = smem_u32(afull), aempty_a0 = smem_u32(aempty);
        const uint32_t cs0 = smem_u32(buf_s + col);

        // ---- seeding pass: group minima over the sampled tiles ----
        // (the 32 running minima live in the still unused candidate buffer column: slots of
        // buf_s, then of buf_i; chunk c of the n-th sampled tile feeds group (CH * n + c) % 32)
        if (n_seed > 0) {
            static_assert(2 * CAP >= TC_GROUPS, "group minima are parked in the candidate buffers");
            float *gcol = buf_s + col;
#pragma unroll
            for (int g = 0; g < TC_GROUPS; ++g) gcol[g * LD] = SK_INF_F;   // buf_i follows buf_s
            for (; t < n_seed; ++t) {
                const int j = t * MT + h, sl = j & (TC_SLOTS - 1);
                const int g0 = (t * CH) & (TC_GROUPS - 1);
                tc_epi_job<CH>(R, tlane + (uint32_t)(sl * TC_N), afull_a0 + 8u * sl, (uint32_t)((j >> 2) & 1),
                               aempty_a0 + 8u * sl, lane,
                               [&](const uint32_t (&r)[32], auto ic) {
                                   constexpr int c = decltype(ic)::value;
                                   float *g = gcol + (g0 + c) * LD;
                                   *g = fminf(*g, tc_min32(r));
                               });
            }
            float gm[TC_GROUPS];
#pragma unroll
            for (int g = 0; g < TC_GROUPS; ++g) gm[g] = gcol[g * LD];
            sort_regs<TC_GROUPS>(gm);
            thr = gm[KC - 1];
        }

        // ---- main pass ----
        // shared-window address of buf_s, derived from the array's own shared address (no generic
        // pointer round trip): uniform
        const uint32_t bs0 = smem_u32(smem_raw) + MT * a_bytes + (uint32_t)nstage * b_bytes + EPI_WARPS * 128u;
        uint32_t pr = 4u * (uint32_t)col;       // offset of this thread's next free slot (4 * col + count * LD * 4)
        for (; t < n_seq; ++t) {
            const int j = t * MT + h, sl = j & (TC_SLOTS - 1);
            const int idb = (t - n_seed) * TC_N + p * CH * 32;   // warp-uniform, like t and p
            tc_epi_job<CH>(R, tlane + (uint32_t)(sl * TC_N), afull_a0 + 8u * sl, (uint32_t)((j >> 2) & 1),
                               aempty_a0 + 8u * sl, lane,
                           [&](const uint32_t (&r)[32], auto ic) {
                               constexpr int c = decltype(ic)::value;
                               tc_process<KC, CAP, LD, EPI_WARPS, DBG>(r, idb + c * 32, bs0, lane, thr, pr, dbg);
                           });
        }

        // ---- final compaction, then every thread writes the candidates of its (query, stream) ----
        int cnt;
        {
            const ThrCnt tc = tc_compact_all<KC, CAP, LD>(cs0, thr, (int)(pr / (uint32_t)(LD * 4)));
            thr = tc.thr;
            cnt = tc.cnt;
        }
        const long long q = qtile * Cfg::QT + qslot;
        if (q < n_q) {
            constexpr int KOUT = 16 / NS;   // list slots per stream in the output (unused ones hold -1)
            int4 *dst = reinterpret_cast<int4 *>(cand_idx + (q * NS + p) * KOUT);
#pragma unroll
            for (int jj = 0; jj < KOUT; jj += 4) {
                int4 v;
                v.x = (jj + 0 < cnt) ? buf_i[(jj + 0) * LD + col] : -1;
                v.y = (jj + 1 < cnt) ? buf_i[(jj + 1) * LD + col] : -1;
                v.z = (jj + 2 < cnt) ? buf_i[(jj + 2) * LD + col] : -1;
                v.w = (jj + 3 < cnt) ? buf_i[(jj + 3) * LD + col] : -1;
                dst[jj / 4] = v;
            }
            cand_thr[q * NS + p] = thr;  // +inf only when the list holds every reference of the stream
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
    if (DBG && (dbg & 8) && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        printf("tc counters (up to the last CTA): chunks %llu events %llu hit-lanes %llu compactions %llu coop %llu\n",
               g_tc_counters[0], g_tc_counters[1], g_tc_counters[2], g_tc_counters[3], g_tc_counters[4]);
}

int g_tc_debug = 0;  // timing experiments only (set through the "tc_debug" option)

// Two configurations:
//   ns = 2: two streams of KCS = 8 candidates (CAP 16) per query, 16 scanner warps -- k (+1) <= 8
//   ns = 1: one stream of 16 candidates (CAP 32), 8 scanner warps              -- k (+1) <= 14
static constexpr int TC_MT = 2;
#ifndef SK_TC_KCS
#define SK_TC_KCS 8   // candidates kept per stream of the dual-stream layout (<= 8)
#endif

size_t search_tc_smem_bytes(int kc_tot, int nstage, int ns) {
    const size_t a = (size_t)kc_tot * TC_M * TC_ROWB, b = (size_t)kc_tot * TC_N * TC_ROWB;
    const size_t ld = (size_t)TC_MT * TC_M * ns, cap = ns == 2 ? 16 : 32;
    return TC_MT * a + nstage * b + (size_t)TC_MT * 4 * ns * 32 * 4 + 2 * cap * ld * 4 +
           (size_t)(2 * nstage + 2 * TC_SLOTS + 1) * 8 + 16;
}

// ring stages for this contraction depth (0: the shape does not fit the engine)
int search_tc_pick_stages(int kc_tot) {
    for (int s = 4; s >= 2; --s)
        if (search_tc_smem_bytes(kc_tot, s, 2) <= 227 * 1024 && search_tc_smem_bytes(kc_tot, s, 1) <= 227 * 1024)
            return s;
    return 0;
}

// sampled tiles of the seeding pass (0 = no seeding: too few references for it to pay)
int search_tc_seed_tiles(int n_rtiles, int seed_stride) {
    if (seed_stride <= 0 || n_rtiles < 64) return 0;
    return (n_rtiles + seed_stride - 1) / seed_stride;
}

template <int KC, int MT, int NS, int CAP, bool DBG>
static cudaError_t launch_tc_dbg(const float *qimg, const float *rimg, int kc_tot, int n_rtiles, int nstage,
                                 int seed_stride, long long n_q, int *cand_idx, float *cand_thr,
                                 cudaStream_t st) {
    using Cfg = TcCfg<MT, NS, CAP>;
    const size_t smem = search_tc_smem_bytes(kc_tot, nstage, NS);
    cudaError_t e = cudaFuncSetAttribute(search_tc_kernel<KC, MT, NS, CAP, DBG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const long long n_qtiles = (n_q + Cfg::QT - 1) / Cfg::QT;
    const int n_seed = search_tc_seed_tiles(n_rtiles, seed_stride);
    search_tc_kernel<KC, MT, NS, CAP, DBG><<<(unsigned)n_qtiles, Cfg::THREADS, smem, st>>>(
        qimg, rimg, kc_tot, n_rtiles, nstage, n_seed, seed_stride, n_q, cand_idx, cand_thr, g_tc_debug);
    return cudaGetLastError();
}

template <int KC, int MT, int NS, int CAP>
static cudaError_t launch_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles, int nstage,
                             int seed_stride, long long n_q, int *cand_idx, float *cand_thr,
                             cudaStream_t st) {
    // the timing-experiment hooks ("tc_debug") live in a separate instantiation: none of their
    // tests is compiled into the product kernel
    if (g_tc_debug)
        return launch_tc_dbg<KC, MT, NS, CAP, true>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q,
                                                    cand_idx, cand_thr, st);
    return launch_tc_dbg<KC, MT, NS, CAP, false>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q,
                                                 cand_idx, cand_thr, st);
}

// cand_idx [n_q][16] (ns lists of 16 / ns entries), cand_thr [n_q][ns]
cudaError_t launch_search_tc(const float *qimg, const float *rimg, int kc_tot, int n_rtiles,
                             long long n_q, int ns, int nstage, int seed_stride, int *cand_idx,
                             float *cand_thr, cudaStream_t st) {
    if (n_q <= 0) return cudaSuccess;
    if (ns == 2)
        return launch_tc<SK_TC_KCS, TC_MT, 2, 16>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q, cand_idx,
                                          cand_thr, st);
    if (ns == 1)
        return launch_tc<16, TC_MT, 1, 32>(qimg, rimg, kc_tot, n_rtiles, nstage, seed_stride, n_q, cand_idx,
                                           cand_thr, st);
    return cudaErrorInvalidValue;
}

}  // namespace sk
