// Device helpers of the tensor-core search kernel (search_tc.cu): tcgen05 / TMEM wrappers, register
// sorting networks, shared-memory accessors by 32-bit address and the rank-of-a-union helper of the
// joint thresholds.
#pragma once
#include "common.cuh"

namespace sk {

constexpr int TC_SLOTS = 4;                     // TMEM accumulator slots of TC_N columns
constexpr int TC_GROUPS = 32;                   // seeding: group minima per query
constexpr uint32_t TC_ROWB = 16;                // bytes of one row of one K chunk (8 FP16)
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
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
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

// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=F16 [7,10)=0,
// B=F16 [10,13)=0, both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
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


struct ThrPr {
    float thr;
    uint32_t pr;
    int n_it;   // tc_drain: the largest number of parked octets any lane of the warp had
};

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
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// predicated dump of one octet of scores (+ the index of its first reference) into the calling
// thread's next pending-queue entry: no branch, lanes whose predicate is off store nothing
__device__ __forceinline__ void tc_dump8(bool on, uint32_t va, uint32_t ia, const float *w, int id) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %0, 0;\n"
        "@p st.shared.v4.f32 [%1], {%3, %4, %5, %6};\n"
        "@p st.shared.v4.f32 [%1+16], {%7, %8, %9, %10};\n"
        "@p st.shared.b32 [%2], %11;\n"
        "}\n" ::"r"((uint32_t)on),
        "r"(va), "r"(ia), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(w[4]), "f"(w[5]), "f"(w[6]),
        "f"(w[7]), "r"(id)
        : "memory");
}


// J-th smallest score of the union of two ascending lists of KC scores each (1 <= J <= 2 KC): the
// minimum over the splits i + j = J (i from `a`, j from `b`) of max(a[i - 1], b[j - 1]).
template <int KC, int J, int NA>
__device__ __forceinline__ float tc_union_rank(const float (&a)[NA], const float (&b)[KC]) {
    static_assert(J >= 1 && J <= 2 * KC && NA >= KC, "rank out of range");
    float best = SK_INF_F;
#pragma unroll
    for (int i = 0; i <= KC; ++i) {
        const int j = J - i;
        if (j < 0 || j > KC) continue;
        float m;
        if (i == 0) m = b[j - 1];
        else if (j == 0) m = a[i - 1];
        else m = fmaxf(a[i - 1], b[j - 1]);
        best = fminf(best, m);
    }
    return best;
}


}  // namespace sk
