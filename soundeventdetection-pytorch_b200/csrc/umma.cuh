// Blackwell (sm_100a) primitives used by the SED kernels: tcgen05 MMA / TMEM, mbarrier,
// bulk async copies (TMA engine), proxy fences.  Hand-written inline PTX; no CUTLASS.
//
// Canonical shared-memory operand layouts (SWIZZLE_NONE, 16-bit elements, 8x8 "core matrices" of
// 128 contiguous bytes).  For an operand tile with MN rows/cols and a K extent:
//
//   K-major  : byte(mn,k) = (mn/8)*SBO + (k/8)*LBO + (mn%8)*16 + (k%8)*2      (8 K-elements contiguous)
//   MN-major : byte(mn,k) = (mn/8)*SBO + (k/8)*LBO + (k%8)*16  + (mn%8)*2     (8 MN-elements contiguous)
//
// i.e. in both modes SBO strides between groups of 8 along MN and LBO between groups of 8 along K.
// (Descriptor bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) = 0 for no swizzle.)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sedb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware (no issue slots consumed) until the phase
// completes or ~hint_ns elapsed, instead of polling.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// What a timed-out wait does.  A printf here is a real call (vprintf): ptxas then treats most registers as clobbered on
// the path that rejoins the wait's fast exit and re-derives thread ids, window bases and descriptors after EVERY wait
// (about a dozen instructions, two of them S2R, per fold chunk of the log-mel kernel).  The default is a bare trap;
// -DSEDB_MBAR_DEBUG=1 brings the message back.
#ifndef SEDB_MBAR_DEBUG
#define SEDB_MBAR_DEBUG 0
#endif
__device__ __forceinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
#if SEDB_MBAR_DEBUG
    printf("sedb: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
#else
    (void)bar;
    (void)parity;
#endif
    __trap();
}

// Bounded wait: a protocol bug must not hang the GPU box.  On timeout (~seconds) the kernel traps, which surfaces
// as a CUDA launch failure on the host instead of a wedged device.  The timeout clock is only consulted every 256
// failed (sleeping) probes so that waiting warps do not steal issue slots from working ones.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    long long t0 = 0;
#ifndef SEDB_MBAR_HINT_NS
#define SEDB_MBAR_HINT_NS 20000u
#endif
    while (!(SEDB_MBAR_HINT_NS ? mbar_try_wait_hint(bar, parity, SEDB_MBAR_HINT_NS) : mbar_try_wait(bar, parity))) {
        if ((++spins & 0xffu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 6000000000LL) {
                mbar_timeout(bar, parity);
            }
        }
    }
}

// One elected lane of a fully converged warp.  Code that issues tcgen05.mma / bulk copies should keep the whole warp
// converged and guard only the instruction with this predicate: operands then stay in uniform registers, whereas a
// lane-divergent `if (lane == 0)` region makes ptxas wrap every UTCHMMA in an R2UR + vote loop (~150 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// The same for a warp that can afford to react late (producers that run ahead, helpers with a frame of slack): sleeps
// between probes, so that the polling loop does not take issue slots from the working warps of its scheduler (a bare
// try_wait loop issues an instruction every ~10 cycles: measured ~10 % of a scheduler's slots per polling warp).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        if ((++spins & 0xffu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 6000000000LL) {
                mbar_timeout(bar, parity);
            }
        }
    }
}

// ------------------------------------------------------------------ programmatic dependent launch
// First statement of a kernel that is launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch_pdl in
// cnn_host.inl): lets the NEXT kernel of the stream be placed as this grid's CTAs retire, then waits until the PREVIOUS
// grid has completed and its memory operations are visible.  Both are no-ops for an ordinary launch.  Every thread of a
// kernel launched that way must pass through the wait before it touches global memory.
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------ proxy fences
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ bulk copy global -> smem (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ cp.async (LDGSTS): 16-byte global -> shared copies
// that do not pass through registers; a thread fires all of them and waits once.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ descriptors
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
    return d;                              // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1 };
enum : uint32_t { kMajorK = 0, kMajorMN = 1 };

// kind::f16 instruction descriptor: fp32 accumulate, A/B formats, majors, optional negation, M and N.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t a_major, uint32_t b_major, uint32_t M,
                                                  uint32_t N, uint32_t a_neg = 0, uint32_t b_neg = 0) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_neg << 13) | (b_neg << 14) | (a_major << 15) |
           (b_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same with the descriptors given as (lo, hi) words: only the start-address field (low 14 bits of lo) differs
// between the operands of a kernel, so the issuing thread does 32-bit adds on lo and keeps hi constant.  The single
// issuing thread is instruction-latency bound (a handful of dependent uniform-datapath instructions cost as much as a
// narrow MMA takes to run: tests/dev/umma_switch.py), so every instruction per MMA counts.
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 lanes, 16-bit elements, two consecutive K per 32-bit column,
// 8 columns per K = 16 step) is read from tensor memory, so it costs no shared-memory bandwidth.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 2 consecutive 32-bit columns, registers -> tensor memory (thread i writes lane base_lane + i)
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ split-precision helpers
// x ~= hi + lo with both halves in a 16-bit float format; three MMAs (hi*hi + lo*hi + hi*lo) then
// reproduce the fp32 product to ~2^-17 (bf16) / ~2^-22 (fp16) relative.
#ifndef SEDB_SPLIT_FP16
#define SEDB_SPLIT_FP16 1
#endif
#if SEDB_SPLIT_FP16
typedef __half split_t;
constexpr uint32_t kSplitFmt = kFmtF16;
__device__ __forceinline__ split_t to_split(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float from_split(split_t h) { return __half2float(h); }
#else
typedef __nv_bfloat16 split_t;
constexpr uint32_t kSplitFmt = kFmtBF16;
__device__ __forceinline__ split_t to_split(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float from_split(split_t h) { return __bfloat162float(h); }
#endif

__device__ __forceinline__ uint32_t pack2(split_t a, split_t b) {
    uint16_t ua = *reinterpret_cast<uint16_t*>(&a);
    uint16_t ub = *reinterpret_cast<uint16_t*>(&b);
    return static_cast<uint32_t>(ua) | (static_cast<uint32_t>(ub) << 16);
}
// Split two floats into packed hi and lo halves.  hi = x truncated to the 16-bit format's mantissa (a mask, exactly
// representable, so its packed conversion is exact); lo = rn16(x - hi).  hi + lo reproduces x to ~2^-21 (fp16) /
// ~2^-16 (bf16) relative.  Two packed conversions per pair instead of six scalar ones (conversions are a slow pipe).
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
#if SEDB_SPLIT_FP16
    const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
    const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
    const __half2 hh = __floats2half2_rn(h0, h1);
    const __half2 ll = __floats2half2_rn(x0 - h0, x1 - h1);
    hi = *reinterpret_cast<const uint32_t*>(&hh);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
#else
    const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
    hi = __byte_perm(u0, u1, 0x7632);                          // {x0[31:16], x1[31:16]}: bf16 truncation, no cvt
    const float h0 = __uint_as_float(u0 & 0xFFFF0000u), h1 = __uint_as_float(u1 & 0xFFFF0000u);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
#endif
}

// ---- packed fp32 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 results per issue slot) ---------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }   // folds into an operand modifier
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

// split_pack2 on a register pair (the lo subtraction is one packed instruction)
__device__ __forceinline__ void split_pack2(float2 x, uint32_t& hi, uint32_t& lo) {
#if SEDB_SPLIT_FP16
    const float2 h = make_float2(__uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u),
                                 __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
    const float2 l = f2sub(x, h);
    const __half2 hh = __floats2half2_rn(h.x, h.y);
    const __half2 ll = __floats2half2_rn(l.x, l.y);
    hi = *reinterpret_cast<const uint32_t*>(&hh);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
#else
    split_pack2(x.x, x.y, hi, lo);
#endif
}

}  // namespace sedb
