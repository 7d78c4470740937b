// tcgen05 / TMEM primitives (sm_100a) for the flow kernels' F x F contractions.
// fp32-grade accuracy comes from the 3xTF32 split: x = hi + lo with hi = rna_tf32(x);
// A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, accumulated in fp32 in tensor memory.
#pragma once
#include "gwtf_common.cuh"

namespace gwtf {

// ---------------------------------------------------------------------------------------------
// tensor memory management (one warp allocates / frees; column count power of two >= 32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one lane of the (converged) warp, chosen by the hardware: lets the compiler keep what the elected thread computes
// in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM <-> registers, 32 lanes x 32 bit x N columns: thread t of warp w owns lane 32*(w%4)+t
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t a, float* d) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a));
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t a, float* d) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t a, float* d) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(a));
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t a, float (&d)[N]) {
    static_assert(N % 8 == 0, "column count must be a multiple of 8");
    int c = 0;
#pragma unroll
    for (; c + 32 <= N; c += 32) tmem_ld32(a + c, d + c);
    if (N - c >= 16) { tmem_ld16(a + c, d + c); c += 16; }
    if (N - c >= 8) { tmem_ld8(a + c, d + c); }
}

__device__ __forceinline__ void tmem_st8(uint32_t a, const float* s) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a),
                 "r"(__float_as_uint(s[0])), "r"(__float_as_uint(s[1])), "r"(__float_as_uint(s[2])),
                 "r"(__float_as_uint(s[3])), "r"(__float_as_uint(s[4])), "r"(__float_as_uint(s[5])),
                 "r"(__float_as_uint(s[6])), "r"(__float_as_uint(s[7]))
                 : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t a, const float (&s)[N]) {
    static_assert(N % 8 == 0, "column count must be a multiple of 8");
#pragma unroll
    for (int c = 0; c < N; c += 8) tmem_st8(a + c, s + c);
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 split
// ---------------------------------------------------------------------------------------------
// hi = x rounded to tf32 (nearest, ties away: add half an ulp of the 10-bit mantissa, clear the 13 low
// bits) in two integer ops -- cvt.rna.tf32.f32 compiles to five (it also guards inf/nan, which the
// O(1) activations and weights of the flow nets never are); lo = x - hi is exact.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// ---------------------------------------------------------------------------------------------
// shared-memory operand layouts (no swizzle) and their descriptors; offsets are in floats
//   K-major  (rows = M or N index, K contiguous in 16-byte chunks):
//     core matrix = 8 rows x 16 B; K-chunks of a row group are adjacent (LBO = 128 B), row groups
//     follow at SBO = (Ktot/4) * 128 B
//   MN-major (rows = K index, M/N contiguous in 16-byte chunks):
//     core matrix = 8 k-rows x 16 B; k-groups adjacent (LBO = 128 B), MN chunks at SBO = (Ktot/8)*128 B
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int kmajor_offset(int row, int k, int Ktot) {
    return (row >> 3) * (Ktot / 4) * 32 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
__host__ __device__ constexpr int mnmajor_offset(int mn, int k, int Ktot) {
    return (mn >> 2) * (Ktot / 8) * 32 + (k >> 3) * 32 + (k & 7) * 4 + (mn & 3);
}
// MN-major tf32 operands only exist in the 128B-swizzle / 32B-base layout: atom = 4 k-rows x 128 B
// (32 tf32 along M/N), 32-byte chunk index XOR k-row; k-atoms adjacent (SBO = 512 B), MN atoms at
// LBO = (Ktot/4) * 512 B.  Offset in floats of element (mn, k); base must be 512-byte aligned.
__host__ __device__ constexpr int mnmajor_sw32_offset(int mn, int k, int Ktot) {
    return (mn >> 5) * (Ktot / 4) * 128 + (k >> 2) * 128 + (k & 3) * 32 + ((((mn >> 3) & 3) ^ (k & 3)) << 3) + (mn & 7);
}
__device__ __forceinline__ uint64_t make_smem_desc(const void* p, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)(layout_type & 7u) << 61;
    d |= (uint64_t)((smem_u32(p) >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;    // descriptor version (Blackwell)
    return d;           // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
__device__ __forceinline__ uint64_t make_smem_desc_kmajor(const void* p, int Ktot) {
    return make_smem_desc(p, 128u, (uint32_t)(Ktot / 4) * 128u);
}
__device__ __forceinline__ uint64_t make_smem_desc_mnmajor_sw32(const void* p, int Ktot) {
    return make_smem_desc(p, (uint32_t)(Ktot / 4) * 512u, 512u, 1u);     // SWIZZLE_128B_BASE32B
}
// instruction descriptor: D fp32, A/B tf32, M x N, major bits (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]     (issued by ONE thread)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

}  // namespace gwtf
