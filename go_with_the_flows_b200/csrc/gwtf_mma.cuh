// Warp-level tensor-core building blocks shared by the forward and backward layer kernels:
// mma.sync.m16n8k8 tf32 on register fragments with the 3xTF32 split (x = hi + lo, hi = rna_tf32(x);
// A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo) for fp32-grade results.
//
// Fragment coordinates of lane (g = lane>>2, t = lane&3):
//   A (16x8):  a0=(row g, k t)  a1=(row g+8, k t)  a2=(row g, k t+4)  a3=(row g+8, k t+4)
//   B (8x8):   b0=(k t, n g)    b1=(k t+4, n g)
//   C (16x8):  c0=(row g, n 2t) c1=(row g, n 2t+1) c2=(row g+8, n 2t) c3=(row g+8, n 2t+1)
// The k index of a contraction may be permuted freely, so operands are staged with A columns (t, t+4)
// standing for channels (2t, 2t+1): a C fragment then IS the next MMA's A fragment.
#pragma once
#include "gwtf_common.cuh"

namespace gwtf {

// hi = x rounded to tf32 (nearest, ties away) in two integer ops (cvt.rna.tf32.f32 compiles to five: it
// also guards inf/nan, which the flow nets' O(1) operands never are); lo = x - hi is exact.
__device__ __forceinline__ void split_tf32_bits(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_tf32p(float* d, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// row stride of the per-warp transpose tiles: == 8 (mod 16) makes the k-major fragment loads of MMA #3
// and the float2 stores of the C fragments conflict-free
__host__ __device__ constexpr int mma_tile_stride(int FP) { return (FP % 16 == 8) ? FP : FP + 8; }

// Kept-activation buffer of the tensor-core engines (the forward apply pass writes it, both backward
// phases read it).  One slab per (component, net), shapes padded to 256 points, stored in MMA C-fragment
// order: [m-tile of 16 points][n-tile][lane = 4g + t][4] floats = (row g: ch 2t, 2t+1 | row g+8: ch 2t, 2t+1)
// of channels 8nt + ..., so that a backward warp reads its m-tile with NT fully coalesced 16-byte loads
// per lane (the buffer is written once and read twice).  Measured on B200 (C2 step): this order 18.9 ms,
// point-major [ch/4][point][4] (cheaper for the one-thread-per-point tcgen05 writer) 19.4 ms, no buffer 19.3 ms.
__host__ __device__ inline size_t keep_npad(int N) { return (size_t)((N + 255) / 256) * 256; }
__host__ __device__ inline size_t keep_points(int B, int N) { return (size_t)B * keep_npad(N); }
// floats per (layer, component)
__host__ __device__ inline size_t mma_keep_floats(int F, int B, int N) {
    return 2 * keep_points(B, N) * (size_t)round_up(F, 8);
}
// slab base of (component j, net), and the address of (channel group f4, point p) inside it
__host__ __device__ inline size_t keep_slab(int F, int B, int N, int j, int net) {
    return ((size_t)j * 2 + net) * keep_points(B, N) * (size_t)round_up(F, 8);
}

// B fragments of h1 = a0 W1^T for one net, pre-split: per (ks, nt, lane) = (b0.hi, b1.hi, b0.lo, b1.lo)
// with k = e (A columns t, t+4 <-> e = 8ks+2t, 8ks+2t+1), n = f = 8nt+g.   W1T[e][f] = sd1.weight[f][e].
template <int FP>
__device__ __forceinline__ void stage_bfrag_h1(float4* bf, const float (&W1T)[FP][FP], int tid, int nthreads) {
    constexpr int KS = FP / 8, NT = FP / 8;
    for (int i = tid; i < KS * NT * 32; i += nthreads) {
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int gg = ln >> 2, tt = ln & 3;
        uint32_t h0, l0, h1, l1;
        split_tf32_bits(W1T[8 * ks + 2 * tt][8 * nt + gg], h0, l0);
        split_tf32_bits(W1T[8 * ks + 2 * tt + 1][8 * nt + gg], h1, l1);
        bf[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
}

// h[mi][nt][2r+i] = h1 of (m-tile mi, row g+8r, channel 8nt+2t+i) for the lane's points x[mi][r]; rows with
// valid false give 0.  The MI m-tiles share every q0 / B-fragment load (shared-memory traffic per MMA / MI)
// and give the scheduler MI independent MMA chains.
template <int FP, int MI>
__device__ __forceinline__ void mma_h1(const float4 (&q0)[FP], const float4* bf, const float (&x)[MI][2][3],
                                       const bool (&valid)[MI][2], int lane, float (&h)[MI][FP / 8][4]) {
    constexpr int KS = FP / 8, NT = FP / 8;
    const int t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) h[mi][nt][0] = h[mi][nt][1] = h[mi][nt][2] = h[mi][nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const float4 qa = q0[8 * ks + 2 * t];
        const float4 qb = q0[8 * ks + 2 * t + 1];
        uint32_t ah[MI][4], al[MI][4];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            float av[4];
            av[0] = fmaxf(fmaf(qa.x, x[mi][0][0], fmaf(qa.y, x[mi][0][1], fmaf(qa.z, x[mi][0][2], qa.w))), 0.f);
            av[1] = fmaxf(fmaf(qa.x, x[mi][1][0], fmaf(qa.y, x[mi][1][1], fmaf(qa.z, x[mi][1][2], qa.w))), 0.f);
            av[2] = fmaxf(fmaf(qb.x, x[mi][0][0], fmaf(qb.y, x[mi][0][1], fmaf(qb.z, x[mi][0][2], qb.w))), 0.f);
            av[3] = fmaxf(fmaf(qb.x, x[mi][1][0], fmaf(qb.y, x[mi][1][1], fmaf(qb.z, x[mi][1][2], qb.w))), 0.f);
            if (!valid[mi][0]) av[0] = av[2] = 0.f;
            if (!valid[mi][1]) av[1] = av[3] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32_bits(av[i], ah[mi][i], al[mi][i]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float4 bq = bf[(ks * NT + nt) * 32 + lane];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma_tf32(h[mi][nt], ah[mi], __float_as_uint(bq.x), __float_as_uint(bq.y));
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma_tf32(h[mi][nt], al[mi], __float_as_uint(bq.x), __float_as_uint(bq.y));
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) mma_tf32(h[mi][nt], ah[mi], __float_as_uint(bq.z), __float_as_uint(bq.w));
        }
    }
}
template <int FP>
__device__ __forceinline__ void mma_h1(const float4 (&q0)[FP], const float4* bf, const float (&x)[2][3],
                                       const bool (&valid)[2], int lane, float (&h)[FP / 8][4]) {
    mma_h1<FP, 1>(q0, bf, reinterpret_cast<const float(&)[1][2][3]>(x), reinterpret_cast<const bool(&)[1][2]>(valid), lane,
                  reinterpret_cast<float(&)[1][FP / 8][4]>(h));
}

// C fragments of one m-tile <-> kept buffer.  `slab` = keep + keep_slab(...), p0 = padded point index of the
// m-tile's first row (b * npad + n0, a multiple of 16).
template <int NT>
__device__ __forceinline__ void load_h1_frag(const float* slab, size_t p0, int lane, float (&h)[NT][4]) {
    const float4* src = reinterpret_cast<const float4*>(slab) + (p0 >> 4) * (NT * 32) + lane;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float4 v = src[nt * 32];
        h[nt][0] = v.x; h[nt][1] = v.y; h[nt][2] = v.z; h[nt][3] = v.w;
    }
}
template <int NT>
__device__ __forceinline__ void store_h1_frag(float* slab, size_t p0, int lane, const float (&h)[NT][4]) {
    float4* dst = reinterpret_cast<float4*>(slab) + (p0 >> 4) * (NT * 32) + lane;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) dst[nt * 32] = make_float4(h[nt][0], h[nt][1], h[nt][2], h[nt][3]);
}
// the same element addressed from a one-thread-per-point kernel: channel pair (8nt + 2t, +1) of padded
// point p is the float2 at keep_point_base(p, NT) + (nt * 32 + t) * 4
__host__ __device__ inline size_t keep_point_base(size_t p, int NT) {
    return ((p >> 4) * (size_t)(NT * 32) + (p & 7) * 4) * 4 + ((p >> 3) & 1) * 2;
}

}  // namespace gwtf
