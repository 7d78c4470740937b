// Backward phase 1 (sd1 / sd0 / bn0 gradients + input gradient) with every contraction on the tensor
// cores through warp-level register fragments (mma.sync.m16n8k8 tf32, 3xTF32 split = fp32-grade).
//
// Replaces the CUDA-core contractions of k_bwd_layer_e (gwtf_bwd.cuh); same BwdArgs, same outputs.
// Reference semantics: lib/networks/flow.py (CondRealNVPFlow3D.forward, autograd of the F->F block).
//
// One warp owns tiles of 16 points (the MMA M dimension) and runs the whole chain on them without any
// CTA-level synchronisation:
//   a0  = relu(q0 x)                       computed straight into the A-fragment layout
//   h1  = a0 W1^T                          MMA #1  (B fragments pre-split hi/lo in shared memory)
//   dh1 = g(h1, dO)                        element-wise on the C fragments
//   da0 = dh1 W1                           MMA #2  (C fragment -> A fragment is free: the k index of a
//                                          contraction may be permuted, so B is staged with k = 2t, 2t+1)
//   dy0 = da0 * [y0 > 0]  -> per-channel sums (registers, reduced once per CTA) and the input gradient
//   dW1 += dh1^T a0                        MMA #3  (k = points: the two tiles take a trip through the
//                                          warp's private shared-memory tile to transpose)
#pragma once
#include "gwtf_bwd.cuh"
#include "gwtf_tc.cuh"
#include "gwtf_mma.cuh"

namespace gwtf {

template <int FP>
struct BwdEMmaSmem {
    LayerW<FP> W;
    LayerWB<FP> WB;
    uint64_t bar;
    uint32_t tmem_base;
    float4 cf1[FP];         // current (net, shape): (st.x, st.y, A1, A0)   dh1 = [y1>0] w2s.dO + h*A1 + A0
    float4 cf2[FP];         //                       w2s = sd2 columns * s * istd1
    float red[round_up(5 * FP, 32)];
};

// Tensor-memory budget of phase 1: the two persistent accumulator sets of a warp (dW1 fragments, per-e
// sums) live in the warp's own TMEM lanes between tiles -- tensor memory as a register file extension,
// which is what lets the kernel run at 128 registers and two CTAs per SM.
template <int FP>
struct BwdETmem {
    static constexpr int MT = (FP + 15) / 16, NT = FP / 8;
    static constexpr int GC = round_up(MT * NT * 4, 8);      // dW1 accumulator columns per thread
    static constexpr int EC = round_up(NT * 10, 8);          // per-e sums
    static constexpr int per_warp = GC + EC;
    static constexpr int need = (kThreads / 128) * per_warp; // warps w and w+4 share a lane quadrant
    static constexpr int alloc = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    static_assert(need <= 512, "tensor memory budget");
};

template <int FP>
__host__ __device__ constexpr size_t bwd_e_mma_smem(int F) {
    return round_up((int)sizeof(BwdEMmaSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 +
           2 * (size_t)(FP / 8) * (FP / 8) * 32 * 16 +                       // B fragments of MMA #1, #2 (one net)
           (size_t)(kThreads / 32) * 2 * 16 * mma_tile_stride(FP) * 4;       // per-warp dh1 / a0 tiles
}

template <int FP>
__global__ void __launch_bounds__(kThreads, 2) k_bwd_layer_e_mma(const BwdArgs a) {
    static_assert(FP % 8 == 0, "feature width padded to the MMA K");
    constexpr int KS = FP / 8, NT = FP / 8, MT = (FP + 15) / 16;
    constexpr int FPS = mma_tile_stride(FP);
    constexpr int NW = kThreads / 32;
    constexpr int TILE = NW * 16;                            // points per CTA tile
    constexpr int GC = BwdETmem<FP>::GC, EC = BwdETmem<FP>::EC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdEMmaSmem<FP>& S = *reinterpret_cast<BwdEMmaSmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(BwdEMmaSmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    float4* bf1 = reinterpret_cast<float4*>(raw + round_up(raw_floats(F), 4));    // [KS][NT][32]
    float4* bf2 = bf1 + KS * NT * 32;
    float* tiles = reinterpret_cast<float*>(bf2 + KS * NT * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    float* Dw = tiles + (size_t)warp * 2 * 16 * FPS;         // [16][FPS] dh1 of the warp's tile
    float* Aw = Dw + 16 * FPS;                               // [16][FPS] a0
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (tid == 0) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&S.tmem_base, BwdETmem<FP>::alloc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();        // only after this CTA owns its tensor-memory columns (dependents allocate too)
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    const uint32_t t_g = S.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * (GC + EC);
    const uint32_t t_e = t_g + GC;
    mbar_wait(&S.bar, 0u);
    stage_w1t<FP>(S.W, raw, F, a.d.warp_mask[l], tid, kThreads);     // parameters only: may overlap phase 0's tail
    pdl_wait();                                                       // phase 0 of this layer is complete from here on
    stage_vectors<FP, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false,
                            train ? a.bsum + (size_t)j * 8 * F : nullptr, tid, kThreads);

    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm), k = 3 - w;
    const NetOffsets o = net_offsets(F, w);
    int col_of_dim[3];
    { int q = 0; for (int dd = 0; dd < 3; ++dd) col_of_dim[dd] = (wm & (1u << dd)) ? -1 : q++; }
    float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
    double* bs = a.bsum + (size_t)j * 8 * F;

    const int tps = (N + TILE - 1) / TILE;
    const long long total_tiles = (long long)B * tps;
    const int t_begin = (int)(total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x);
    const size_t npad = keep_npad(N);

#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        __syncthreads();                                     // W ready / previous net done with bf*, tiles
        // ---- B fragments of this net, pre-split: (b0.hi, b1.hi, b0.lo, b1.lo) per lane
        for (int i = tid; i < KS * NT * 32; i += kThreads) {
            const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
            const int gg = ln >> 2, tt = ln & 3;
            uint32_t h0, l0, h1, l1;
            // MMA #1: k = e (A columns t, t+4 <-> e = 8ks+2t, 8ks+2t+1), n = f = 8nt+g  (not needed when h1 is kept)
            if (!a.y1in) {
                split_tf32_bits(S.W.W1T[net][8 * ks + 2 * tt][8 * nt + gg], h0, l0);
                split_tf32_bits(S.W.W1T[net][8 * ks + 2 * tt + 1][8 * nt + gg], h1, l1);
                bf1[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
            }
            // MMA #2: k = f (8ks+2t, 8ks+2t+1), n = e = 8nt+g
            split_tf32_bits(S.W.W1T[net][8 * nt + gg][8 * ks + 2 * tt], h0, l0);
            split_tf32_bits(S.W.W1T[net][8 * nt + gg][8 * ks + 2 * tt + 1], h1, l1);
            bf2[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
        }
        {   // zero the tensor-memory accumulators of this warp
            float z[GC > EC ? GC : EC];
#pragma unroll
            for (int i = 0; i < (GC > EC ? GC : EC); ++i) z[i] = 0.f;
            tmem_st<GC>(t_g, reinterpret_cast<const float(&)[GC]>(z));
            tmem_st<EC>(t_e, reinterpret_cast<const float(&)[EC]>(z));
            tmem_wait_st();
        }

        int cur_b = -1;
#pragma unroll 1
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int b = tile / tps;
            const int tin = tile - b * tps;
            const int n0 = tin * TILE + warp * 16;
            if (b != cur_b) {
                __syncthreads();
                if (tid < FP) {
                    const int f = tid;
                    float4 c1 = make_float4(0.f, 0.f, 0.f, 0.f), c2 = c1;
                    if (f < F) {
                        const float* film = a.film + ((size_t)(b * K + j) * L + l) * 4 * F;
                        const float s = film[net * 2 * F + f], tt = film[net * 2 * F + F + f];
                        const float2 mi = S.W.mi1[net][f];
                        const float2 ab = S.WB.ab1[net][f];
                        const float4 w2 = S.W.w2[net][f];
                        const float sc = s * mi.y;
                        if (a.kept_y1)      // the kept value is y1 itself: n1 = (y1 - t) / s
                            c1 = make_float4(1.f, 0.f, -mi.y * ab.y / s, mi.y * (ab.y * tt / s - ab.x));
                        else
                            c1 = make_float4(sc, tt - s * mi.x, -mi.y * mi.y * ab.y, mi.y * (mi.x * ab.y - ab.x));
                        c2 = make_float4(w2.x * sc, w2.y * sc, w2.z * sc, 0.f);
                    }
                    S.cf1[f] = c1;
                    S.cf2[f] = c2;
                }
                __syncthreads();
                cur_b = b;
            }
            // ---- this lane's two rows (points) of the tile
            const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
            const size_t sb = ((size_t)j * B + b) * 3 * N;
            const float* dob = a.dobuf + ((size_t)j * B + b) * 6 * N + (size_t)net * 3 * N;
            float x[2][3], dO[2][3], gold[2];
            bool valid[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int n = n0 + g + 8 * r;
                valid[r] = n < N;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    x[r][d] = valid[r] ? xin[(size_t)d * N + n] : 0.f;
                    dO[r][d] = valid[r] ? dob[(size_t)d * N + n] : 0.f;
                }
                gold[r] = (valid[r] && t < 3) ? a.gbuf[sb + (size_t)t * N + n] : 0.f;
            }
            const bool ragged = n0 + 16 > N;                  // some rows of this m-tile lie beyond the shape
            // ---- a0 (A fragments) and MMA #1: h1[row][f]  (or h1 kept by the forward pass)
            float h[NT][4];
            const bool kept = a.y1in != nullptr;
            if (kept) {
                load_h1_frag<NT>(a.y1in + keep_slab(F, B, N, j, net), (size_t)b * npad + n0, lane, h);
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) h[nt][0] = h[nt][1] = h[nt][2] = h[nt][3] = 0.f;
            }
            uint32_t mask0 = 0u, mask1 = 0u;                  // [y0 > 0] of rows g, g+8; bit = 2ks+i
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float4 qa = S.W.q0[net][8 * ks + 2 * t];
                const float4 qb = S.W.q0[net][8 * ks + 2 * t + 1];
                float av[4];
                av[0] = fmaf(qa.x, x[0][0], fmaf(qa.y, x[0][1], fmaf(qa.z, x[0][2], qa.w)));   // (row g,   e0)
                av[1] = fmaf(qa.x, x[1][0], fmaf(qa.y, x[1][1], fmaf(qa.z, x[1][2], qa.w)));   // (row g+8, e0)
                av[2] = fmaf(qb.x, x[0][0], fmaf(qb.y, x[0][1], fmaf(qb.z, x[0][2], qb.w)));   // (row g,   e1)
                av[3] = fmaf(qb.x, x[1][0], fmaf(qb.y, x[1][1], fmaf(qb.z, x[1][2], qb.w)));   // (row g+8, e1)
                mask0 |= (av[0] > 0.f ? 1u : 0u) << (2 * ks) | (av[2] > 0.f ? 2u : 0u) << (2 * ks);
                mask1 |= (av[1] > 0.f ? 1u : 0u) << (2 * ks) | (av[3] > 0.f ? 2u : 0u) << (2 * ks);
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = fmaxf(av[i], 0.f);
                if (ragged) {                                   // warp-uniform: only the last tile of a shape
                    if (!valid[0]) av[0] = av[2] = 0.f;
                    if (!valid[1]) av[1] = av[3] = 0.f;
                }
                *reinterpret_cast<float2*>(Aw + g * FPS + 8 * ks + 2 * t) = make_float2(av[0], av[2]);
                *reinterpret_cast<float2*>(Aw + (g + 8) * FPS + 8 * ks + 2 * t) = make_float2(av[1], av[3]);
                if (!kept) {
                    uint32_t ah[4], al[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split_tf32_bits(av[i], ah[i], al[i]);
                    float4 bq[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) bq[nt] = bf1[(ks * NT + nt) * 32 + lane];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(h[nt], ah, __float_as_uint(bq[nt].x), __float_as_uint(bq[nt].y));
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(h[nt], al, __float_as_uint(bq[nt].x), __float_as_uint(bq[nt].y));
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) mma_tf32(h[nt], ah, __float_as_uint(bq[nt].z), __float_as_uint(bq[nt].w));
                }
            }
            // ---- h1 -> dh1 on the C fragments: h[nt][2r+i] = (row g+8r, f = 8nt+2t+i)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float4 c1 = S.cf1[8 * nt + 2 * t + i];
                    const float4 c2 = S.cf2[8 * nt + 2 * t + i];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float hv = h[nt][2 * r + i];
                        const float y1 = fmaf(c1.x, hv, c1.y);
                        const float da = fmaf(c2.x, dO[r][0], fmaf(c2.y, dO[r][1], c2.z * dO[r][2]));
                        const float dh = (y1 > 0.f ? da : 0.f) + fmaf(hv, c1.z, c1.w);
                        h[nt][2 * r + i] = (ragged && !valid[r]) ? 0.f : dh;
                    }
                }
                *reinterpret_cast<float2*>(Dw + g * FPS + 8 * nt + 2 * t) = make_float2(h[nt][0], h[nt][1]);
                *reinterpret_cast<float2*>(Dw + (g + 8) * FPS + 8 * nt + 2 * t) = make_float2(h[nt][2], h[nt][3]);
            }
            // ---- MMA #2: da0[row][e] = sum_f dh1[row][f] W1[f][e]
            float da0[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) da0[nt][0] = da0[nt][1] = da0[nt][2] = da0[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t ah[4], al[4];
                split_tf32_bits(h[ks][0], ah[0], al[0]);      // (row g,   f = 8ks+2t)
                split_tf32_bits(h[ks][2], ah[1], al[1]);      // (row g+8, f = 8ks+2t)
                split_tf32_bits(h[ks][1], ah[2], al[2]);      // (row g,   f = 8ks+2t+1)
                split_tf32_bits(h[ks][3], ah[3], al[3]);      // (row g+8, f = 8ks+2t+1)
                float4 bq[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bq[nt] = bf2[(ks * NT + nt) * 32 + lane];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_tf32(da0[nt], ah, __float_as_uint(bq[nt].x), __float_as_uint(bq[nt].y));
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_tf32(da0[nt], al, __float_as_uint(bq[nt].x), __float_as_uint(bq[nt].y));
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_tf32(da0[nt], ah, __float_as_uint(bq[nt].z), __float_as_uint(bq[nt].w));
            }
            // ---- relu/bn0/sd0 pieces: da0[nt][2r+i] = (row g+8r, e = 8nt+2t+i); the running per-e sums
            // come from tensor memory and go back there
            {
                float es[EC];                                 // [(nt*2+i)*5+c]
                tmem_wait_st();
                tmem_ld<EC>(t_e, es);
                float vin[2][3];
#pragma unroll
                for (int r = 0; r < 2; ++r) vin[r][0] = vin[r][1] = vin[r][2] = 0.f;
                tmem_wait_ld();
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float4 qv = S.W.q0[net][8 * nt + 2 * t + i];
                        const float4 rv = S.WB.r0[net][8 * nt + 2 * t + i];
                        float* e5 = es + (nt * 2 + i) * 5;
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const uint32_t m = r == 0 ? mask0 : mask1;
                            const float dy0 = ((m >> (2 * nt + i)) & 1u) ? da0[nt][2 * r + i] : 0.f;
                            const float hh = fmaf(rv.x, x[r][0], fmaf(rv.y, x[r][1], fmaf(rv.z, x[r][2], rv.w)));
                            vin[r][0] = fmaf(qv.x, dy0, vin[r][0]);
                            vin[r][1] = fmaf(qv.y, dy0, vin[r][1]);
                            vin[r][2] = fmaf(qv.z, dy0, vin[r][2]);
                            e5[0] = fmaf(dy0, hh, e5[0]);          // d gamma0
                            e5[1] += dy0;                          // d beta0
                            e5[2] = fmaf(dy0, x[r][0], e5[2]);     // raw sd0 weight sums
                            e5[3] = fmaf(dy0, x[r][1], e5[3]);
                            e5[4] = fmaf(dy0, x[r][2], e5[4]);
                        }
                    }
                }
                tmem_st<EC>(t_e, es);
                // input gradient: sum over the 4 lanes that share a row, lane t < 3 writes dimension t
#pragma unroll
                for (int r = 0; r < 2; ++r) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        vin[r][d] += __shfl_xor_sync(0xffffffffu, vin[r][d], 1);
                        vin[r][d] += __shfl_xor_sync(0xffffffffu, vin[r][d], 2);
                    }
                    const float v = t == 0 ? vin[r][0] : (t == 1 ? vin[r][1] : vin[r][2]);
                    const int n = n0 + g + 8 * r;
                    if (valid[r] && t < 3) a.gbuf[sb + (size_t)t * N + n] = gold[r] + v;
                }
            }
            __syncwarp();
            // ---- MMA #3: dW1[f][e] += sum_rows dh1[row][f] a0[row][e]   (m = f, n = e, k = row);
            // accumulator fragments [(mt*NT+nt)*4+i] round-trip through tensor memory
            {
                float gacc[GC];
                tmem_ld<GC>(t_g, gacc);
                tmem_wait_ld();
#pragma unroll
                for (int kk = 0; kk < 16; kk += 8) {
                    const float* dr0 = Dw + (kk + t) * FPS;
                    const float* dr1 = dr0 + 4 * FPS;
                    const float* ar0 = Aw + (kk + t) * FPS;
                    const float* ar1 = ar0 + 4 * FPS;
                    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        split_tf32_bits(ar0[nt * 8 + g], bh[nt][0], bl[nt][0]);
                        split_tf32_bits(ar1[nt * 8 + g], bh[nt][1], bl[nt][1]);
                    }
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        constexpr bool kHalfLast = (FP % 16) != 0;   // last m-tile has only 8 real rows
                        const bool upper = !(kHalfLast && mt == MT - 1);
                        uint32_t ah[4], al[4];
                        split_tf32_bits(dr0[mt * 16 + g], ah[0], al[0]);
                        split_tf32_bits(dr1[mt * 16 + g], ah[2], al[2]);
                        if (upper) {
                            split_tf32_bits(dr0[mt * 16 + g + 8], ah[1], al[1]);
                            split_tf32_bits(dr1[mt * 16 + g + 8], ah[3], al[3]);
                        } else {
                            ah[1] = al[1] = ah[3] = al[3] = 0u;
                        }
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_tf32p(gacc + (mt * NT + nt) * 4, ah, bh[nt][0], bh[nt][1]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_tf32p(gacc + (mt * NT + nt) * 4, al, bh[nt][0], bh[nt][1]);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) mma_tf32p(gacc + (mt * NT + nt) * 4, ah, bl[nt][0], bl[nt][1]);
                    }
                }
                tmem_st<GC>(t_g, gacc);
            }
            __syncwarp();
        }

        // ---- flush the per-e sums of this net: lanes with the same t hold the same channels
        tmem_wait_st();
        __syncthreads();
        for (int i = tid; i < round_up(5 * FP, 32); i += kThreads) S.red[i] = 0.f;
        __syncthreads();
        {
            float es[EC];
            tmem_ld<EC>(t_e, es);
            tmem_wait_ld();
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        float v = es[(nt * 2 + i) * 5 + c];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (g == 0) atomicAdd(&S.red[(8 * nt + 2 * t + i) * 5 + c], v);
                    }
        }
        __syncthreads();
        for (int i = tid; i < 5 * FP; i += kThreads) {
            const int e = i / 5, c = i - 5 * e;
            if (e >= F) continue;
            const float val = S.red[i];
            if (c == 0) {
                atomicAdd(&dpr[net * o.stride + o.g0 + e], val);
                atomicAdd(&bs[(net * 4 + 3) * F + e], (double)val);
            } else if (c == 1) {
                atomicAdd(&dpr[net * o.stride + o.b0 + e], val);
                atomicAdd(&bs[(net * 4 + 2) * F + e], (double)val);
            } else {
                const int col = col_of_dim[c - 2];
                if (col >= 0) atomicAdd(&dpr[net * o.stride + o.W0 + e * k + col], val);
            }
        }
        // ---- flush dW1: per-warp partial outputs through shared memory (reuses the tile area), four
        // warps at a time
        constexpr int PM = MT * 16, PN = NT * 8;
        static_assert((size_t)4 * PM * PN <= (size_t)NW * 2 * 16 * FPS, "dW1 staging fits the tile area");
        float* part = tiles;                                  // [4][PM][PN]
        float gacc[GC];
        tmem_ld<GC>(t_g, gacc);
        tmem_wait_ld();
#pragma unroll 1
        for (int half = 0; half < NW / 4; ++half) {
            __syncthreads();
            if ((warp >> 2) == half) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int f = mt * 16 + g + ((i & 2) ? 8 : 0), e = nt * 8 + 2 * t + (i & 1);
                            part[((size_t)(warp & 3) * PM + f) * PN + e] = gacc[(mt * NT + nt) * 4 + i];
                        }
            }
            __syncthreads();
            for (int i = tid; i < F * F; i += kThreads) {
                const int f = i / F, e = i - f * F;
                float s = 0.f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) s += part[((size_t)qq * PM + f) * PN + e];
                atomicAdd(&dpr[net * o.stride + o.W1 + f * F + e], s);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(S.tmem_base, BwdETmem<FP>::alloc);
}

// =============================================================================================
// Backward phase 0 on register fragments: d(o_mu, o_lv), FiLM + sd2 gradients, sd1_bn backward sums.
// Same outputs as k_bwd_layer_d.  Two passes over the CTA's tiles: the logvar net first (its output
// fixes sigma, hence d o_mu and d o_lv, written to dobuf), then the mu net.  The 5 sums per channel
// live in registers in the C-fragment layout (10 channels per lane) and are reduced over the 8 row
// groups only when the shape changes (FiLM gradients are per shape) -- no per-tile reduce-scatter.
// =============================================================================================
template <int FP>
__host__ __device__ constexpr size_t bwd_d_mma_smem(int F) {
    return round_up((int)sizeof(BwdSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 +
           (size_t)(FP / 8) * (FP / 8) * 32 * 16;
}

// Tensor-memory budget of phase 0: the per-channel sums of a warp live in its own TMEM lanes, in chunks of
// two n-tiles (20 sums in 24 columns); the sd2 bias sums ride in the spare columns of the last chunk.  A tile
// brings one chunk at a time into registers, so the kernel fits 80 registers and three CTAs (24 warps) per SM
// (measured: 125 us per launch against 129 us with register accumulators at two CTAs; prefetching the next tile's
// kept activations into registers at two CTAs is slower, 136 us).
template <int FP>
struct BwdDTmem {
    static constexpr int NT = FP / 8;
    static constexpr int NCH = NT / 2 + 1;                    // full chunks + the tail chunk (odd n-tile and/or bias sums)
    static constexpr int TAIL = (NT & 1) ? 16 : 8;            // tail chunk: [10 sums of the odd n-tile] + 3 bias sums
    static constexpr int DC = 24 * (NT / 2) + TAIL;           // columns per warp
    static constexpr int need = (kThreads / 128) * DC;
    static constexpr int alloc = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    static constexpr int ctas_per_sm = 512 / alloc < 3 ? 512 / alloc : 3;
};

template <int FP>
__global__ void __launch_bounds__(kThreads, BwdDTmem<FP>::ctas_per_sm) k_bwd_layer_d_mma(const BwdArgs a) {
    static_assert(FP % 8 == 0, "feature width padded to the MMA K");
    constexpr int KS = FP / 8, NT = FP / 8;
    constexpr int NW = kThreads / 32;
    constexpr int TILE = NW * 16;
    constexpr int NV = 5 * FP + 3;
    using TM = BwdDTmem<FP>;
    constexpr int TAILB = (NT & 1) ? 10 : 0;                 // offset of the bias sums inside the tail chunk
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<FP>& S = *reinterpret_cast<BwdSmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(BwdSmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    float4* bf1 = reinterpret_cast<float4*>(raw + round_up(raw_floats(F), 4));    // [KS][NT][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int td = t < 3 ? t : 2;                            // the xyz dimension this lane finishes
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (tid == 0) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&S.tmem_base, TM::alloc);
    if (tid < 12) S.corr[tid] = 0.f;
    for (int i = tid; i < 2 * round_up(NV, 32); i += kThreads) (&S.red[0][0])[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();        // only after this CTA owns its tensor-memory columns (dependents allocate too)
    const uint32_t t_acc = S.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * TM::DC;
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    mbar_wait(&S.bar, 0u);
    if (!a.y1in) stage_w1t<FP>(S.W, raw, F, a.d.warp_mask[l], tid, kThreads);   // parameters only
    pdl_wait();                                               // the previous layer's phase 1 is complete from here on
    const bool correct = train && a.mom_prev != nullptr;
    if (correct)
        bn0_correction<FP>(a.d, a.params, j, l - 1, a.mom_prev + j * GWTF_MOM_STRIDE, a.bsum_prev + (size_t)j * 8 * F,
                           a.n_total, S.corr, tid);
    stage_vectors<FP, true>(S.W, &S.WB, raw, src, F, a.d.warp_mask[l], train, false, nullptr, tid, kThreads);
    __syncthreads();
    float Mrow[3], ccd;
#pragma unroll
    for (int i = 0; i < 3; ++i) Mrow[i] = correct ? S.corr[td * 3 + i] : 0.f;
    ccd = correct ? S.corr[9 + td] : 0.f;

    const unsigned wm = a.d.warp_mask[l];
    const int w = popc3(wm);
    const NetOffsets o = net_offsets(F, w);
    int row_of_dim[3];
    { int q = 0; for (int dd = 0; dd < 3; ++dd) row_of_dim[dd] = (wm & (1u << dd)) ? q++ : -1; }

    const int tps = (N + TILE - 1) / TILE;
    const long long total_tiles = (long long)B * tps;
    const int t_begin = (int)(total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x);
    const size_t npad = keep_npad(N);

#pragma unroll 1
    for (int net = 1; net >= 0; --net) {
        __syncthreads();
        if (!a.y1in) stage_bfrag_h1<FP>(bf1, S.W.W1T[net], tid, kThreads);
        // zero this warp's tensor-memory sums: (ds, dt, dW2 xyz) of f = 8nt+2t+i over this lane's rows
        {
            float z[24];
#pragma unroll
            for (int i = 0; i < 24; ++i) z[i] = 0.f;
#pragma unroll
            for (int c = 0; c < NT / 2; ++c) tmem_st<24>(t_acc + 24 * c, z);
            tmem_st<TM::TAIL>(t_acc + 24 * (NT / 2), reinterpret_cast<const float(&)[TM::TAIL]>(z));
            tmem_wait_st();
        }

        auto flush = [&](int b) {
            // tensor memory -> block partials (reduce over the 8 row groups) -> global; zero the sums
            tmem_wait_st();
#pragma unroll
            for (int c = 0; c < TM::NCH; ++c) {
                constexpr int W = 24;
                float v[W];
                const bool tail = c == NT / 2;
                if (tail) tmem_ld<TM::TAIL>(t_acc + 24 * c, reinterpret_cast<float(&)[TM::TAIL]>(v));
                else tmem_ld<24>(t_acc + 24 * c, v);
                tmem_wait_ld();
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const int nt = 2 * c + nn;
                    if (nt < NT && !(tail && nn == 1)) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int k5 = 0; k5 < 5; ++k5) {
                                float s = v[nn * 10 + i * 5 + k5];
                                s += __shfl_xor_sync(0xffffffffu, s, 4);
                                s += __shfl_xor_sync(0xffffffffu, s, 8);
                                s += __shfl_xor_sync(0xffffffffu, s, 16);
                                if (g == 0) atomicAdd(&S.red[net][(8 * nt + 2 * t + i) * 5 + k5], s);
                            }
                    }
                }
                if (tail) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        float s = v[TAILB + d];
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 8);
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (lane == 0) atomicAdd(&S.red[net][5 * FP + d], s);
                    }
                }
#pragma unroll
                for (int i = 0; i < W; ++i) v[i] = 0.f;
                if (tail) tmem_st<TM::TAIL>(t_acc + 24 * c, reinterpret_cast<const float(&)[TM::TAIL]>(v));
                else tmem_st<24>(t_acc + 24 * c, v);
            }
            tmem_wait_st();
            __syncthreads();
            float* dfl = a.dfilm + ((size_t)(b * K + j) * L + l) * 4 * F;
            float* dpr = a.dparams + (size_t)(j * L + l) * a.d.rec_stride;
            double* bs = a.bsum + (size_t)j * 8 * F;
            for (int idx = tid; idx < NV; idx += kThreads) {
                const float val = S.red[net][idx];
                if (idx < 5 * FP) {
                    const int f = idx / 5, c = idx - 5 * f;
                    if (f < F) {
                        if (c == 0) {            // ds
                            atomicAdd(&dfl[net * 2 * F + f], val);
                            if (train) atomicAdd(&bs[(net * 4 + 1) * F + f], (double)(val * S.WB.sg[net][f].x));
                        } else if (c == 1) {     // dt
                            atomicAdd(&dfl[net * 2 * F + F + f], val);
                            if (train) atomicAdd(&bs[(net * 4 + 0) * F + f], (double)(val * S.WB.sg[net][f].x));
                        } else {
                            const int row = row_of_dim[c - 2];
                            if (row >= 0) atomicAdd(&dpr[net * o.stride + o.W2 + row * F + f], val);
                        }
                    }
                } else {
                    const int row = row_of_dim[idx - 5 * FP];
                    if (row >= 0) atomicAdd(&dpr[net * o.stride + o.b2 + row], val);
                }
            }
            __syncthreads();
            for (int i = tid; i < round_up(NV, 32); i += kThreads) S.red[net][i] = 0.f;
            __syncthreads();
        };

        int cur_b = -1;
#pragma unroll 1
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int b = tile / tps;
            const int tin = tile - b * tps;
            const int n0 = tin * TILE + warp * 16;
            if (b != cur_b) {
                if (cur_b >= 0) flush(cur_b);
                else __syncthreads();
                const float* film = a.film + ((size_t)(b * K + j) * L + l) * 4 * F;
                stage_film<FP, true>(S.W, &S.WB, film, F, tid, kThreads);
                __syncthreads();
                // per-shape view of the value v the tile loop works on: v = h1 (y1 = st.x v + st.y,
                // n1 = v istd - mean istd), or v = kept y1 (y1 = v, n1 = (v - t) / s)
                for (int i = tid; i < 2 * FP; i += kThreads) {
                    const int nn = i / FP, c = i - nn * FP;
                    float2 mf = S.W.mi1[nn][c];
                    if (a.kept_y1) {
                        const float s = c < F ? film[nn * 2 * F + c] : 0.f, tt = c < F ? film[nn * 2 * F + F + c] : 0.f;
                        mf = c < F ? make_float2(tt / s, 1.0f / s) : make_float2(0.f, 0.f);
                        S.W.st[nn][c] = c < F ? make_float2(1.f, 0.f) : make_float2(0.f, 0.f);
                    }
                    S.mif[nn][c] = mf;
                }
                __syncthreads();
                cur_b = b;
            }
            const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
            const size_t sb = ((size_t)j * B + b) * 3 * N;
            float* dob = a.dobuf + ((size_t)j * B + b) * 6 * N;
            bool valid[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) valid[r] = n0 + g + 8 * r < N;
            float outd[2] = {0.f, 0.f}, gcd[2] = {0.f, 0.f}, gsd[2] = {0.f, 0.f};   // dimension td of this lane
            float dO[2][3];
            if (net == 1) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int n = n0 + g + 8 * r;
                    if (valid[r]) {
                        const float o0 = a.xout[sb + n], o1 = a.xout[sb + (size_t)N + n], o2 = a.xout[sb + 2 * (size_t)N + n];
                        outd[r] = td == 0 ? o0 : (td == 1 ? o1 : o2);
                        gcd[r] = a.gbuf[sb + (size_t)td * N + n] - (Mrow[0] * o0 + Mrow[1] * o1 + Mrow[2] * o2) + ccd;
                        gsd[r] = a.gs[sb + (size_t)td * N + n];
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int n = n0 + g + 8 * r;
#pragma unroll
                    for (int d = 0; d < 3; ++d) dO[r][d] = valid[r] ? dob[(size_t)d * N + n] : 0.f;
                }
            }
            float h[NT][4];
            if (a.y1in) {
                load_h1_frag<NT>(a.y1in + keep_slab(F, B, N, j, net), (size_t)b * npad + n0, lane, h);
            } else {
                float x[2][3];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int n = n0 + g + 8 * r;
#pragma unroll
                    for (int d = 0; d < 3; ++d) x[r][d] = valid[r] ? xin[(size_t)d * N + n] : 0.f;
                }
                mma_h1<FP>(S.W.q0[net], bf1, x, valid, lane, h);
            }
            if (net == 1) {
                // head: o_lv = W2 relu(y1) + b2 (partial over this lane's channels, then over the 4 lanes)
                float ol[2][3];
#pragma unroll
                for (int r = 0; r < 2; ++r) ol[r][0] = ol[r][1] = ol[r][2] = 0.f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float2 st = S.W.st[1][8 * nt + 2 * t + i];
                        const float4 w2 = S.W.w2[1][8 * nt + 2 * t + i];
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const float a1 = fmaxf(fmaf(st.x, h[nt][2 * r + i], st.y), 0.f);
                            ol[r][0] = fmaf(w2.x, a1, ol[r][0]);
                            ol[r][1] = fmaf(w2.y, a1, ol[r][1]);
                            ol[r][2] = fmaf(w2.z, a1, ol[r][2]);
                        }
                    }
                const float4 b2 = S.W.b2[1];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        ol[r][d] += __shfl_xor_sync(0xffffffffu, ol[r][d], 1);
                        ol[r][d] += __shfl_xor_sync(0xffffffffu, ol[r][d], 2);
                    }
                    const float olv = (td == 0 ? ol[r][0] + b2.x : (td == 1 ? ol[r][1] + b2.y : ol[r][2] + b2.z));
                    const float lam = softsign(olv);
                    const float ex = expf(lam);
                    const float sig2 = GWTF_FLOW_EPS + ex;
                    const float sig = sqrtf(sig2);
                    const float gd = valid[r] ? gcd[r] : 0.f;
                    const float gin = gd / sig;
                    const float dlam = (valid[r] ? gsd[r] : 0.f) - gd * outd[r] * ex / (2.0f * sig2);
                    const float den = 1.0f + fabsf(olv);
                    const float dov = dlam / (den * den);
                    const int n = n0 + g + 8 * r;
                    if (valid[r] && t < 3) {
                        a.gbuf[sb + (size_t)td * N + n] = gin;
                        dob[(size_t)td * N + n] = -gin;
                        dob[(size_t)(3 + td) * N + n] = dov;
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d) dO[r][d] = __shfl_sync(0xffffffffu, dov, (lane & ~3) + d);
                }
            }
            // ---- per-channel sums: one tensor-memory chunk (two n-tiles) at a time
            tmem_wait_st();
#pragma unroll
            for (int c = 0; c < TM::NCH; ++c) {
                float v[24];
                const bool tail = c == NT / 2;
                if (tail) tmem_ld<TM::TAIL>(t_acc + 24 * c, reinterpret_cast<float(&)[TM::TAIL]>(v));
                else tmem_ld<24>(t_acc + 24 * c, v);
                tmem_wait_ld();
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const int nt = 2 * c + nn;
                    if (nt < NT && !(tail && nn == 1)) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float2 st = S.W.st[net][8 * nt + 2 * t + i];
                            const float2 mi = S.mif[net][8 * nt + 2 * t + i];
                            const float4 w2 = S.W.w2[net][8 * nt + 2 * t + i];
                            float* acc = v + nn * 10 + i * 5;
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                const float hv = h[nt][2 * r + i];
                                const float y1 = fmaf(st.x, hv, st.y);
                                const float da1 = w2.x * dO[r][0] + w2.y * dO[r][1] + w2.z * dO[r][2];
                                const float dy1 = y1 > 0.f ? da1 : 0.f;
                                const float a1 = fmaxf(y1, 0.f);
                                acc[0] = fmaf(dy1, fmaf(hv, mi.y, -mi.x), acc[0]);
                                acc[1] += dy1;
                                acc[2] = fmaf(dO[r][0], a1, acc[2]);
                                acc[3] = fmaf(dO[r][1], a1, acc[3]);
                                acc[4] = fmaf(dO[r][2], a1, acc[4]);
                            }
                        }
                    }
                }
                if (tail) {
                    if (t == 0) {                              // sd2 bias sums (identical on the 4 lanes of a row)
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            v[TAILB + 0] += dO[r][0]; v[TAILB + 1] += dO[r][1]; v[TAILB + 2] += dO[r][2];
                        }
                    }
                    tmem_st<TM::TAIL>(t_acc + 24 * c, reinterpret_cast<const float(&)[TM::TAIL]>(v));
                } else {
                    tmem_st<24>(t_acc + 24 * c, v);
                }
            }
        }
        if (cur_b >= 0) flush(cur_b);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(S.tmem_base, TM::alloc);
}

}  // namespace gwtf
