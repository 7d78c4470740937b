// Forward layer phases on warp-level tensor-core fragments (mma.sync.m16n8k8 tf32, 3xTF32 split).
//
// Same LayerArgs / outputs as k_fwd_layer (gwtf_fwd.cuh):
//   PHASE 0: sd1_bn batch statistics: sum h1, sum h1^2 per channel and net       (flows.py:96-104, train mode)
//   PHASE 1: apply: heads -> softsign -> affine update of the points, log-det sums, moments of the output
// One warp owns 16 points (MMA M); a0 = relu(q0 x) is computed straight into A fragments, h1 = a0 W1^T
// comes back as C fragments (rows g, g+8; channels 8nt+2t, 8nt+2t+1), everything after is element-wise
// on those fragments.  PHASE 1 can keep h1 for the backward pass (`hkeep`, layout in gwtf_mma.cuh).
#pragma once
#include "gwtf_fwd.cuh"
#include "gwtf_mma.cuh"

namespace gwtf {

template <int FP>
__host__ __device__ constexpr size_t fwd_mma_smem(int F) {
    return round_up((int)sizeof(PhaseSmem<FP>), 16) + (size_t)round_up(raw_floats(F), 4) * 4 +
           2 * (size_t)(FP / 8) * (FP / 8) * 32 * 16;
}

template <int FP, int PHASE, int MI>
__global__ void __launch_bounds__(kThreads, MI == 1 ? 2 : 1) k_fwd_layer_mma(const LayerArgs a, float* hkeep) {
    static_assert(FP % 8 == 0, "feature width padded to the MMA K");
    constexpr int NT = FP / 8;
    constexpr int NW = kThreads / 32;
    constexpr int TILE = NW * 16 * MI;                       // MI m-tiles per warp iteration
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem<FP>& S = *reinterpret_cast<PhaseSmem<FP>*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(PhaseSmem<FP>), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    float4* bf = reinterpret_cast<float4*>(raw + round_up(raw_floats(F), 4));     // [2 nets][KS][NT][32]
    constexpr int BFN = NT * NT * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int td = t < 3 ? t : 2;
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
    for (int i = tid; i < 2 * (2 * FP + 32); i += kThreads) (&S.red[0][0])[i] = 0.f;
    if (tid < 16) S.dred[tid] = 0.0;
    __syncthreads();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar);
    mbar_wait(&S.bar, 0u);
    stage_layer<FP, false>(S.W, nullptr, raw, src, F, a.d.warp_mask[l], train, PHASE == 0, nullptr, tid, kThreads);
    __syncthreads();
    stage_bfrag_h1<FP>(bf, S.W.W1T[0], tid, kThreads);
    stage_bfrag_h1<FP>(bf + BFN, S.W.W1T[1], tid, kThreads);
    __syncthreads();

    const int tps = (N + TILE - 1) / TILE;
    const long long total_tiles = (long long)B * tps;
    const int t_begin = (int)(total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)(total_tiles * (blockIdx.x + 1) / gridDim.x);
    const size_t npad = keep_npad(N);

    if constexpr (PHASE == 0) {
        // one net at a time (two passes over the tiles): 20 accumulators per lane instead of 40
#pragma unroll 1
        for (int net = 0; net < 2; ++net) {
            float s1[NT][2], s2[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
#pragma unroll 1
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int b = tile / tps;
                const int n0 = (tile - b * tps) * TILE + warp * (16 * MI);
                const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
                float x[MI][2][3];
                bool valid[MI][2];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int n = n0 + mi * 16 + g + 8 * r;
                        valid[mi][r] = n < N;
#pragma unroll
                        for (int d = 0; d < 3; ++d) x[mi][r][d] = valid[mi][r] ? xin[(size_t)d * N + n] : 0.f;
                    }
                float h[MI][NT][4];
                mma_h1<FP, MI>(S.W.q0[net], bf + net * BFN, x, valid, lane, h);
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            s1[nt][i] += h[mi][nt][i] + h[mi][nt][2 + i];
                            s2[nt][i] = fmaf(h[mi][nt][i], h[mi][nt][i], fmaf(h[mi][nt][2 + i], h[mi][nt][2 + i], s2[nt][i]));
                        }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float u = s1[nt][i], v = s2[nt][i];
#pragma unroll
                    for (int m = 4; m <= 16; m <<= 1) {
                        u += __shfl_xor_sync(0xffffffffu, u, m);
                        v += __shfl_xor_sync(0xffffffffu, v, m);
                    }
                    if (g == 0) {
                        atomicAdd(&S.red[net][2 * (8 * nt + 2 * t + i)], u);
                        atomicAdd(&S.red[net][2 * (8 * nt + 2 * t + i) + 1], v);
                    }
                }
        }
        __syncthreads();
        for (int i = tid; i < 2 * 2 * FP; i += kThreads) {
            const int net = i / (2 * FP), idx = i - net * 2 * FP, f = idx >> 1, which = idx & 1;
            if (f < F) atomicAdd(&a.sum1[((size_t)j * 2 + net) * 2 * F + which * F + f], (double)S.red[net][idx]);
        }
    } else {
        float mo[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) mo[i] = 0.f;
        int cur_b = -1;
#pragma unroll 1
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int b = tile / tps;
            const int n0 = (tile - b * tps) * TILE + warp * (16 * MI);
            if (b != cur_b) {
                __syncthreads();
                stage_film<FP, false>(S.W, (LayerWB<FP>*)nullptr, a.film + ((size_t)(b * K + j) * L + l) * 4 * F, F, tid, kThreads);
                __syncthreads();
                cur_b = b;
            }
            const float* xin = a.xin_shared ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
            const size_t base = ((size_t)j * B + b) * 3 * N;
            float x[MI][2][3], ssold[MI][2], ldold[MI][2];
            bool valid[MI][2];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int n = n0 + mi * 16 + g + 8 * r;
                    valid[mi][r] = n < N;
#pragma unroll
                    for (int d = 0; d < 3; ++d) x[mi][r][d] = valid[mi][r] ? xin[(size_t)d * N + n] : 0.f;
                    ssold[mi][r] = (valid[mi][r] && a.ssum && t < 3) ? a.ssum[base + (size_t)t * N + n] : 0.f;
                    ldold[mi][r] = (valid[mi][r] && a.ld && t == 0) ? a.ld[((size_t)j * B + b) * N + n] : 0.f;
                }
            float o[2][MI][2][3];                              // [net][m-tile][row][dim] head outputs
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                float h[MI][NT][4];
                mma_h1<FP, MI>(S.W.q0[net], bf + net * BFN, x, valid, lane, h);
                if (hkeep) {
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
                        store_h1_frag<NT>(hkeep + keep_slab(F, B, N, j, net), (size_t)b * npad + n0 + mi * 16, lane, h[mi]);
                }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int r = 0; r < 2; ++r) o[net][mi][r][0] = o[net][mi][r][1] = o[net][mi][r][2] = 0.f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float2 st = S.W.st[net][8 * nt + 2 * t + i];
                        const float4 w2 = S.W.w2[net][8 * nt + 2 * t + i];
#pragma unroll
                        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                const float a1 = fmaxf(fmaf(st.x, h[mi][nt][2 * r + i], st.y), 0.f);
                                o[net][mi][r][0] = fmaf(w2.x, a1, o[net][mi][r][0]);
                                o[net][mi][r][1] = fmaf(w2.y, a1, o[net][mi][r][1]);
                                o[net][mi][r][2] = fmaf(w2.z, a1, o[net][mi][r][2]);
                            }
                    }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            o[net][mi][r][d] += __shfl_xor_sync(0xffffffffu, o[net][mi][r][d], 1);
                            o[net][mi][r][d] += __shfl_xor_sync(0xffffffffu, o[net][mi][r][d], 2);
                        }
            }
            // ---- lane t < 3 finishes dimension t of its rows
            const float4 b2m = S.W.b2[0], b2v = S.W.b2[1];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int n = n0 + mi * 16 + g + 8 * r;
                    const float (&om)[3] = o[0][mi][r];
                    const float (&ov)[3] = o[1][mi][r];
                    const float omu = td == 0 ? om[0] + b2m.x : (td == 1 ? om[1] + b2m.y : om[2] + b2m.z);
                    const float olv = td == 0 ? ov[0] + b2v.x : (td == 1 ? ov[1] + b2v.y : ov[2] + b2v.z);
                    const float xd = td == 0 ? x[mi][r][0] : (td == 1 ? x[mi][r][1] : x[mi][r][2]);
                    const float lam = softsign(olv);
                    const float sig = sqrtf(GWTF_FLOW_EPS + expf(lam));
                    const float xn = a.direct ? fmaf(sig, xd, omu) : (xd - omu) / sig;
                    const bool vr = valid[mi][r];
                    if (vr && t < 3) {
                        a.xout[base + (size_t)td * N + n] = xn;
                        if (a.ssum) a.ssum[base + (size_t)td * N + n] = ssold[mi][r] + lam;
                        if (a.trio) {
                            const size_t tb = (((size_t)j * 3) * B + b) * 3 * N + n;
                            const size_t ts = (size_t)B * 3 * N;
                            a.trio[tb + (size_t)td * N] = xn;
                            a.trio[tb + ts + (size_t)td * N] = omu;
                            a.trio[tb + 2 * ts + (size_t)td * N] = lam;
                        }
                    }
                    // gather the row's three dims on every lane of the row group
                    const int l0 = lane & ~3;
                    const float x0 = __shfl_sync(0xffffffffu, xn, l0), x1 = __shfl_sync(0xffffffffu, xn, l0 + 1),
                                x2 = __shfl_sync(0xffffffffu, xn, l0 + 2);
                    const float lsum = __shfl_sync(0xffffffffu, lam, l0) + __shfl_sync(0xffffffffu, lam, l0 + 1) +
                                       __shfl_sync(0xffffffffu, lam, l0 + 2);
                    if (vr && t == 0) {
                        if (a.ld) a.ld[((size_t)j * B + b) * N + n] = ldold[mi][r] + lsum;
                        mo[0] += x0; mo[1] += x1; mo[2] += x2;
                        mo[3] = fmaf(x0, x0, mo[3]); mo[4] = fmaf(x0, x1, mo[4]); mo[5] = fmaf(x0, x2, mo[5]);
                        mo[6] = fmaf(x1, x1, mo[6]); mo[7] = fmaf(x1, x2, mo[7]); mo[8] = fmaf(x2, x2, mo[8]);
                    }
                }
        }
        if (a.mom_out) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float v = mo[i];                              // only lanes with t == 0 hold sums
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane == 0) atomicAdd(&S.dred[i], (double)v);
            }
            __syncthreads();
            if (tid < 9) atomicAdd(&a.mom_out[j * GWTF_MOM_STRIDE + tid], S.dred[tid]);
        }
    }
}

}  // namespace gwtf
