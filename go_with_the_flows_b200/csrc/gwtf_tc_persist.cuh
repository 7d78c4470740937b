// Persistent, warp-specialised tcgen05 forward of one coupling layer.
//
// One CTA per SM: 4 compute warpgroups (128 threads = 128 TMEM lanes each, one 128-point tile in flight
// per warpgroup, 128 TMEM columns per slot) + one MMA-issuer warp per slot.  The warpgroups never synchronise
// with each other on the hot path (only at a shape boundary, to restage the per-shape operands): a warpgroup
// hands its operands to its issuer through an mbarrier (`req[s]`, 128 arrivals) and sleeps on `done[s]`, which
// the issuer's tcgen05.commit completes, so the tensor pipe works on whichever tiles are ready while the other
// warpgroups run their relu / split / statistics on the CUDA cores.  The issuer is one ELECTED thread
// (elect.sync) working from warp-uniform values: it then issues tcgen05.mma back to back at the pipe's rate.
// Same math and operand layouts as gwtf_tc_fwd.cuh (all-GEMM chain MMA0 -> relu -> MMA1 -> relu -> MMA2).
#pragma once
#include "gwtf_tc_fwd.cuh"

namespace gwtf {

constexpr int kSlots = 4;
constexpr int kPersistThreads = kSlots * 128 + kSlots * 32;   // 4 compute warpgroups + one issuer warp per slot

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// named barrier over the compute warpgroups only (the issuer warp never joins it)
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSlots * 128) : "memory"); }

template <int FPK, int FPN>
struct TcPersistSmem {
    LayerT<FPN> W;
    TcLayerOps<FPK, FPN> ops[2];
    float x_hi[kSlots][128 * 8], x_lo[kSlots][128 * 8];
    uint64_t bar_tma, req[kSlots], done[kSlots];
    uint32_t tmem_base;
    float red[2][2 * FPN + 32];
    double dred[16];
    alignas(16) float col[kSlots * 4][kColTile];     // statistics pass: per-warp tiles of the column sums (gwtf_tc_fwd.cuh)
};

// the CTA's contiguous tile range cut into rounds of <= `slots` tiles that never straddle a shape.  Dense mode:
// every shape has `tps` tiles; segmented mode: shape b of this component owns tiles [pre[b], pre[b+1]).
struct RoundIter {
    int t, t_end, tps;
    const int32_t* pre;          // null = dense
    int b, s_begin, s_end, slots;
    __host__ __device__ __forceinline__ RoundIter(int t0, int t1, int tps_, const int32_t* pre_, int B, int slots_ = kSlots)
        : t(t0), t_end(t1), tps(tps_), pre(pre_), slots(slots_) {
        if (!pre) { b = t0 / tps; s_begin = b * tps; s_end = s_begin + tps; return; }
        int lo = 0, hi = B;                  // largest b with pre[b] <= t0
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= t0) lo = mid; else hi = mid; }
        b = lo; s_begin = pre[lo]; s_end = pre[lo + 1];
    }
    __host__ __device__ __forceinline__ bool next(int& base, int& count, int& bb) {
        if (t >= t_end) return false;
        while (t >= s_end) { ++b; s_begin = s_end; s_end = pre ? pre[b + 1] : s_end + tps; }
        const int ti = t - s_begin;
        int end = s_begin + min(s_end - s_begin, (ti / slots + 1) * slots);
        end = min(end, t_end);
        base = t;
        count = end - t;
        bb = b;
        t = end;
        return true;
    }
    __host__ __device__ __forceinline__ int shape_begin() const { return s_begin; }
};

// PASSES: 3 = fp32-grade 3xTF32 (training, anything that feeds gradients), 1 = single-pass TF32 for the
// no-grad eval-mode NLL and the sampling pass (gwtf_stack_desc.eval_precision)
template <int FPK, int FPN, int PHASE, int PASSES = 3>
__global__ void __launch_bounds__(kPersistThreads, 1) k_fwd_layer_tcp(const LayerArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = TcPersistSmem<FPK, FPN>;
    using C = TcCols<FPK, FPN>;
    constexpr int STAGES = PHASE == 0 ? 4 : 6;          // MMA batches per tile (both nets)
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    float* raw = reinterpret_cast<float*>(smem_raw + round_up((int)sizeof(SM), 16));
    const int F = a.d.n_features, K = a.d.n_components, L = a.d.n_layers;
    const int tid = threadIdx.x, lane = tid & 31;
    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything the MMA
    // issuer derives from it (tensor-memory addresses, shared-memory descriptors) in uniform registers
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int slot = warp < kSlots * 4 ? (warp >> 2) : kSlots;   // 0..3 compute warpgroups, 4 = issuer warps
    const int wtid = tid & 127;                         // thread within its warpgroup = TMEM lane
    const int j = blockIdx.y, l = a.layer;
    const int N = a.N, B = a.B;
    const bool train = a.train != 0;
    constexpr int CT = kSlots * 128;                    // compute threads
    const bool is_compute = slot < kSlots;

    LayerSrc src;
    src.params = a.params + (size_t)(j * L + l) * a.d.rec_stride;
    src.bn = a.bnbuf + (size_t)(j * L + l) * 8 * F;
    src.film = nullptr;
    src.mom = a.mom_in ? a.mom_in + j * GWTF_MOM_STRIDE : nullptr;
    src.sum1 = a.sum1 ? a.sum1 + (size_t)j * 4 * F : nullptr;
    src.n_total = a.n_total;

    if (warp == kSlots * 4) tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        mbar_init(&S.bar_tma, 1);
        for (int s = 0; s < kSlots; ++s) { mbar_init(&S.req[s], 128); mbar_init(&S.done[s], 1); }
        mbar_fence_init();
    }
    for (int i = tid; i < 2 * (2 * FPN + 32); i += kPersistThreads) (&S.red[0][0])[i] = 0.f;
    for (int i = tid; i < kSlots * 128 * 8; i += kPersistThreads) { (&S.x_hi[0][0])[i] = 0.f; (&S.x_lo[0][0])[i] = 0.f; }
    if (tid < 16) S.dred[tid] = 0.0;
    __syncthreads();
    // dependents may start launching only now: their CTAs allocate tensor memory too, and a CTA of THIS grid that
    // had triggered before its own tcgen05.alloc could starve behind them
    pdl_trigger();
    if (tid == 0) issue_layer_copy(raw, src, F, !train, false, &S.bar_tma);
    mbar_wait(&S.bar_tma, 0u);
    const NetOffsets o = net_offsets(F, popc3(a.d.warp_mask[l]));
    if (PHASE == 0) {                 // parameters only: may overlap the previous kernel's tail
#pragma unroll
        for (int net = 0; net < 2; ++net)
            stage_b1<FPK, FPN>(S.ops[net], raw + net * o.stride + o.W1, nullptr, F, tid, kPersistThreads);
    }
    pdl_wait();                       // the statistics / points written by the previous kernel are complete from here on
    stage_vectors<FPN, false>(S.W, (LayerWB<FPN>*)nullptr, raw, src, F, a.d.warp_mask[l], train, PHASE == 0, nullptr,
                              tid, kPersistThreads);
    __syncthreads();
#pragma unroll
    for (int net = 0; net < 2; ++net) {
        stage_b0<FPK, FPN>(S.ops[net], S.W.q0[net], F, tid, kPersistThreads);
        if (PHASE != 0) stage_b2<FPK, FPN>(S.ops[net], S.W.w2[net], S.W.b2[net], F, tid, kPersistThreads);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = S.tmem_base;

    const int32_t* pre = a.seg_tiles ? a.seg_tiles + (size_t)j * (B + 1) : nullptr;
    const int total_tiles = pre ? pre[B] : B * a.tiles_per_shape;
    const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = min(blockIdx.x * per_cta, total_tiles), t_end = min(t_begin + per_cta, total_tiles);

    if (!is_compute) {
        // ================= MMA issuers: warp (16 + s) serves slot s =================
        // One elected thread per issuer warp (elect.sync, not `lane == 0`: with a data-dependent lane test the compiler
        // re-materialises every descriptor through R2UR and a single thread then issues one tcgen05.mma per ~125
        // cycles; elected, the same loop issues them back to back at the tensor pipe's own rate -- tools/tc_probe4.cu)
        const int s = warp - kSlots * 4;
        if (elect_one()) {
            int remaining = 0;
            {
                RoundIter it(t_begin, t_end, a.tiles_per_shape, pre, B);
                int base, count, b;
                while (it.next(base, count, b)) remaining += (s < count) ? STAGES : 0;
            }
            uint32_t req_phase = 0u;
            int st = 0;
            const uint32_t tslot = tbase + s * kTcCols;
            while (remaining > 0) {
                mbar_wait(&S.req[s], req_phase);
                req_phase ^= 1u;
                tc_fence_after();
                const int net = st / (STAGES / 2), k = st - net * (STAGES / 2);
                if (k == 0) issue_ss_k8<FPN>(tslot + C::D, S.x_hi[s], S.x_lo[s], S.ops[net].B0.hi, S.ops[net].B0.lo);
                else if (k == 1) issue_ts<FPK, FPN, PASSES>(tslot + C::D, tslot + C::Ahi, tslot + C::Alo, S.ops[net].B1.hi, S.ops[net].B1.lo);
                else issue_ts<FPK, 16, PASSES>(tslot + C::D, tslot + C::Ahi, tslot + C::Alo, S.ops[net].B2.hi, S.ops[net].B2.lo);
                tc_commit(&S.done[s]);
                st = st + 1 == STAGES ? 0 : st + 1;
                --remaining;
            }
        }
        __syncwarp();
    } else {
        // ================= compute warpgroups =================
        const uint32_t trow = tbase + slot * kTcCols + ((uint32_t)((warp & 3) * 32) << 16);
        uint64_t* req = &S.req[slot];
        uint64_t* done = &S.done[slot];
        uint32_t done_phase = 0u;
        auto request = [&]() { tc_fence_before(); mbar_arrive(req); };
        auto wait_done = [&]() { mbar_wait(done, done_phase); done_phase ^= 1u; tc_fence_after(); };
        float mv[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) mv[i] = 0.f;
        // statistics pass: [net][sum h1, sum h1^2 of channel `lane` | of channel 32 + (lane & 7)], kept across tiles
        float st_acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float* coltile = S.col[warp];
        int cur_b = -1;
        RoundIter it(t_begin, t_end, a.tiles_per_shape, pre, B);
        FilmAhead<FPN> film;
        film.b = -1; film.s = 0.f; film.t = 0.f;
        if (PHASE == 1 && t_begin < t_end) film.fetch(a.film, it.b, B, K, j, L, l, F, tid);
        int base, count, b;
        while (it.next(base, count, b)) {
            if (PHASE == 1 && b != cur_b) {
                // new shape: every warpgroup has drained its MMAs (it waited on `done` for each request)
                compute_barrier();
                if (film.b != b) film.fetch(a.film, b, B, K, j, L, l, F, tid);
                film.template stage<false>(S.W, (LayerWB<FPN>*)nullptr, F, tid);
                film.fetch(a.film, b + 1, B, K, j, L, l, F, tid);      // in flight until the next shape boundary
                compute_barrier();
#pragma unroll
                for (int net = 0; net < 2; ++net)
                    stage_b1<FPK, FPN>(S.ops[net], raw + net * o.stride + o.W1, S.W.st[net], F, tid, CT);
                fence_proxy_async();
                compute_barrier();
                cur_b = b;
            }
            if (slot >= count) continue;
            const int t = base + slot;
            int n = (t - it.shape_begin()) * 128 + wtid;
            bool valid = n < N;
            if (a.seg) {                                            // segmented rows: (offset, count) of (j, b)
                const int2 sg = *reinterpret_cast<const int2*>(a.seg + ((size_t)j * B + b) * 2);
                valid = n < sg.y;
                n += sg.x;
            }
            float x[3], s3[3];
            const bool shared_rows = a.xin_shared || a.seg;
            const float* xin = shared_rows ? a.xin + (size_t)b * 3 * N : a.xin + ((size_t)j * B + b) * 3 * N;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                x[d] = valid ? xin[(size_t)d * N + n] : 0.f;
                s3[d] = (PHASE == 1 && a.ssum && valid) ? a.ssum[((size_t)j * B + b) * 3 * N + (size_t)d * N + n] : 0.f;
            }
            write_x_operand(S.x_hi[slot], S.x_lo[slot], x, wtid);
            fence_proxy_async();
            request();                                              // -> MMA0 (net 0)
            if (PHASE == 0) {
#pragma unroll 1
                for (int net = 0; net < 2; ++net) {
                    wait_done();                                    // y0
                    relu_to_operand<FPK, FPN>(trow);
                    request();                                      // -> MMA1
                    wait_done();                                    // h1
                    float h[FPK];                                   // (channels >= F are zero: F < FPK)
                    tmem_ld<FPK>(trow + C::D, h);
                    tmem_wait_ld();
                    if (net == 0) request();                        // -> MMA0 (net 1): accumulator is in registers now
                    // per-channel sum h1, sum h1^2 over the warp's 32 points: rows into the warp's tile, lane L sums column L
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < FPK; ++i) h[i] = 0.f;
                    }
                    __syncwarp();
                    col_write_row<FPK>(coltile, lane, h, 0.f, 0.f);
                    __syncwarp();
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                    for (int pp = 0; pp < 32; ++pp) {
                        const float v = lane < FPK ? coltile[pp * kColPitch + lane] : 0.f;
                        s1 += v;
                        s2 = fmaf(v, v, s2);
                    }
                    st_acc[net][0] += s1; st_acc[net][1] += s2;
                    if (FPK > 32) {
                        float t1 = 0.f, t2 = 0.f;
                        const int ch = 32 + (lane & 7), grp = lane >> 3;
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp) {
                            const float v = ch < FPK ? coltile[col_tail_point(grp, pp) * kColPitch + ch] : 0.f;
                            t1 += v;
                            t2 = fmaf(v, v, t2);
                        }
                        t1 += __shfl_xor_sync(0xffffffffu, t1, 8);  t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                        t2 += __shfl_xor_sync(0xffffffffu, t2, 8);  t2 += __shfl_xor_sync(0xffffffffu, t2, 16);
                        st_acc[net][2] += t1; st_acc[net][3] += t2;
                    }
                }
            } else {
                float o3[2][3];
#pragma unroll                                                      // (unrolled: o3[net] stays in registers)
                for (int net = 0; net < 2; ++net) {
                    wait_done();                                    // y0
                    relu_to_operand<FPK, FPN, PASSES>(trow);
                    request();                                      // -> MMA1
                    wait_done();                                    // y1
                    relu_to_operand<FPK, FPN, PASSES>(trow, keep_ptr(a.y1out, j, net, F, B, N, b, n), (F + 7) / 8, valid);
                    request();                                      // -> MMA2
                    wait_done();                                    // o
                    float ov[8];
                    tmem_ld8(trow + C::D, ov);
                    tmem_wait_ld();
                    o3[net][0] = ov[0]; o3[net][1] = ov[1]; o3[net][2] = ov[2];
                    if (net == 0) request();                        // -> MMA0 (net 1)
                }
                float lam[3];
                if (a.direct) warp_point<true>(x, o3[0], o3[1], lam);
                else warp_point<false>(x, o3[0], o3[1], lam);
                if (valid) {
                    const size_t gb = (a.seg ? (size_t)b : (size_t)j * B + b) * 3 * N + n;
#pragma unroll
                    for (int d = 0; d < 3; ++d) a.xout[gb + (size_t)d * N] = x[d];
                    if (a.ld) a.ld[((size_t)j * B + b) * N + n] += lam[0] + lam[1] + lam[2];
                    if (a.ssum)
#pragma unroll
                        for (int d = 0; d < 3; ++d) a.ssum[gb + (size_t)d * N] = s3[d] + lam[d];
                    if (a.trio) {
                        const size_t tb = (((size_t)j * 3) * B + b) * 3 * N + n;
                        const size_t ts = (size_t)B * 3 * N;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            a.trio[tb + (size_t)d * N] = x[d];
                            a.trio[tb + ts + (size_t)d * N] = o3[0][d];
                            a.trio[tb + 2 * ts + (size_t)d * N] = lam[d];
                        }
                    }
                    mv[0] += x[0]; mv[1] += x[1]; mv[2] += x[2];
                    mv[3] += x[0] * x[0]; mv[4] += x[0] * x[1]; mv[5] += x[0] * x[2];
                    mv[6] += x[1] * x[1]; mv[7] += x[1] * x[2]; mv[8] += x[2] * x[2];
                }
            }
        }
        if (PHASE == 0) {
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                if (lane < FPK) { atomicAdd(&S.red[net][2 * lane], st_acc[net][0]); atomicAdd(&S.red[net][2 * lane + 1], st_acc[net][1]); }
                if (lane < 8 && 32 + lane < FPK) {
                    atomicAdd(&S.red[net][2 * (32 + lane)], st_acc[net][2]);
                    atomicAdd(&S.red[net][2 * (32 + lane) + 1], st_acc[net][3]);
                }
            }
        }
        if (PHASE == 1 && a.mom_out) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = i < 9 ? mv[i] : 0.f;
            const float r = warp_reduce_scatter32(v, lane);
            if (lane < 9) atomicAdd(&S.dred[lane], (double)r);
        }
    }
    // ---- flush block partials
    tc_fence_before();
    __syncthreads();
    if (PHASE == 0) {
        for (int i = tid; i < 2 * 2 * FPN; i += kPersistThreads) {
            const int net = i / (2 * FPN), idx = i - net * 2 * FPN, f = idx >> 1, which = idx & 1;
            if (f < F) atomicAdd(&a.sum1[((size_t)j * 2 + net) * 2 * F + which * F + f], (double)S.red[net][idx]);
        }
    } else if (a.mom_out) {
        if (tid < 9) atomicAdd(&a.mom_out[j * GWTF_MOM_STRIDE + tid], S.dred[tid]);
    }
    if (warp == kSlots * 4) tmem_dealloc(tbase, 512);
    exchange_tail(a.tail);
}

}  // namespace gwtf
