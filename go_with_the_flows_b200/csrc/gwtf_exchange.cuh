// Batch-statistic exchange between ranks over NVLink peer memory (one process per GPU).
//
// Train-mode BatchNorm over the point dimension makes every layer phase end in a few hundred fp64 sums
// that all ranks must add up before the next phase can start (SyncBatchNorm semantics, the reference's
// train_ae.py:77-78 + torch.nn.SyncBatchNorm): 4 dependent exchanges per layer, 132 per step.  They are
// latency, not bandwidth: instead of a library all-reduce per exchange (launch + protocol latency, and a
// Python call in between), one tiny kernel per exchange PUSHES this rank's partial sums straight into a
// slot of every peer's receive buffer with P2P stores, publishes a monotonically increasing sequence
// number on every peer (release, system scope), waits until every peer's number has arrived here, and adds
// the R slots in rank order -- every rank gets bit-identical totals.  Slots are double-buffered by the
// parity of the sequence number: a rank can be at most one exchange ahead of the slowest peer, because it
// cannot finish exchange e+1 without that peer's flag for e+1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gwtf {

constexpr int kMaxRanks = 16;

struct ExchangeArgs {
    int rank = 0, world = 0, n = 0, slot = 0; // n doubles to add up, slot = doubles per (parity, rank) slot
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 0;        // how long to wait for the peers before giving up
    double* data = nullptr;                   // in: this rank's partial sums, out: the totals
    double* recv[kMaxRanks] = {};             // receive buffer of every rank: [2][world][slot] doubles
    unsigned long long* flags[kMaxRanks] = {};  // flag array of every rank: [0, world) flags, [kTicketSlot] scratch
};

// The exchange proper, by all threads of one CTA.  `data` is read through L2 (the partial sums were accumulated with
// atomics by other CTAs).
__device__ __forceinline__ void exchange_body(const ExchangeArgs& a) {
    const int tid = threadIdx.x;
    const size_t par = (size_t)(a.seq & 1ull);
    for (int r = 0; r < a.world; ++r) {
        double* dst = a.recv[r] + (par * a.world + a.rank) * a.slot;
        for (int i = tid; i < a.n; i += blockDim.x) dst[i] = __ldcg(a.data + i);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) {
        unsigned long long* theirs = a.flags[tid] + a.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(a.seq) : "memory");
        const unsigned long long* mine = a.flags[a.rank] + tid;
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > a.timeout_ns) __trap();        // a peer never showed up: fail (sticky error) instead of hanging
        } while (v < a.seq);
    }
    __syncthreads();
    const double* in = a.recv[a.rank] + par * a.world * a.slot;
    for (int i = tid; i < a.n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += __ldcg(in + (size_t)r * a.slot + i);
        a.data[i] = s;
    }
}

static __global__ void __launch_bounds__(256) k_exchange_sum(const ExchangeArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the consumer's prologue may overlap the exchange
    asm volatile("griddepcontrol.wait;" ::: "memory");                // the producer's sums are complete
    exchange_body(a);
}

// The same exchange folded into the tail of the kernel that produces the sums (the tcgen05 layer kernels): the last CTA
// of the grid to get here -- a ticket in flags[rank][kTicketSlot] -- runs it, so a multi-rank step has no launches a
// single-rank step does not have.  Every thread of every CTA calls this once, after its last atomic on `x.data`.
// world <= 1: no exchange.
constexpr int kTicketSlot = 31;               // the flag arrays hold >= 32 uint64; [0, world) are the flags proper
struct ExchangeTail {
    ExchangeArgs x;
};
__device__ __forceinline__ void exchange_tail(const ExchangeTail& t) {
    if (t.x.world <= 1) return;
    __shared__ int s_last;
    __threadfence();                           // this CTA's atomics before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* ticket = t.x.flags[t.x.rank] + kTicketSlot;
        const unsigned long long k = atomicAdd(ticket, 1ull);
        s_last = k == (unsigned long long)(gridDim.x * gridDim.y) - 1ull;
        if (s_last) *ticket = 0ull;            // (nobody else touches it until the next launch)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    exchange_body(t.x);
}

}  // namespace gwtf
