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
    int ll = 0;                               // 1: 16-byte cells carrying their own sequence number (exchange_body_ll)
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

// Low-latency variant (the default): no separate flag, no system fence between data and flag.  Every double travels as
// one 16-byte cell {lo32, seq32, hi32, seq32} written with a single vector store; each 8-byte half carries its own copy
// of the (32-bit, non-zero) sequence number, so a reader that sees both copies has both halves -- the only atomicity
// assumed of a peer store is 8 bytes (the scheme of NCCL's LL protocol).  The receive buffer then holds
// [2][world][slot] cells = twice the bytes; parity double-buffering and the "at most one exchange ahead" argument are
// unchanged.  One NVLink latency per exchange instead of data -> fence -> flag -> poll.
__device__ __forceinline__ void exchange_body_ll(const ExchangeArgs& a) {
    const int tid = threadIdx.x;
    const size_t par = (size_t)(a.seq & 1ull);
    const uint32_t tag = (uint32_t)a.seq;                      // never 0: sequence numbers start at 1, buffers at 0
    for (int r = 0; r < a.world; ++r) {
        uint4* dst = reinterpret_cast<uint4*>(a.recv[r]) + (par * a.world + a.rank) * a.slot;
        for (int i = tid; i < a.n; i += blockDim.x) {
            const unsigned long long v = (unsigned long long)__double_as_longlong(__ldcg(a.data + i));
            const uint4 cell = make_uint4((uint32_t)v, tag, (uint32_t)(v >> 32), tag);
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(cell.x), "r"(cell.y),
                         "r"(cell.z), "r"(cell.w) : "memory");
        }
    }
    const uint4* in = reinterpret_cast<const uint4*>(a.recv[a.rank]) + par * a.world * a.slot;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = tid; i < a.n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) {
            const uint4* src = in + (size_t)r * a.slot + i;
            uint4 c;
            int spins = 0;
            do {
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
                             : "l"(src) : "memory");
                if (((++spins) & 1023) == 0) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > a.timeout_ns) __trap();       // a peer never showed up: fail (sticky error), do not hang
                }
            } while (c.y != tag || c.w != tag);
            s += __longlong_as_double((long long)(((unsigned long long)c.z << 32) | c.x));
        }
        a.data[i] = s;
    }
}

__device__ __forceinline__ void exchange_run(const ExchangeArgs& a) {
    if (a.ll) exchange_body_ll(a); else exchange_body(a);
}

static __global__ void __launch_bounds__(1024) k_exchange_sum(const ExchangeArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the consumer's prologue may overlap the exchange
    asm volatile("griddepcontrol.wait;" ::: "memory");                // the producer's sums are complete
    exchange_run(a);
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
    exchange_run(t.x);
}

}  // namespace gwtf
