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
    int rank, world, n, slot;                 // n doubles to add up, slot = doubles per (parity, rank) slot
    unsigned long long seq;
    unsigned long long timeout_ns;            // how long to wait for the peers before giving up
    double* data;                             // in: this rank's partial sums, out: the totals
    double* recv[kMaxRanks];                  // receive buffer of every rank: [2][world][slot] doubles
    unsigned long long* flags[kMaxRanks];     // flag array of every rank: [world]
};

static __global__ void __launch_bounds__(256) k_exchange_sum(const ExchangeArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the consumer's prologue may overlap the exchange
    asm volatile("griddepcontrol.wait;" ::: "memory");                // the producer's sums are complete
    const int tid = threadIdx.x;
    const size_t par = (size_t)(a.seq & 1ull);
    for (int r = 0; r < a.world; ++r) {
        double* dst = a.recv[r] + (par * a.world + a.rank) * a.slot;
        for (int i = tid; i < a.n; i += blockDim.x) dst[i] = a.data[i];
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

}  // namespace gwtf
