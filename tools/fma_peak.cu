// Microbenchmark: FP32 FMA-pipe peak on this GPU (scalar FFMA vs packed fma.rn.f32x2) and
// the contraction inner loop shape used by the flow kernels (LDS.128 broadcast : FFMA ratio).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak tools/fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long acc[NACC];
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float v = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// contraction-shaped loop: per e: 1 LDS.128 (q) + FP/4 LDS.128 (weights), P*(3 + FP) FFMA
template <int FP, int P>
__global__ void __launch_bounds__(256) k_contract(float* out, int iters, int F) {
    extern __shared__ float4 sm[];
    float4* q0 = sm;                 // [64]
    float4* w1 = sm + 64;            // [64][FP/4]
    for (int i = threadIdx.x; i < 64 + 64 * FP / 4; i += blockDim.x)
        sm[i] = make_float4(1e-3f * i, 2e-3f, -1e-3f, 1e-4f * i);
    __syncthreads();
    float x[P][3];
#pragma unroll
    for (int p = 0; p < P; ++p) { x[p][0] = threadIdx.x * 1e-3f + p; x[p][1] = 0.5f * p; x[p][2] = 0.25f; }
    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        float acc[P][FP];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int f = 0; f < FP; ++f) acc[p][f] = 0.f;
#pragma unroll 1
        for (int e = 0; e < F; ++e) {
            float4 q = q0[e];
            float a[P];
#pragma unroll
            for (int p = 0; p < P; ++p)
                a[p] = fmaxf(fmaf(q.x, x[p][0], fmaf(q.y, x[p][1], fmaf(q.z, x[p][2], q.w))), 0.f);
#pragma unroll
            for (int f4 = 0; f4 < FP / 4; ++f4) {
                float4 w = w1[e * (FP / 4) + f4];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    acc[p][4 * f4 + 0] = fmaf(w.x, a[p], acc[p][4 * f4 + 0]);
                    acc[p][4 * f4 + 1] = fmaf(w.y, a[p], acc[p][4 * f4 + 1]);
                    acc[p][4 * f4 + 2] = fmaf(w.z, a[p], acc[p][4 * f4 + 2]);
                    acc[p][4 * f4 + 3] = fmaf(w.w, a[p], acc[p][4 * f4 + 3]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float s = 0.f;
#pragma unroll
            for (int f = 0; f < FP; ++f) s += fmaxf(acc[p][f], 0.f);
            x[p][0] = s * 1e-6f;
            tot += s;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

// same, packed: acc pairs over features, fma.rn.f32x2 with {a,a}
template <int FP, int P>
__global__ void __launch_bounds__(256) k_contract2(float* out, int iters, int F) {
    extern __shared__ float4 sm[];
    float4* q0 = sm;
    float4* w1 = sm + 64;
    for (int i = threadIdx.x; i < 64 + 64 * FP / 4; i += blockDim.x)
        sm[i] = make_float4(1e-3f * i, 2e-3f, -1e-3f, 1e-4f * i);
    __syncthreads();
    float x[P][3];
#pragma unroll
    for (int p = 0; p < P; ++p) { x[p][0] = threadIdx.x * 1e-3f + p; x[p][1] = 0.5f * p; x[p][2] = 0.25f; }
    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        unsigned long long acc[P][FP / 2];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int f = 0; f < FP / 2; ++f) acc[p][f] = 0ull;
#pragma unroll 1
        for (int e = 0; e < F; ++e) {
            float4 q = q0[e];
            unsigned long long a2[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float a = fmaxf(fmaf(q.x, x[p][0], fmaf(q.y, x[p][1], fmaf(q.z, x[p][2], q.w))), 0.f);
                asm("mov.b64 %0, {%1, %1};" : "=l"(a2[p]) : "f"(a));
            }
#pragma unroll
            for (int f4 = 0; f4 < FP / 4; ++f4) {
                ulonglong2 w = reinterpret_cast<const ulonglong2*>(w1)[e * (FP / 4) + f4];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * f4 + 0]) : "l"(w.x), "l"(a2[p]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * f4 + 1]) : "l"(w.y), "l"(a2[p]));
                }
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float s = 0.f;
#pragma unroll
            for (int f = 0; f < FP / 2; ++f) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[p][f]));
                s += fmaxf(lo, 0.f) + fmaxf(hi, 0.f);
            }
            x[p][0] = s * 1e-6f;
            tot += s;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <typename Fn>
float time_ms(Fn fn) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fn(); fn();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); fn(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    printf("device %s SMs %d clock %d kHz\n", pr.name, sms, pr.clockRate);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256 * 4));
    const int iters = 20000;
    {
        for (int bps = 1; bps <= 4; bps *= 2) {
            int grid = sms * bps;
            float ms = time_ms([&] { k_ffma<32><<<grid, 256>>>(out, iters, 1.0001f, 1e-4f); });
            double fl = 2.0 * 32 * iters * 256.0 * grid;
            printf("ffma   scalar  blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
            ms = time_ms([&] { k_ffma2<16><<<grid, 256>>>(out, iters, 1.0001f, 1e-4f); });
            printf("ffma2  packed  blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());
    {
        const int F = 37, it2 = 400;
        size_t smem = sizeof(float4) * (64 + 64 * 10);
        auto report = [&](const char* name, float ms, int P, int grid) {
            double fl = 2.0 * (double)F * (40 + 3) * P * it2 * 256.0 * grid;
            double useful = 2.0 * (double)F * (37 + 3) * P * it2 * 256.0 * grid;
            printf("%s P=%d grid %d: %.3f ms  issued %.2f TFLOP/s  useful %.2f\n", name, P, grid, ms, fl / ms * 1e-9, useful / ms * 1e-9);
        };
        for (int bps = 1; bps <= 2; ++bps) {
            int grid = sms * bps;
            report("contract scalar", time_ms([&] { k_contract<40, 2><<<grid, 256, smem>>>(out, it2, F); }), 2, grid);
            report("contract scalar", time_ms([&] { k_contract<40, 4><<<grid, 256, smem>>>(out, it2, F); }), 4, grid);
            report("contract packed", time_ms([&] { k_contract2<40, 2><<<grid, 256, smem>>>(out, it2, F); }), 2, grid);
            report("contract packed", time_ms([&] { k_contract2<40, 4><<<grid, 256, smem>>>(out, it2, F); }), 4, grid);
        }
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
