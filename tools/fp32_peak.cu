// fp32_peak.cu -- measured FP32 FMA peak of the device ("FP32 peak is not in MEASURED_PEAKS.json
// -- builder must microbenchmark it").  Every thread runs ILP independent chains of dependent FFMAs; the grid is a
// whole number of waves (SM count x resident CTAs, both queried), long enough that launch overhead vanishes.
// Reported: the best short run ("burst", what a kernel timed alone can reach) and a back-to-back run of a few
// seconds ("sustained", power-limited clocks), with the SM clock sampled through NVML by the Python wrapper
// (tools/fp32_peak.py), which also writes profiles/fp32_peak.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp32_peak.cu -o tools/fp32_peak
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int ILP>
__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float a, float b)
{
    float x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true: keeps the chains alive
}

// Register-operand form (what a real kernel issues: both multiplicands in registers) and the packed form of
// sm_100 (fma.rn.f32x2 -> SASS FFMA2: two FMAs per lane and instruction, one issue slot).
template <int ILP>
__global__ void __launch_bounds__(256) k_ffma_rrr(float *out, int iters, float a, float b)
{
    float x[ILP], m[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) { x[k] = (float)(threadIdx.x + k) * 1e-3f; m[k] = a + (float)(k + (int)threadIdx.x) * 1e-7f; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fmaf(x[k], m[k], m[(k + 1) % ILP]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long aa, bb, cc, rr;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(aa), "l"(bb), "l"(cc));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rr));
    return r;
}

// BCAST = 1: one multiplicand is a scalar register broadcast to both halves (FFMA2 R, Rw.F32, Rx.F32x2, Racc)
template <int ILP, int BCAST>
__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, float a, float b)
{
    float2 x[ILP], m[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) {
        x[k] = make_float2((float)(threadIdx.x + k) * 1e-3f, (float)(threadIdx.x + k) * 2e-3f);
        m[k] = make_float2(a + (float)(k + (int)threadIdx.x) * 1e-7f, a - (float)(k + (int)threadIdx.x) * 1e-7f);
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++)
                x[k] = BCAST ? ffma2(make_float2(m[k].x, m[k].x), x[k], m[(k + 1) % ILP]) : ffma2(m[k], x[k], m[(k + 1) % ILP]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k].x + x[k].y;
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}

#define CK(c) do { cudaError_t e = (c); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #c, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv)
{
    const double sustain_s = argc > 1 ? atof(argv[1]) : 3.0;
    constexpr int ILP = 8;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ffma<ILP>, 256, 0));
    const int grid = p.multiProcessorCount * per_sm;
    float *out;
    CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 20000; // x 8 x ILP FFMAs per thread
    const double flop = 2.0 * (double)grid * 256.0 * (double)iters * 8.0 * ILP;
    for (int w = 0; w < 3; w++) k_ffma<ILP><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
    CK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int r = 0; r < 10; r++) {
        CK(cudaEventRecord(e0));
        k_ffma<ILP><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    // sustained: back to back for sustain_s seconds
    int n = 0;
    float ms_tot = 0.f;
    CK(cudaEventRecord(e0));
    do {
        for (int k = 0; k < 20; k++) k_ffma<ILP><<<grid, 256>>>(out, iters, 0.999f, 1e-3f);
        n += 20;
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_tot, e0, e1));
    } while (ms_tot < sustain_s * 1e3);
    const double sustained = flop * n / (ms_tot * 1e-3) / 1e12;
    // the other instruction forms, same grid, best of 5 (flop per launch differs: FFMA2 does two FMAs per lane)
    auto best_of = [&](auto launch, double fl) -> double {
        double bst = 0.0;
        for (int r = 0; r < 6; r++) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r) bst = fl / (ms * 1e-3) / 1e12 > bst ? fl / (ms * 1e-3) / 1e12 : bst;
        }
        return bst;
    };
    const double rrr = best_of([&] { k_ffma_rrr<ILP><<<grid, 256>>>(out, iters, 0.999f, 1e-3f); }, flop);
    const double p2 = best_of([&] { k_ffma2<ILP, 0><<<grid, 256>>>(out, iters, 0.999f, 1e-3f); }, 2.0 * flop);
    const double p2b = best_of([&] { k_ffma2<ILP, 1><<<grid, 256>>>(out, iters, 0.999f, 1e-3f); }, 2.0 * flop);
    CK(cudaGetLastError());
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"ctas_per_sm\": %d, \"ilp\": %d, \"fp32_tflops_burst\": %.2f, "
           "\"fp32_tflops_sustained\": %.2f, \"sustained_seconds\": %.2f, \"ffma_reg_operands_tflops\": %.2f, "
           "\"ffma2_packed_tflops\": %.2f, \"ffma2_packed_broadcast_tflops\": %.2f, \"theoretical_tflops_at_max_clock\": %.2f, "
           "\"max_clock_mhz\": %.0f}\n",
           p.name, p.multiProcessorCount, per_sm, ILP, best, sustained, ms_tot * 1e-3, rrr, p2, p2b,
           2.0 * p.multiProcessorCount * 128.0 * clk_khz * 1e3 / 1e12, clk_khz / 1e3);
    return 0;
}
