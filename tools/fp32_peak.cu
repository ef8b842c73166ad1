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
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"ctas_per_sm\": %d, \"ilp\": %d, \"fp32_tflops_burst\": %.2f, "
           "\"fp32_tflops_sustained\": %.2f, \"sustained_seconds\": %.2f, \"theoretical_tflops_at_max_clock\": %.2f, "
           "\"max_clock_mhz\": %.0f}\n",
           p.name, p.multiProcessorCount, per_sm, ILP, best, sustained, ms_tot * 1e-3,
           2.0 * p.multiProcessorCount * 128.0 * clk_khz * 1e3 / 1e12, clk_khz / 1e3);
    return 0;
}
