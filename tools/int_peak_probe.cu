// Measurement probe (not product code): the integer issue roofline of this B200 -- how many warp instructions per
// second the SMs retire for the instruction kinds the Blokus legality kernels are made of (LOP3, SHF, IADD/IMAD, LDS,
// POPC), each as long independent chains so that only the issue / pipe rate limits.  SURVEY.md 8(d): "measure an int32
// micro-benchmark on the box for the real ALU peak".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak_probe tools/int_peak_probe.cu && tools/int_peak_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 4096, ILP = 8;

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[256 * 2];
    sm[threadIdx.x] = seed + threadIdx.x; sm[256 + threadIdx.x] = seed ^ threadIdx.x;
    __syncthreads();
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = seed + threadIdx.x * 977u + i;
    const uint32_t a = seed | 1u, b = seed * 3u + 5u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const uint32_t y = x[(i + 1) % ILP], z = x[(i + 3) % ILP];                 // (operands rotate: nothing folds)
            if (KIND == 0) x[i] = (x[i] ^ y) & (z | a);                             // LOP3 (3 register inputs)
            if (KIND == 1) x[i] = __funnelshift_r(x[i], y, z);                      // SHF
            if (KIND == 2) x[i] = x[i] + y + z;                                     // IADD3
            if (KIND == 3) x[i] = x[i] * y + z;                                     // IMAD
            if (KIND == 4) x[i] = __popc(x[i] ^ y) + z;                             // POPC (+ LOP3 + IADD)
            if (KIND == 5) x[i] = sm[(x[i] + y) & 511u];                            // LDS (+ IADD + LOP3)
            if (KIND == 6) x[i] = ((x[i] >> (y & 7u)) & z) | (x[i] + y);            // SHF + IADD + LOP3 x2 mix
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= x[i];
    if (r == 0x12345678u) out[0] = r;
}

static double g_rate[8];   // warp-inst / clk / SM per kind (for the JSON line)

template <int KIND>
int run(const char *name, double inst_per_iter, int sms, double mhz) {
    uint32_t *out;
    CK(cudaMalloc(&out, 4));
    const int blocks = sms * 8;
    k<KIND><<<blocks, 256>>>(out, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        k<KIND><<<blocks, 256>>>(out, 12345u + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double warp_inst = (double)blocks * 8 * ITERS * ILP * inst_per_iter;
    const double rate = warp_inst / (best * 1e-3);
    printf("%-28s %8.3f ms  %7.2f T warp-inst/s  %5.2f warp-inst / clk / SM (at %.0f MHz)\n", name, best, rate / 1e12,
           rate / (sms * mhz * 1e6), mhz);
    g_rate[KIND] = rate / (sms * mhz * 1e6);
    cudaFree(out);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double mhz = khz / 1e3;
    printf("%s: %d SMs, max SM clock %.0f MHz; peak issue = 4 warp-inst / clk / SM = %.2f T warp-inst/s\n", p.name,
           p.multiProcessorCount, mhz, 4.0 * p.multiProcessorCount * mhz * 1e6 / 1e12);
    const int s = p.multiProcessorCount;
    run<0>("LOP3", 1, s, mhz); run<1>("SHF", 1, s, mhz); run<2>("IADD", 1, s, mhz); run<3>("IMAD", 1, s, mhz);
    run<4>("POPC + LOP3 + IADD", 3, s, mhz); run<5>("LDS + IADD + LOP3", 3, s, mhz); run<6>("SHF + IADD + 2 LOP3 mix", 4, s, mhz);
    // one JSON line (stderr) for INT_PEAKS.json: the issue roof bench.py reports Blokus against
    fprintf(stderr, "{\"gpu\": \"%s\", \"sms\": %d, \"sm_mhz\": %.0f, \"issue_warp_inst_per_clk_per_sm\": 4, "
                    "\"issue_peak_t_warp_inst_s\": %.4f, \"lop3\": %.3f, \"shf\": %.3f, \"iadd\": %.3f, \"imad\": %.3f, "
                    "\"popc_lop3_iadd\": %.3f, \"lds_iadd_lop3\": %.3f, \"mix_warp_inst_per_clk_per_sm\": %.3f, "
                    "\"how\": \"tools/int_peak_probe.cu: long independent chains, 8 CTAs x 256 threads per SM, best of 5, CUDA events\"}\n",
            p.name, s, mhz, 4.0 * s * mhz * 1e6 / 1e12, g_rate[0], g_rate[1], g_rate[2], g_rate[3], g_rate[4], g_rate[5], g_rate[6]);
    return 0;
}
