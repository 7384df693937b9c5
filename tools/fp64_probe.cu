// fp64 pipe probe for B200 (sm_100a): what is the best sustained DFMA rate, and what is the DFMA dependent-issue
// latency?  Not part of the product; results are recorded in profiles/ and DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu && tools/fp64_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__constant__ double cA = 0.999999, cB = 1e-9;
__constant__ double cT[8] = {0.999999, 0.999998, 0.999997, 0.999996, 0.999995, 0.999994, 0.999993, 0.999992};

template <int CH, int MODE>
__global__ void chain(double *out, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-9 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < (CH == 3 ? 42 : 128 / CH); ++u) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (MODE == 0) x[c] = fma(x[c], a, b);          // 3 register operands
                else if (MODE == 1) x[c] = fma(x[c], cA, cB);   // constant-bank operand(s)
                else if (MODE == 2) x[c] = fma(x[c], x[c], b);  // 2 distinct registers
                else if (MODE == 3) x[c] = fma(x[c], cT[(c + u) & 7], b);  // constant-bank operand + 2 registers
                else if (MODE == 4) x[c] = x[c] * a;            // DMUL
                else x[c] = x[c] + b;                           // DADD
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = s;
}

// Issue model: can other pipes issue in the shadow of a (2-cycle) fp64 warp instruction?  8 DFMA chains (2-register
// form) interleaved with K independent FP32/INT instructions per DFMA.
template <int K, int KIND>
__global__ void mix(double *out, int iters, double b, float fa, int ia) {
    double x[8];
    float f[8];
    int n[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { x[c] = threadIdx.x * 1e-9 + c; f[c] = threadIdx.x + c; n[c] = threadIdx.x + c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                x[c] = fma(x[c], x[c], b);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (KIND == 0) f[(c + k) & 7] = fmaf(f[(c + k) & 7], fa, 1.0f);
                    else if (KIND == 1) n[(c + k) & 7] = n[(c + k) & 7] * ia + 1;
                    else n[(c + k) & 7] = (n[(c + k) & 7] ^ ia) + k;
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c] + f[c] + n[c];
    out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = s;
}

template <int K, int KIND>
static void runmix(const char *name, double *d) {
    const int iters = 2048, threads = 256, blocks = 148 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) mix<K, KIND><<<blocks, threads>>>(d, iters, 1e-9, 0.999f, 3);
    float best = 1e30f;
    for (int r = 0; r < 10; ++r) {
        cudaEventRecord(e0);
        mix<K, KIND><<<blocks, threads>>>(d, iters, 1e-9, 0.999f, 3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double dfma = 128.0 * iters * threads * blocks;
    printf("%-40s %.3f ms  %.2f TFLOP/s fp64  (%.3f DFMA warp-inst/clk/SM)\n", name, best, 2 * dfma / (best * 1e-3) / 1e12,
           dfma / 32 / (best * 1e-3) / 1.965e9 / 148);
}

__global__ void latency(double *out, long long *cyc, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 64; ++u) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int CH, int MODE>
static void run(const char *name, int threads, int blocks_per_sm, double *d, int nsm = 0) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2048;
    const int blocks = nsm ? nsm : sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) chain<CH, MODE><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    float best = 1e30f;
    for (int r = 0; r < 10; ++r) {
        cudaEventRecord(e0);
        chain<CH, MODE><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 128 * (double)iters * threads * blocks;
    printf("%-28s threads %4d x %4d blocks: %.3f ms  %.2f TFLOP/s  = %.3f warp-inst/clk/SM-equivalent (@1965 MHz, %d blocks)\n", name, threads, blocks,
           best, flops / (best * 1e-3) / 1e12, flops / 2 / 32 / (best * 1e-3) / 1.965e9 / (nsm ? (nsm < sms ? nsm : sms) : sms), blocks);
}

int main() {
    double *d;
    long long *cyc;
    cudaMalloc(&d, sizeof(double) * 1024 * 148 * 16);
    cudaMalloc(&cyc, 8);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clockRate %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    run<8, 0>("8 chains, reg operands", 256, 8, d);
    run<8, 1>("8 chains, const operands", 256, 8, d);
    run<8, 2>("8 chains, x*x+b", 256, 8, d);
    run<4, 0>("4 chains, reg operands", 256, 8, d);
    run<16, 0>("16 chains, reg operands", 256, 4, d);
    run<16, 1>("16 chains, const operands", 256, 4, d);
    run<8, 0>("8 chains, reg operands", 128, 4, d);
    run<8, 0>("8 chains, reg operands", 128, 2, d);
    run<8, 0>("8 chains, reg operands", 128, 1, d);
    run<2, 0>("2 chains, reg operands", 128, 5, d);
    run<1, 0>("1 chain, reg operands", 128, 5, d);
    run<2, 1>("2 chains, const operands", 128, 5, d);
    run<4, 1>("4 chains, const operands", 128, 5, d);
    run<1, 0>("1 chain, reg operands", 1024, 2, d);
    run<2, 0>("2 chains, reg operands", 1024, 2, d);
    run<8, 3>("8 chains, const-bank operand", 256, 8, d);
    run<4, 3>("4 chains, const-bank operand", 128, 5, d);
    run<8, 4>("8 chains, DMUL", 256, 8, d);
    run<8, 5>("8 chains, DADD", 256, 8, d);
    // does TLP hide the dependent-issue latency?  1 chain per warp, varying warps per SM and number of busy SMs
    run<1, 0>("1 chain  1 block total", 512, 1, d, 1);
    run<1, 0>("1 chain  1 block total", 1024, 1, d, 1);
    run<1, 0>("1 chain  2 blocks total", 512, 1, d, 2);
    run<1, 0>("1 chain  74 blocks", 512, 1, d, 74);
    run<1, 0>("1 chain  148 blocks", 512, 1, d, 148);
    run<1, 0>("1 chain  148 blocks", 1024, 1, d, 148);
    run<1, 0>("1 chain  296 blocks", 512, 1, d, 296);
    run<2, 0>("2 chains 148 blocks", 512, 1, d, 148);
    run<2, 2>("2 chains x*x+b 148 blocks", 512, 1, d, 148);
    run<1, 2>("1 chain x*x+b 148 blocks", 512, 1, d, 148);
    run<1, 2>("1 chain x*x+b 148 blocks", 1024, 1, d, 148);
    run<4, 2>("4 chains x*x+b", 128, 5, d);
    run<2, 2>("2 chains x*x+b", 128, 5, d);
    run<3, 2>("3 chains x*x+b", 128, 5, d);
    runmix<0, 0>("DFMA only (32 warps/SM)", d);
    runmix<1, 0>("DFMA + 1 FFMA each", d);
    runmix<2, 0>("DFMA + 2 FFMA each", d);
    runmix<3, 0>("DFMA + 3 FFMA each", d);
    runmix<1, 1>("DFMA + 1 IMAD each", d);
    runmix<2, 1>("DFMA + 2 IMAD each", d);
    runmix<1, 2>("DFMA + 1 (LOP3,IADD) each", d);
    // dependent-issue latency: one warp, one chain
    latency<<<1, 32>>>(d, cyc, 1000, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency: %.2f cycles (1 warp)\n", (double)c / (1000.0 * 64));
    latency<<<1, 128>>>(d, cyc, 1000, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA, 4 warps (1 per SMSP): %.2f cycles per DFMA per warp\n", (double)c / (1000.0 * 64));
    latency<<<1, 256>>>(d, cyc, 1000, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA, 8 warps (2 per SMSP): %.2f cycles per DFMA per warp\n", (double)c / (1000.0 * 64));
    latency<<<1, 512>>>(d, cyc, 1000, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA, 16 warps (4 per SMSP): %.2f cycles per DFMA per warp\n", (double)c / (1000.0 * 64));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
