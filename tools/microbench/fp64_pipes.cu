// Microbenchmark of the FP64 paths of one SM (B200): DMMA m8n8k4 throughput / latency, DFMA throughput / latency, whether they
// share a pipe, double<->float conversions, 64-bit shared-memory loads.   nvcc -arch=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i + 0.5f;
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 8 independent DMMA chains
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        } else if (MODE == 1) {     // 1 dependent DMMA chain, 8 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[0], c[1], a, b);
        } else if (MODE == 2) {     // 16 independent DFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        } else if (MODE == 3) {     // dependent DFMA chain, 16 per iteration
#pragma unroll
            for (int i = 0; i < 16; ++i) c[0] = fma(c[0], a, b);
        } else if (MODE == 4) {     // 8 DMMA + 16 DFMA independent
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
            for (int i = 8; i < 16; ++i) c[i] = fma(c[i], a, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
            for (int i = 8; i < 16; ++i) c[i] = fma(c[i], a, b);
        } else if (MODE == 5) {     // 8 double->float->double round trips (independent)
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = (double)((float)c[i] + 1.0f);
        } else if (MODE == 6) {     // 16 independent 64-bit shared loads (conflict-free) + add
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] += sm[(threadIdx.x + 32 * i + (int)c[15]) & 4095];
        } else if (MODE == 7) {     // dependent chain: DP mul -> rsqrtf seed -> 2 Newton steps (jb_rsqrt)
            double x = c[0] + 2.0;
            double y = (double)rsqrtf((float)x);
            const double h = 0.5 * x;
            y = y * (1.5 - h * y * y);
            y = y * (1.5 - h * y * y);
            c[0] = y;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, sizeof(long long) * 148);
    const int iters = 2000;
    k<MODE><<<148, threads, 32768>>>(out, cyc, 10);
    k<MODE><<<148, threads, 32768>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const int warps = threads / 32;
    printf("%-44s warps/SM %2d: %8.2f cycles per warp-instr-slot, %7.2f cycles/instr/SM (all warps)\n", name, warps,
           avg / iters / per_iter, avg / iters / per_iter / warps);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {32, 128, 256, 576, 1024}) {
        run<0>("DMMA independent x8", 8, th);
        run<1>("DMMA dependent chain", 8, th);
        run<2>("DFMA independent x16", 16, th);
        run<3>("DFMA dependent chain", 16, th);
        run<4>("8 DMMA + 16 DFMA mixed (per 24 instr)", 24, th);
        run<5>("F2F d->f->d (+FADD) x8 (per round trip)", 8, th);
        run<6>("LDS.64 x16 + DADD", 16, th);
        run<7>("jb_rsqrt chain (per call)", 1, th);
    }
    return 0;
}
