// ubench_dmma.cu -- what do the FP64 pipes and the shared-memory pipe of one B200 SM deliver?  Decides whether the two
// in-plane contractions of the operator kernel (D u and u D^T on 8x8 planes) belong on mma.sync.m8n8k4.f64 (DMMA):
//   * DFMA and DMMA issue rates at 2 / 4 warps per SM sub-partition with 8 independent accumulators,
//   * cycles per LDS.64 / LDS.128 / STS.64 for the access patterns the operator kernel uses (row broadcast, column,
//     fragment loads).
// Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_dmma scripts/ubench_dmma.cu && /tmp/ubench_dmma
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int ITERS = 4096;

// MODE 0: DFMA, 16 independent chains per thread.  MODE 1: DMMA, 8 independent accumulator tiles per warp.
template <int MODE>
__global__ void k_math(double *out, long long *cyc, double seed)
{
    double c[16];
#pragma unroll
    for (int q = 0; q < 16; q++) c[q] = seed * (q + threadIdx.x);
    const double a = seed + 1e-9 * threadIdx.x, b = seed - 1e-9 * threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 16; q++) c[q] = fma(a, c[q], b);
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) dmma(c[2 * q], c[2 * q + 1], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; q++) s += c[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory patterns on an 8x8x8 tile of doubles per warp (4 KB), one warp = lanes (i = lane & 7, jj = lane >> 3)
// PAT 0: LDS.64 column  t[k][m][i]           (8 distinct words per warp: broadcast over jj)
// PAT 1: LDS.128 row    t[k][jj][2m..2m+1]   (4 distinct 16-byte words per warp)
// PAT 2: LDS.64 all distinct, contiguous 256 B  t[k][m4 + (lane>>3)][lane & 7]
// PAT 3: LDS.128 all distinct, contiguous 512 B
// PAT 4: STS.64 all distinct contiguous
// PAT 5: LDS.64 row broadcast t[k][jj][m] (4 distinct words)
template <int PAT>
__global__ void k_lds(double *out, long long *cyc)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *t = sm + warp * 512;
    for (int q = lane; q < 512; q += 32) t[q] = q * 0.5 + warp;
    __syncthreads();
    const int i = lane & 7, jj = lane >> 3;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                if (PAT == 0) acc += *(volatile double *)&t[k * 64 + m * 8 + i];
                if (PAT == 5) acc += *(volatile double *)&t[k * 64 + jj * 8 + m];
                if (PAT == 2) acc += *(volatile double *)&t[k * 64 + ((m & 1) * 4 + jj) * 8 + i];
                if (PAT == 4) *(volatile double *)&t[k * 64 + ((m & 1) * 4 + jj) * 8 + i] = acc + m;
                if (PAT == 1 && m < 4) {
                    double2 v;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(&t[k * 64 + jj * 8 + 2 * m])));
                    acc += v.x + v.y;
                }
                if (PAT == 3 && m < 4) {
                    double2 v;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(&t[((k * 4 + m) & 7) * 64 + lane * 2])));
                    acc += v.x + v.y;
                }
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out;
    long long *cyc, h_cyc[1024];
    CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    printf("device %s, %d SMs\n", prop.name, sms);
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0)
                    k_math<0><<<sms, warps * 32>>>(out, cyc, 1.0000001);
                else
                    k_math<1><<<sms, warps * 32>>>(out, cyc, 1.0000001);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h_cyc, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int q = 0; q < sms; q++) mean += h_cyc[q];
            mean /= sms;
            const double instr = (double)ITERS * (mode == 0 ? 16 : 8) * warps;   // warp instructions per SM
            const double flop = instr * (mode == 0 ? 64.0 : 512.0);
            printf("%s  warps/SM %2d : %.3f warp-instr/clk/SM, %.1f flop/clk/SM  (x %d SMs x 1.9 GHz = %.1f TFLOP/s)\n",
                   mode == 0 ? "DFMA" : "DMMA", warps, instr / mean, flop / mean, sms, flop / mean * sms * 1.9e-3);
        }
    }
    const char *names[6] = {"LDS.64 column (8 distinct words, broadcast)", "LDS.128 row (4 distinct 16 B, broadcast)",
                            "LDS.64 256 B contiguous", "LDS.128 512 B contiguous", "STS.64 256 B contiguous",
                            "LDS.64 row (4 distinct words, broadcast)"};
    for (int warps : {2, 8}) {
        for (int pat = 0; pat < 6; pat++) {
            const size_t smem = (size_t)warps * 4096;
            for (int rep = 0; rep < 2; rep++) {
                switch (pat) {
                    case 0: k_lds<0><<<sms, warps * 32, smem>>>(out, cyc); break;
                    case 1: k_lds<1><<<sms, warps * 32, smem>>>(out, cyc); break;
                    case 2: k_lds<2><<<sms, warps * 32, smem>>>(out, cyc); break;
                    case 3: k_lds<3><<<sms, warps * 32, smem>>>(out, cyc); break;
                    case 4: k_lds<4><<<sms, warps * 32, smem>>>(out, cyc); break;
                    default: k_lds<5><<<sms, warps * 32, smem>>>(out, cyc); break;
                }
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h_cyc, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int q = 0; q < sms; q++) mean += h_cyc[q];
            mean /= sms;
            const double instr = (double)(ITERS / 8) * 8 * ((pat == 1 || pat == 3) ? 4 : 8) * warps;
            printf("warps/SM %d  %-46s : %.2f clk per warp instruction (SM-wide)\n", warps, names[pat], mean / instr);
        }
    }
    return 0;
}
