// ubench_stream.cu -- micro-benchmark of the CG "update" streaming pattern (r -= alpha * mask * w ; s = sum wt r^2):
// 2 reads + 1 write of 8-byte words + one code byte per node, n = 512 * 262,144 nodes.  Which form gets closest to the copy
// bandwidth?  Variants: grid-stride loads/stores through registers (block size, registers, unroll, cache hints) against a
// TMA (cp.async.bulk) ring with bulk stores.  Build + run (GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_stream scripts/ubench_stream.cu && /tmp/ubench_stream
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void block_out(double s, double *out)
{
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[blockIdx.x] = v;
    }
}

__device__ __forceinline__ double upd(double r, double a, unsigned char c, double alpha, double &s)
{
    const double rn = fma(-alpha, (c & 0x80) ? 0.0 : a, r);
    s = fma((double)(c & 0x7f) * rn, rn, s);
    return rn;
}

// V0: the library's form -- one quad per thread per trip
template <int HINT>
__global__ void k_quad(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code, int64_t n,
                       double alpha, double *out)
{
    double s = 0.0;
    const int64_t n4 = n >> 2;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        double2 ra, rb, aa, ab;
        uchar4 c;
        if (HINT) {
            ra = __ldcs(r2 + 2 * t), rb = __ldcs(r2 + 2 * t + 1), aa = __ldcs(a2 + 2 * t), ab = __ldcs(a2 + 2 * t + 1);
            c = __ldcs(c4 + t);
        } else {
            ra = r2[2 * t], rb = r2[2 * t + 1], aa = a2[2 * t], ab = a2[2 * t + 1];
            c = c4[t];
        }
        ra.x = upd(ra.x, aa.x, c.x, alpha, s), ra.y = upd(ra.y, aa.y, c.y, alpha, s);
        rb.x = upd(rb.x, ab.x, c.z, alpha, s), rb.y = upd(rb.y, ab.y, c.w, alpha, s);
        if (HINT) {
            __stcs(r2 + 2 * t, ra), __stcs(r2 + 2 * t + 1, rb);
        } else {
            r2[2 * t] = ra, r2[2 * t + 1] = rb;
        }
    }
    block_out(s, out);
}

// V2: a thread owns double2 lanes strided by the block (fully coalesced 16-byte accesses), UNR independent pairs per trip
template <int UNR>
__global__ void k_lane(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code, int64_t n,
                       double alpha, double *out)
{
    double s = 0.0;
    const int64_t n2 = n >> 1;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar2 *c2 = reinterpret_cast<const uchar2 *>(code);
    const int64_t chunk = (int64_t)blockDim.x * UNR;
    for (int64_t base = blockIdx.x * chunk; base < n2; base += (int64_t)gridDim.x * chunk) {
        double2 rv[UNR], av[UNR];
        uchar2 cv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int64_t t = base + u * blockDim.x + threadIdx.x;
            if (t < n2) rv[u] = r2[t], av[u] = a2[t], cv[u] = c2[t];
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int64_t t = base + u * blockDim.x + threadIdx.x;
            if (t < n2) {
                rv[u].x = upd(rv[u].x, av[u].x, cv[u].x, alpha, s), rv[u].y = upd(rv[u].y, av[u].y, cv[u].y, alpha, s);
                r2[t] = rv[u];
            }
        }
    }
    block_out(s, out);
}

// V5: TMA ring.  A stage = TILE doubles of r and of ap + TILE code bytes, fetched with cp.async.bulk by one thread; the block
// updates the stage in shared memory and sends r back with a bulk store.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int TILE, int STAGES, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
    k_tma(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code, int64_t n, double alpha,
          double *out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sr = reinterpret_cast<double *>(smem_raw);                      // [STAGES][TILE]
    double *sa = sr + (size_t)STAGES * TILE;                                // [STAGES][TILE]
    unsigned char *sc = reinterpret_cast<unsigned char *>(sa + (size_t)STAGES * TILE);   // [STAGES][TILE]
    uint64_t *full = reinterpret_cast<uint64_t *>(sc + (size_t)STAGES * TILE);
    if (threadIdx.x == 0) {
        for (int q = 0; q < STAGES; q++) mbar_init(&full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = n / TILE;
    auto issue = [&](int stage, int64_t tile) {
        mbar_expect_tx(&full[stage], TILE * 17);
        bulk_g2s(sr + (size_t)stage * TILE, r + tile * TILE, TILE * 8, &full[stage]);
        bulk_g2s(sa + (size_t)stage * TILE, ap + tile * TILE, TILE * 8, &full[stage]);
        bulk_g2s(sc + (size_t)stage * TILE, code + tile * TILE, TILE, &full[stage]);
    };
    if (threadIdx.x == 0)
        for (int q = 0; q < STAGES; q++) {
            const int64_t tile = blockIdx.x + (int64_t)q * gridDim.x;
            if (tile < ntiles) issue(q, tile);
        }
    double s = 0.0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int stage = it % STAGES;
        mbar_wait(&full[stage], (uint32_t)(it / STAGES) & 1u);
        double2 *r2 = reinterpret_cast<double2 *>(sr + (size_t)stage * TILE);
        const double2 *a2 = reinterpret_cast<const double2 *>(sa + (size_t)stage * TILE);
        const uchar2 *c2 = reinterpret_cast<const uchar2 *>(sc + (size_t)stage * TILE);
#pragma unroll
        for (int q = threadIdx.x; q < TILE / 2; q += THREADS) {
            double2 rv = r2[q];
            const double2 av = a2[q];
            const uchar2 c = c2[q];
            rv.x = upd(rv.x, av.x, c.x, alpha, s), rv.y = upd(rv.y, av.y, c.y, alpha, s);
            r2[q] = rv;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(r + tile * TILE, sr + (size_t)stage * TILE, TILE * 8);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            const int64_t nxt = tile + (int64_t)STAGES * gridDim.x;
            if (nxt < ntiles) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the stage: refill it
                issue(stage, nxt);
            }
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    block_out(s, out);
}

__global__ void k_copy(double2 *__restrict__ dst, const double2 *__restrict__ src, int64_t n2)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) dst[t] = src[t];
}

int main()
{
    const int64_t n = 512LL * 262144;
    double *r, *ap, *out;
    unsigned char *code;
    CK(cudaMalloc(&r, n * 8));
    CK(cudaMalloc(&ap, n * 8));
    CK(cudaMalloc(&code, n));
    CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMemset(r, 0, n * 8));
    CK(cudaMemset(ap, 0, n * 8));
    CK(cudaMemset(code, 1, n));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double bytes = 25.0 * n;   // 2 reads + 1 write of 8 B + 1 code byte
    auto time = [&](const char *name, auto launch, double b) {
        for (int w = 0; w < 3; w++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int w = 0; w < reps; w++) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("%-44s %8.4f ms  %7.1f GB/s\n", name, ms / reps, b / (ms / reps * 1e-3) / 1e9);
    };
    time("copy 16B/thread grid 148*8 (r+w 16 B/node)", [&] { k_copy<<<148 * 8, 256>>>((double2 *)r, (const double2 *)ap, n / 2); }, 16.0 * n);
    for (int g : {6, 8, 12, 16}) {
        char nm[96];
        snprintf(nm, sizeof nm, "quad/thread, 256 thr, %d CTA/SM", g);
        time(nm, [&] { k_quad<0><<<148 * g, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    }
    time("quad/thread, 256 thr, 8 CTA/SM, ldcs/stcs", [&] { k_quad<1><<<148 * 8, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("quad/thread, 512 thr, 4 CTA/SM", [&] { k_quad<0><<<148 * 4, 512>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("quad/thread, 1024 thr, 2 CTA/SM", [&] { k_quad<0><<<148 * 2, 1024>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("lane x1, 256 thr, 8 CTA/SM", [&] { k_lane<1><<<148 * 8, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("lane x2, 256 thr, 8 CTA/SM", [&] { k_lane<2><<<148 * 8, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("lane x4, 256 thr, 8 CTA/SM", [&] { k_lane<4><<<148 * 8, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("lane x4, 256 thr, 4 CTA/SM", [&] { k_lane<4><<<148 * 4, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    time("lane x8, 256 thr, 4 CTA/SM", [&] { k_lane<8><<<148 * 4, 256>>>(r, ap, code, n, 1e-3, out); }, bytes);
    {
        constexpr int TILE = 2048, ST = 6, TH = 256;
        const size_t sm = (size_t)ST * TILE * 17 + 64;
        CK(cudaFuncSetAttribute(k_tma<TILE, ST, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        time("TMA ring 2048 x 6 stages, 256 thr, 1 CTA/SM", [&] { k_tma<TILE, ST, TH><<<148, TH, sm>>>(r, ap, code, n, 1e-3, out); }, bytes);
    }
    {
        constexpr int TILE = 4096, ST = 3, TH = 512;
        const size_t sm = (size_t)ST * TILE * 17 + 64;
        CK(cudaFuncSetAttribute(k_tma<TILE, ST, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        time("TMA ring 4096 x 3 stages, 512 thr, 1 CTA/SM", [&] { k_tma<TILE, ST, TH><<<148, TH, sm>>>(r, ap, code, n, 1e-3, out); }, bytes);
    }
    {
        constexpr int TILE = 1024, ST = 12, TH = 256;
        const size_t sm = (size_t)ST * TILE * 17 + 128;
        CK(cudaFuncSetAttribute(k_tma<TILE, ST, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        time("TMA ring 1024 x 12 stages, 256 thr, 1 CTA/SM", [&] { k_tma<TILE, ST, TH><<<148, TH, sm>>>(r, ap, code, n, 1e-3, out); }, bytes);
    }
    return 0;
}
