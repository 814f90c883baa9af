// common.cuh -- error handling, device buffers, deterministic reductions.
// Part of libnekb200.so (unity build: see nekb200.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace nekb {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const char *file, int line, const std::string &msg)
{
    char buf[1024];
    snprintf(buf, sizeof buf, "%s:%d %s", file, line, msg.c_str());
    throw Error(buf);
}

#define NEKB_CUDA(call)                                                      \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) ::nekb::fail(__FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

#define NEKB_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) ::nekb::fail(__FILE__, __LINE__, (msg));     \
    } while (0)

// Kernel launches issued by the library (bench.py reports them as gpu_launches).
inline int64_t &launch_counter()
{
    static int64_t c = 0;
    return c;
}
#define NEKB_LAUNCHED()                        \
    do {                                       \
        ++::nekb::launch_counter();            \
        NEKB_CUDA(cudaPeekAtLastError());      \
    } while (0)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) {
            release();
            p = o.p, n = o.n;
            o.p = nullptr, o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr, n = 0;
    }
    // (Re)allocate to exactly `count` elements unless already that size.
    void alloc(size_t count)
    {
        if (count == n && p) return;
        release();
        if (count) NEKB_CUDA(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count)
    {
        if (count > n) alloc(count);
    }
    void zero(cudaStream_t s)
    {
        if (n) NEKB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void upload(const T *host, size_t count, cudaStream_t s)
    {
        alloc(count);
        if (count) NEKB_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *host, size_t count, cudaStream_t s) const
    {
        NEKB_REQUIRE(count <= n, "download larger than buffer");
        if (count) NEKB_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
    }
};

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum with a fixed combination tree (deterministic for a given block size).
// `red` is shared scratch of >= 32 doubles.  Result valid in thread 0.
template <bool IS_MAX = false>
__device__ __forceinline__ double block_reduce(double v, double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();  // protect `red` against a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < nw) ? red[lane] : (IS_MAX ? -1.0e300 : 0.0);
        r = IS_MAX ? warp_max(r) : warp_sum(r);
    }
    return r;
}

// Grid-wide reduction: every block deposits its partial; the last block to arrive combines all
// partials in index order (fixed tree => run-to-run deterministic, independent of which block is
// last) and calls `fin(total)` on thread 0.  `counter` must be 0 on entry and is reset on exit.
// All threads of the block must call.  `red` >= 33 doubles of shared scratch.
template <bool IS_MAX = false, class Fin>
__device__ __forceinline__ void grid_reduce(double block_val, double *partials, unsigned *counter, double *red,
                                            Fin fin)
{
    __shared__ int s_last;
    __syncthreads();  // a previous grid_reduce in the same kernel may still be reading s_last
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_val;
        __threadfence();
        unsigned t = atomicInc(counter, gridDim.x - 1);  // wraps back to 0 after the last arrival
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double a = IS_MAX ? -1.0e300 : 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            double q = __ldcg(partials + i);
            a = IS_MAX ? fmax(a, q) : a + q;
        }
        double tot = block_reduce<IS_MAX>(a, red);
        if (threadIdx.x == 0) fin(tot);
    }
}

}  // namespace nekb
