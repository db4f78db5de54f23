// Shared helpers for the spaa_b200 kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#define SPAA_HD __host__ __device__ __forceinline__
#define SPAA_D __device__ __forceinline__
#else
#define SPAA_HD inline
#define SPAA_D inline
#endif

namespace spaa {

// error codes returned across the C ABI (0 = ok)
enum : int {
    SPAA_OK = 0,
    SPAA_ERR_ARG = -1,       // bad argument (null pointer, size, unsupported combination)
    SPAA_ERR_CUDA = -2,      // a CUDA runtime call / launch failed
    SPAA_ERR_UNSUPPORTED = -3
};

void set_last_error(const char* fmt, ...);

#if defined(__CUDACC__)

#define SPAA_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            spaa::set_last_error(__VA_ARGS__);    \
            return spaa::SPAA_ERR_ARG;            \
        }                                         \
    } while (0)

#define SPAA_CHECK_LAUNCH(name)                                                          \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            spaa::set_last_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return spaa::SPAA_ERR_CUDA;                                                  \
        }                                                                                \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// Dynamic shared memory opt-in of one kernel, tracked PER DEVICE and under a mutex: cudaFuncSetAttribute applies to the current device
// only, so a process that drives several GPUs (nn.DataParallel callers, device='cuda:1' with current device 0) must opt in on each.
// Usage: `static SmemOptIn opt; if (!opt.ensure(kernel, bytes, carveout)) error;` right before the launch.
struct SmemOptIn {
    static constexpr int kMaxDevices = 64;
    size_t reserved[kMaxDevices];
    int lock;
    template <class K> bool ensure(K kernel, size_t bytes, bool max_carveout = false) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
        while (__atomic_exchange_n(&lock, 1, __ATOMIC_ACQUIRE)) {}
        bool ok = true;
        if (bytes > reserved[dev]) {
            ok = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
            if (ok && max_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (ok) reserved[dev] = bytes;
        }
        __atomic_store_n(&lock, 0, __ATOMIC_RELEASE);
        return ok;
    }
};

SPAA_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` is >= 32 floats of shared memory.
SPAA_D float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (wid == 0) v = warp_sum(v);
    return v;
}


// Deterministic two-stage per-row reduction of N values: every block of row `row` (gridDim.x blocks per row) calls this
// with its thread-local partial sums; block partials go to `partial` and the last block to arrive adds them in a fixed
// order into out[row*N + i].  `counter` (one unsigned per row) must be zero on entry and is reset on exit, so the
// workspace can be reused by the next launch on the same stream.  Layout of a workspace made by row_reduce_ws_bytes:
// [rows * nblk * N floats][rows unsigned].
template <int N>
SPAA_D void row_reduce_finish(float (&v)[N], int row, float* __restrict__ partial, unsigned* __restrict__ counter,
                              float* __restrict__ out, float* red /* >=32 floats smem */, bool* is_last /* smem */) {
    const int nblk = gridDim.x;
    float* pp = partial + ((int64_t)row * nblk + blockIdx.x) * N;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float s = block_sum(v[i], red);
        if (threadIdx.x == 0) pp[i] = s;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(counter + row, 1u);
        *is_last = (t == (unsigned)nblk - 1);
    }
    __syncthreads();
    if (*is_last && threadIdx.x < 32) {
        __threadfence();
        const volatile float* q = partial + (int64_t)row * nblk * N;
        float a[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = 0.f;
        for (int k = threadIdx.x; k < nblk; k += 32)
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] += q[k * N + i];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            a[i] = warp_sum(a[i]);
            if (threadIdx.x == 0) out[row * N + i] = a[i];
        }
        if (threadIdx.x == 0) counter[row] = 0u;
    }
}
inline int row_reduce_nblk(int64_t n, int per_block) {
    int64_t k = (n + per_block - 1) / per_block;
    return (int)(k < 1 ? 1 : (k > 256 ? 256 : k));
}
inline int64_t row_reduce_ws_bytes(int64_t rows, int nblk, int N) {
    return rows * nblk * N * (int64_t)sizeof(float) + rows * (int64_t)sizeof(unsigned);
}

SPAA_D float ld_f(const float* p) { return __ldg(p); }
SPAA_D float ld_f(const __nv_bfloat16* p) { return __bfloat162float(*p); }
SPAA_D void st_f(float* p, float v) { *p = v; }
SPAA_D void st_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
SPAA_D float ld_f(const __half* p) { return __half2float(*p); }
SPAA_D void st_f(__half* p, float v) { *p = __float2half_rn(v); }

#endif  // __CUDACC__

}  // namespace spaa
