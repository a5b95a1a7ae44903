// util.cu -- error plumbing, device checks and the boundary pack/unpack kernels (K0).
#include <stdarg.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define TSR_HOST_SSE2 1
#endif

#include "common.cuh"

namespace tsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return TSR_ERR_CUDA;
    }
    return TSR_OK;
}

int sm_count() {
    static int cache[64];                       // per device ordinal; 0 = not queried yet
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64) cache[dev] = n;
    return n;
}

int require_sm100(const char* what) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        set_error("%s: no CUDA device available (this library has no CPU fallback)", what);
        return TSR_ERR_UNSUPPORTED;
    }
    if (major != 10) {
        set_error("%s: device compute capability %d.x is not sm_100 (B200 only, no fallback)", what, major);
        return TSR_ERR_UNSUPPORTED;
    }
    return TSR_OK;
}

// ----------------------------------------------------------------------------- K0 kernels
// One thread per (b,t,f): gathers E edge bytes into one packed byte.  Reads are E-byte
// contiguous per thread (8 B vector when E == 8), writes are 1 B coalesced.
__global__ void pack_edges_kernel(const uint8_t* __restrict__ X, uint8_t* __restrict__ bits, long long n, int E) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b = 0;
    if (E == 8) {
        uint2 v = *reinterpret_cast<const uint2*>(X + i * 8);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            b |= (((v.x >> (8 * e)) & 0xFF) != 0) << e;
            b |= (((v.y >> (8 * e)) & 0xFF) != 0) << (e + 4);
        }
    } else {
        for (int e = 0; e < E; ++e) b |= (X[i * E + e] != 0) << e;
    }
    bits[i] = (uint8_t)b;
}

__global__ void unpack_edges_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ X, long long n, int E) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b = bits[i];
    if (E == 8) {
        uint2 v;
        v.x = (b & 1) | ((b >> 1 & 1) << 8) | ((b >> 2 & 1) << 16) | ((b >> 3 & 1) << 24);
        v.y = (b >> 4 & 1) | ((b >> 5 & 1) << 8) | ((b >> 6 & 1) << 16) | ((b >> 7 & 1) << 24);
        *reinterpret_cast<uint2*>(X + i * 8) = v;
    } else {
        for (int e = 0; e < E; ++e) X[i * E + e] = (b >> e) & 1;
    }
}

// One warp per 32-feature word group: lane l tests feature 32*w + l, ballot packs the word.
__global__ void pack_features_kernel(const uint8_t* __restrict__ Fm, uint32_t* __restrict__ bits, long long R, int D,
                                     int rowwords) {
    const int lane = threadIdx.x & 31;
    long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long total = R * rowwords;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; warp < total; warp += nwarps) {
        long long r = warp / rowwords;
        int w = (int)(warp % rowwords);
        int d = w * 32 + lane;
        bool on = d < D && Fm[r * D + d] != 0;
        uint32_t word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) bits[warp] = word;
    }
}

__global__ void unpack_features_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ Fm, long long R, int D,
                                       int rowwords) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = R * D;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        long long r = i / D;
        int d = (int)(i % D);
        Fm[i] = (bits[r * rowwords + (d >> 5)] >> (d & 31)) & 1;
    }
}

// ----------------------------------------------------------------------------- host-side packer
// The boundary conversion for callers whose edge tensors live in HOST memory: packing before the copy moves 1 byte
// per (t,f) over PCIe instead of 8.  16 input bytes (two cells) -> compare with zero -> byte mask -> two packed bytes.
extern "C" int tsr_hostpack_has_avx512(void);                                                   // host_pack.cpp
extern "C" void tsr_hostpack_range_avx512(const uint8_t* X, uint8_t* bits, long long c0, long long c1);
static void host_pack_range(const uint8_t* X, uint8_t* bits, long long c0, long long c1) {
    static const bool avx512 = tsr_hostpack_has_avx512() != 0;
    if (avx512) { tsr_hostpack_range_avx512(X, bits, c0, c1); return; }
    long long i = c0;
#ifdef TSR_HOST_SSE2
    const __m128i zero = _mm_setzero_si128();
    for (; i + 8 <= c1; i += 8) {                      // 64 input bytes -> 8 packed bytes
        uint64_t out = 0;
#pragma GCC unroll 4
        for (int k = 0; k < 4; ++k) {
            const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(X + (i + 2 * k) * 8));
            const uint32_t nz = ~(uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, zero)) & 0xFFFFu;
            out |= (uint64_t)nz << (16 * k);
        }
        memcpy(bits + i, &out, 8);
    }
#endif
    for (; i < c1; ++i) {
        uint32_t b = 0;
        for (int e = 0; e < 8; ++e) b |= (uint32_t)(X[i * 8 + e] != 0) << e;
        bits[i] = (uint8_t)b;
    }
}

// Persistent worker pool (created on first use): a call publishes a job, wakes the workers and waits for them.
class HostPool {
public:
    void run(const uint8_t* X, uint8_t* bits, long long n, int nthreads) {
        std::unique_lock<std::mutex> call_lock(call_mu_);              // one job at a time
        if (nthreads < 1) nthreads = 1;
        if (nthreads > 256) nthreads = 256;
        {
            std::unique_lock<std::mutex> lk(mu_);
            while ((int)workers_.size() < nthreads - 1) {
                const int id = (int)workers_.size();
                workers_.emplace_back([this, id] { loop(id); });
            }
            X_ = X; bits_ = bits; n_ = n; parts_ = nthreads; pending_ = nthreads - 1; ++gen_;
        }
        cv_.notify_all();
        slice(nthreads - 1);                                           // the caller takes the last slice
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void slice(int id) {
        const long long per = ((n_ + parts_ - 1) / parts_ + 7) & ~7LL;   // 8-cell granules
        long long c0 = per * id, c1 = c0 + per;
        if (c0 > n_) c0 = n_;
        if (c1 > n_) c1 = n_;
        host_pack_range(X_, bits_, c0, c1);
    }
    void loop(int id) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (id >= parts_ - 1) continue;                        // not part of this job
            }
            slice(id);
            std::unique_lock<std::mutex> lk(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> workers_;
    const uint8_t* X_ = nullptr;
    uint8_t* bits_ = nullptr;
    long long n_ = 0;
    int parts_ = 1, pending_ = 0;
    unsigned long long gen_ = 0;
};

}  // namespace tsr

using namespace tsr;

extern "C" {

int tsr_host_pack_edges(const uint8_t* X_host, uint8_t* bits_host, long long n_cells, int E, int nthreads) {
    TSR_REQUIRE(X_host && bits_host, "tsr_host_pack_edges: null pointer");
    TSR_REQUIRE(n_cells >= 0 && E == 8, "tsr_host_pack_edges: needs E == 8 orientation bytes per cell (got %d)", E);
    static HostPool* pool = new HostPool();                           // never destroyed: workers outlive exit handlers
    if (n_cells < 65536 || nthreads <= 1) host_pack_range(X_host, bits_host, 0, n_cells);
    else pool->run(X_host, bits_host, n_cells, nthreads);
    return TSR_OK;
}

int tsr_abi_version(void) { return TSR_ABI_VERSION; }
const char* tsr_last_error(void) { return g_err; }
int tsr_device_supported(void) { return require_sm100("tsr_device_supported") == TSR_OK ? 1 : 0; }
int tsr_rowwords(int D) { return ((D + 63) / 64) * 2; }

int tsr_pack_edges(const uint8_t* X, uint8_t* bits, int B, int T, int F, int E, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(X && bits, "tsr_pack_edges: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 0 && F >= 0 && E >= 1 && E <= 8, "tsr_pack_edges: need 1 <= E <= 8 (got %d)", E);
    long long n = (long long)B * T * F;
    if (n == 0) return TSR_OK;
    int rc = require_sm100("tsr_pack_edges");
    if (rc) return rc;
    pack_edges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, bits, n, E);
    return check_launch("tsr_pack_edges");
}

int tsr_unpack_edges(const uint8_t* bits, uint8_t* X, int B, int T, int F, int E, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(X && bits, "tsr_unpack_edges: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 0 && F >= 0 && E >= 1 && E <= 8, "tsr_unpack_edges: need 1 <= E <= 8 (got %d)", E);
    long long n = (long long)B * T * F;
    if (n == 0) return TSR_OK;
    int rc = require_sm100("tsr_unpack_edges");
    if (rc) return rc;
    unpack_edges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, X, n, E);
    return check_launch("tsr_unpack_edges");
}

int tsr_pack_features_u8(const uint8_t* Fm, uint32_t* bits, long long R, int D, int rowwords, void* stream) {
    TSR_REQUIRE(Fm && bits, "tsr_pack_features_u8: null pointer");
    TSR_REQUIRE(R >= 0 && D >= 1 && rowwords == tsr_rowwords(D), "tsr_pack_features_u8: rowwords must be tsr_rowwords(D)");
    if (R == 0) return TSR_OK;
    int rc = require_sm100("tsr_pack_features_u8");
    if (rc) return rc;
    long long warps = R * rowwords;
    long long blocks = (warps + 7) / 8;
    if (blocks > (long long)sm_count() * 64) blocks = (long long)sm_count() * 64;
    pack_features_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(Fm, bits, R, D, rowwords);
    return check_launch("tsr_pack_features_u8");
}

int tsr_unpack_features_u8(const uint32_t* bits, uint8_t* Fm, long long R, int D, int rowwords, void* stream) {
    TSR_REQUIRE(Fm && bits, "tsr_unpack_features_u8: null pointer");
    TSR_REQUIRE(R >= 0 && D >= 1 && rowwords == tsr_rowwords(D), "tsr_unpack_features_u8: rowwords must be tsr_rowwords(D)");
    if (R == 0) return TSR_OK;
    int rc = require_sm100("tsr_unpack_features_u8");
    if (rc) return rc;
    long long n = R * D;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 64) blocks = (long long)sm_count() * 64;
    unpack_features_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, Fm, R, D, rowwords);
    return check_launch("tsr_unpack_features_u8");
}

}  // extern "C"
