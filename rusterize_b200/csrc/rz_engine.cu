// rz_engine.cu — device context, pipeline orchestration and the C ABI of librz_b200.so.
//
// Replaces the slab  rust/src/rasterize.rs:71-196 (ArrayBuilder::build + process)
//                  + rust/src/rasterization/*  + rust/src/encoding/writers.rs
// of the reference with:  host flattening (rz_host.cpp) -> CUDA kernels (rz_kernels.cuh) -> copy-back.
// There is no CPU fallback: without a usable CUDA device every compute entry point fails.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <cctype>
#include <unistd.h>
#include <chrono>
#include <future>
#include <thread>
#include <type_traits>
#include <vector>

#include "rz_host.hpp"
#include "rz_kernels.cuh"
#include "rz_sparse.cuh"
#include "rz_burn.cuh"
#include "rz_tiles.cuh"
#include "rz_dispatch.hpp"

namespace rz {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};
static void set_err(char* err, size_t n, const std::string& m) {
    if (err && n) {
        std::strncpy(err, m.c_str(), n - 1);
        err[n - 1] = 0;
    }
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            (void)cudaGetLastError();                                                                    \
            throw Error{RZ_RUNTIME_ERROR, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                              __FILE__ + ":" + std::to_string(__LINE__)};              \
        }                                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device memory
// ------------------------------------------------------------------------------------------------
static bool trim_device_pool_current();  // frees the pooled geometry blocks of the current device (defined below)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CUDA_TRY(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            if (cudaMalloc(&p, want) != cudaSuccess) {
                (void)cudaGetLastError();
                p = nullptr;
                trim_device_pool_current();
                CUDA_TRY(cudaMalloc(&p, want));
            }
        }
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

// ------------------------------------------------------------------------------------------------
// page-locked host blocks
// ------------------------------------------------------------------------------------------------
// Host memory for sparse results and for the vertex pools of geometry sets.  Handing out gigabytes of FRESH
// memory per call costs far more than the whole device pipeline (measured for 125 M triplets = 2.5 GB: 3.8 ms of
// kernels against 700 ms to fault the pages in and page-lock them for the copy).  Blocks are therefore huge-page
// backed, faulted in by several threads, page-locked once (portable: every device's copy engine may use them),
// and recycled through a pool when their owner is freed: steady-state calls copy straight into / out of
// resident, pinned memory at the PCIe rate.  RZ_HOST_POOL_BYTES caps what the pool keeps (default: a quarter of
// the machine's memory, between 8 and 64 GiB).
struct HostBlock {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    bool interleaved = false;  // pages placed on the memory nodes of the GPUs (rz_host_alloc, sparse results)
};

// Memory nodes the CUDA devices in use (RZ_DEVICES, else every visible one) hang off
// (/sys/bus/pci/devices/<bus id>/numa_node), distinct, ascending.
static const std::vector<int>& device_memory_nodes() {
    static const std::vector<int> nodes = []() {
        std::vector<int> v;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            (void)cudaGetLastError();
            return v;
        }
        std::vector<int> devs;
        if (const char* e = std::getenv("RZ_DEVICES")) {  // the devices calls use by default (same rule as the bindings)
            for (const char* q = e; *q;) {
                char* end = nullptr;
                const long d = std::strtol(q, &end, 10);
                if (end == q) break;
                if (d >= 0 && d < n) devs.push_back((int)d);
                q = *end == ',' ? end + 1 : end;
            }
        }
        if (devs.empty())
            for (int d = 0; d < n; d++) devs.push_back(d);
        for (int d : devs) {
            char bus[32] = {0};
            if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, d) != cudaSuccess) {
                (void)cudaGetLastError();
                continue;
            }
            for (char* q = bus; *q; q++) *q = (char)std::tolower((unsigned char)*q);
            const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
            std::FILE* f = std::fopen(path.c_str(), "r");
            if (!f) continue;
            int node = -1;
            if (std::fscanf(f, "%d", &node) != 1) node = -1;
            std::fclose(f);
            if (node >= 0 && node < 1024 && std::find(v.begin(), v.end(), node) == v.end()) v.push_back(node);
        }
        std::sort(v.begin(), v.end());
        return v;
    }();
    return nodes;
}

// Place the pages of [p, p+bytes) - before they are touched - on the memory nodes the visible GPUs are attached to:
// round-robin over them (mbind, MPOL_INTERLEAVE) when they hang off several nodes, on the one node otherwise
// (MPOL_PREFERRED).  A raster written by the copy engines of GPUs on BOTH sockets otherwise lives on the socket of
// the thread that first touched it, and the devices of the other socket copy into it at 60 % of the rate (measured
// at 8 GPUs: 170 ms against 105 ms for the same 2.1 GB; at 4 GPUs 161 against 94).  Returns false when the nodes
// are unknown or the kernel refuses (containers without CAP_SYS_NICE): the pages then follow first touch.
static bool place_pages_near_devices(void* p, size_t bytes) {
    if (const char* e = std::getenv("RZ_HOST_INTERLEAVE"))
        if (std::atoi(e) == 0) return false;
    const std::vector<int>& nodes = device_memory_nodes();
    if (nodes.empty()) return false;
    const unsigned long bits = 8 * sizeof(unsigned long);
    const unsigned long max_node = (unsigned long)nodes.back();
    std::vector<unsigned long> mask(max_node / bits + 2, 0ul);
    for (int k : nodes) mask[(unsigned long)k / bits] |= 1ul << ((unsigned long)k % bits);
    const int MPOL_PREFERRED_ = 1, MPOL_INTERLEAVE_ = 3;
    return syscall(SYS_mbind, p, bytes, nodes.size() > 1 ? MPOL_INTERLEAVE_ : MPOL_PREFERRED_, mask.data(), max_node + 2, 0u) == 0;
}
class HostPool {
  public:
    HostBlock get(size_t bytes, bool interleave = false) {
        if (bytes == 0) return HostBlock{};
        {
            std::lock_guard<std::mutex> lk(mu_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].cap >= bytes && free_[i].cap <= 2 * bytes + (1u << 20) && free_[i].interleaved == interleave &&
                    (best == free_.size() || free_[i].cap < free_[best].cap))
                    best = i;
            if (best != free_.size()) {
                HostBlock b = free_[best];
                free_.erase(free_.begin() + best);
                pooled_ -= b.cap;
                return b;
            }
        }
        HostBlock b;
        const size_t huge = (size_t)2 << 20;
        b.cap = bytes >= huge ? (bytes + huge - 1) & ~(huge - 1) : bytes;
        if (bytes >= huge) {
            if (posix_memalign(&b.p, huge, b.cap) != 0) throw std::bad_alloc();
            madvise(b.p, b.cap, MADV_HUGEPAGE);
            if (interleave) place_pages_near_devices(b.p, b.cap);  // (the flag is kept either way: it is the pool's key)
            b.interleaved = interleave;
            // first touch in parallel: the kernel clears the pages on the faulting thread
            const unsigned nt = std::min<unsigned>({std::max(1u, std::thread::hardware_concurrency()), 16u,
                                                   (unsigned)(b.cap >> 26) + 1u});
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nt; t++)
                th.emplace_back([&, t]() {
                    char* q = (char*)b.p;
                    const size_t lo = b.cap / nt * t, hi = t + 1 == nt ? b.cap : b.cap / nt * (t + 1);
                    for (size_t o = lo; o < hi; o += 4096) q[o] = 0;
                });
            for (auto& x : th) x.join();
            if (cudaHostRegister(b.p, b.cap, cudaHostRegisterPortable) == cudaSuccess) b.pinned = true;
            else (void)cudaGetLastError();
        } else {
            b.p = std::malloc(b.cap);
            if (!b.p) throw std::bad_alloc();
        }
        return b;
    }
    void put(HostBlock b) {
        if (!b.p) return;
        std::vector<HostBlock> evicted;
        bool kept = false;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (b.cap >= ((size_t)2 << 20) && b.cap <= limit()) {
                // a full pool gives up its oldest blocks (free_ is in order of return): the block handed back last is
                // the one the next call of the same shape asks for
                while (pooled_ + b.cap > limit() && !free_.empty()) {
                    evicted.push_back(free_.front());
                    pooled_ -= free_.front().cap;
                    free_.erase(free_.begin());
                }
                free_.push_back(b);
                pooled_ += b.cap;
                kept = true;
            }
        }
        for (auto& e : evicted) release(e);
        if (!kept) release(b);
    }
    // give the free blocks back to the system until at most keep_bytes stay pooled (oldest first)
    void trim(size_t keep_bytes) {
        std::vector<HostBlock> evicted;
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (pooled_ > keep_bytes && !free_.empty()) {
                evicted.push_back(free_.front());
                pooled_ -= free_.front().cap;
                free_.erase(free_.begin());
            }
        }
        for (auto& e : evicted) release(e);
    }
    size_t pooled() {
        std::lock_guard<std::mutex> lk(mu_);
        return pooled_;
    }
    // blocks lent to a container that only knows the pointer (the vertex pools' allocator)
    void* lease(size_t bytes, bool interleave = false) {
        HostBlock b = get(bytes, interleave);
        std::lock_guard<std::mutex> lk(mu_);
        leased_[b.p] = b;
        return b.p;
    }
    bool unlease(void* p) {
        HostBlock b;
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto it = leased_.find(p);
            if (it == leased_.end()) return false;
            b = it->second;
            leased_.erase(it);
        }
        put(b);
        return true;
    }
    bool leased(const void* p, HostBlock* out) {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = leased_.find(const_cast<void*>(p));
        if (it == leased_.end()) return false;
        if (out) *out = it->second;
        return true;
    }

  private:
    static void release(HostBlock& b) {
        if (b.pinned && cudaHostUnregister(b.p) != cudaSuccess) (void)cudaGetLastError();
        std::free(b.p);
        b.p = nullptr;
    }
    static size_t limit() {
        if (const char* e = std::getenv("RZ_HOST_POOL_BYTES")) return (size_t)std::strtoull(e, nullptr, 10);
        // a quarter of the machine's memory, between 8 and 64 GiB: the triplet stream of BASELINE config 5 is 20 GB,
        // and re-faulting + re-locking it on every call costs three times the whole device pipeline
        static const size_t def = []() {
            const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
            size_t quarter = pages > 0 && psz > 0 ? (size_t)pages / 4 * (size_t)psz : (size_t)8 << 30;
            return std::min<size_t>(std::max<size_t>(quarter, (size_t)8 << 30), (size_t)64 << 30);
        }();
        return def;
    }
    std::mutex mu_;
    std::vector<HostBlock> free_;
    std::map<void*, HostBlock> leased_;
    size_t pooled_ = 0;
};
// never destroyed: geometry sets and sparse results may be freed while the process shuts down
static HostPool& g_host_pool = *new HostPool();
static const bool g_pinned_hooks_set = []() {
    g_pinned_hooks.alloc = [](size_t bytes) -> void* {
        try {
            return g_host_pool.lease(bytes);
        } catch (const std::bad_alloc&) {
            return nullptr;
        }
    };
    g_pinned_hooks.release = [](void* p) -> bool { return g_host_pool.unlease(p); };
    return true;
}();

// Device blocks of freed geometry sets, kept per device for the next set of about the same size: a one-shot call
// (flatten -> upload -> burn -> free) otherwise pays a cudaMalloc and a cudaFree every time, and both go through the
// driver's global lock - measured on shared boxes as stalls of tens to hundreds of milliseconds around a 3 ms call,
// and serialising the host threads of a multi-device call.  RZ_DEVICE_POOL_BYTES caps what is kept per device
// (default 8 GiB, oldest blocks go first); DevBuf::ensure and alloc_device_geoms empty the pool and retry when a
// cudaMalloc fails.
class DeviceBlockPool {
  public:
    void* get(int dev, size_t bytes, size_t* cap) {
        std::lock_guard<std::mutex> lk(mu_);
        auto& v = free_[dev];
        size_t best = v.size();
        for (size_t i = 0; i < v.size(); i++)
            if (v[i].cap >= bytes && v[i].cap <= 2 * bytes + (1u << 20) && (best == v.size() || v[i].cap < v[best].cap)) best = i;
        if (best == v.size()) return nullptr;
        void* p = v[best].p;
        *cap = v[best].cap;
        pooled_[dev] -= v[best].cap;
        v.erase(v.begin() + best);
        return p;
    }
    // the caller has made sure that no work still uses the block (cudaDeviceSynchronize, as cudaFree would)
    void put(int dev, void* p, size_t cap) {
        std::vector<Blk> evicted;
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto& v = free_[dev];
            size_t& pooled = pooled_[dev];
            if (cap <= limit()) {
                while (pooled + cap > limit() && !v.empty()) {
                    evicted.push_back(v.front());
                    pooled -= v.front().cap;
                    v.erase(v.begin());
                }
                v.push_back(Blk{p, cap});
                pooled += cap;
                p = nullptr;
            }
        }
        for (auto& e : evicted) cudaFree(e.p);
        if (p) cudaFree(p);
        (void)cudaGetLastError();
    }
    // cudaFree every pooled block of the CURRENT device `dev`; returns whether anything was freed
    bool trim(int dev) {
        std::vector<Blk> evicted;
        {
            std::lock_guard<std::mutex> lk(mu_);
            evicted.swap(free_[dev]);
            pooled_[dev] = 0;
        }
        for (auto& e : evicted) cudaFree(e.p);
        (void)cudaGetLastError();
        return !evicted.empty();
    }

  private:
    struct Blk {
        void* p;
        size_t cap;
    };
    static size_t limit() {
        static const size_t v = []() -> size_t {
            if (const char* e = std::getenv("RZ_DEVICE_POOL_BYTES")) return (size_t)std::strtoull(e, nullptr, 10);
            return (size_t)8 << 30;
        }();
        return v;
    }
    std::mutex mu_;
    std::map<int, std::vector<Blk>> free_;
    std::map<int, size_t> pooled_;
};
static DeviceBlockPool& g_device_pool = *new DeviceBlockPool();
static bool trim_device_pool_current() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return g_device_pool.trim(dev);
}

struct DeviceGeoms {
    int dev = 0;
    // every array below lives in ONE device allocation (a cudaMalloc per array made a fresh geometry set pay ~20
    // driver calls, which serialise across the host threads of a multi-device call)
    void* block = nullptr;
    double* x[3] = {nullptr, nullptr, nullptr};
    double* y[3] = {nullptr, nullptr, nullptr};
    uint32_t* tag[3] = {nullptr, nullptr, nullptr};
    uint32_t* seq_end[3] = {nullptr, nullptr, nullptr};
    uint8_t* seq_closed[3] = {nullptr, nullptr, nullptr};
    uint8_t* part_kind = nullptr;
    uint32_t* part_geom = nullptr;
    double* part_xlo = nullptr;
    double* part_xhi = nullptr;
    double* part_ylo = nullptr;
    double* part_yhi = nullptr;
    uint32_t* part_vbeg = nullptr;
    uint32_t* part_vend = nullptr;
    size_t bytes = 0;
    size_t block_cap = 0;
    ~DeviceGeoms() {
        if (!block) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(dev);
        cudaDeviceSynchronize();  // (what cudaFree implies: no kernel of any stream still reads the block)
        g_device_pool.put(dev, block, block_cap);
        if (prev >= 0) cudaSetDevice(prev);
        (void)cudaGetLastError();
    }
};

struct DeviceCtx {
    int dev = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int host_ptr_ok = 0;  // kernels may dereference cudaHostRegister'ed host pointers
    DevBuf keys_a, keys_b, hist, digit_total, task_start, part_info, field, valid, band, last_kept, counters, win_out,
        block_total, vs_keys, vs_first, sp_raw, tile_cnt, tile_cnt2, tile_off, tile_off2, tile_ctr,
        tile_pt, tile_units, tile_masks, tile_desc, tile_desc2, lb_status, sp_a, sp_b, sp_c, sp_d, sp_e, sp_f, sp_g, sp_rows, sp_cols, sp_data, sp_partial, sp_w, sp_wraw, sp_ws, cache_acc, cache_box;
    Counters* h_counters = nullptr;  // pinned + mapped: written by readback_kernel
    void* h_tile_ctr = nullptr;      // same, for TileCounters
    cudaEvent_t ev[16];
    cudaStream_t copy_stream = nullptr;  // device->host copies of finished row windows overlap the next window
    cudaStream_t copy_stream2 = nullptr; // second half of every window copy (two copy engines in flight)
    cudaStream_t upload_stream = nullptr; // geometry pools copied while they are flattened; independent of the calls
                                          // in flight on this device (one-shot calls flatten the next rows meanwhile)
    cudaEvent_t ev_half = nullptr, ev_first_fill = nullptr;
    cudaEvent_t ev_last = nullptr;  // end of the previous call on this device (scratch buffers are shared by all streams)
    bool ev_last_valid = false;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev_filled[2], ev_copied[2], ev_d2h[2], ev_bounce[2];
    DevBuf win_out2, win_t;
};

static std::mutex g_ctx_mu;
static std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;

static DeviceCtx& device_ctx(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(dev);
    if (it != g_ctx.end()) return *it->second;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        throw Error{RZ_RUNTIME_ERROR, "No CUDA device available: librz_b200 has no CPU fallback."};
    }
    if (dev < 0 || dev >= n) throw Error{RZ_RUNTIME_ERROR, "Invalid CUDA device ordinal."};
    CUDA_TRY(cudaSetDevice(dev));
    std::unique_ptr<DeviceCtx> c(new DeviceCtx());
    c->dev = dev;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cudaDeviceGetAttribute(&c->host_ptr_ok, cudaDevAttrCanUseHostPointerForRegisteredMem, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        c->host_ptr_ok = 0;
    }
    CUDA_TRY(cudaHostAlloc((void**)&c->h_counters, sizeof(Counters), cudaHostAllocMapped));
    CUDA_TRY(cudaHostAlloc(&c->h_tile_ctr, 256, cudaHostAllocMapped));
    for (auto& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_half, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&c->ev_first_fill));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_filled[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&c->ev_d2h[k]));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_bounce[k], cudaEventDisableTiming));
    }
    DeviceCtx& ref = *c;
    g_ctx[dev] = std::move(c);
    return ref;
}

// Host -> device copy of [src, src+bytes) that never spans two host registrations (CUDA rejects such a copy):
// the range is cut at the borders of the page-locked ranges; cuts outside any of them go as pageable copies.
static void copy_h2d_split(void* dst, const void* src, size_t bytes, const std::vector<std::pair<void*, size_t>>& pinned,
                           cudaStream_t s) {
    uintptr_t a = (uintptr_t)src;
    const uintptr_t end = a + bytes;
    while (a < end) {
        uintptr_t b = end;
        for (const auto& r : pinned) {
            const uintptr_t r0 = (uintptr_t)r.first, r1 = r0 + r.second;
            if (a >= r0 && a < r1) { b = std::min(end, r1); break; }  // inside a registration: up to its end
            if (r0 > a && r0 < b) b = r0;                             // outside: up to the next one
        }
        CUDA_TRY(cudaMemcpyAsync((char*)dst + (a - (uintptr_t)src), (const void*)a, b - a, cudaMemcpyHostToDevice, s));
        a = b;
    }
}

// (re)upload a host vector into its place inside the geometry set's device block (a geometry set is immutable, so a
// forced re-upload only pays the copy)
template <typename T, typename A>
static void upload_vec(T* dst, const std::vector<T, A>& v, cudaStream_t s, size_t& bytes,
                       const std::vector<std::pair<void*, size_t>>& pinned) {
    if (v.empty()) return;
    copy_h2d_split(dst, v.data(), v.size() * sizeof(T), pinned, s);
    bytes += v.size() * sizeof(T);
}

// Page-lock the vertex pools where they lie so uploads run at PCIe speed.  cudaHostRegister was seen to
// refuse one 1.6 GB vector out of three ("OS call failed"), which silently left that pool pageable (2.5x
// slower upload); a refused vector is therefore moved to freshly allocated storage and tried again, and as a
// last resort registered in page-aligned 256 MiB pieces so that only the refused pieces stay pageable.
// RZ_VERBOSE=1 reports what was refused.
template <typename T, typename A>
static bool pin_vec(rz_geoms* g, std::vector<T, A>& v, bool verbose) {  // true: fully page-locked
    const size_t bytes = v.size() * sizeof(T);
    // small arrays go as pageable copies: a cudaHostRegister call costs more than staging a few megabytes, and the
    // calls of the host threads of a multi-device job serialise inside the driver
    if (bytes < ((size_t)4 << 20)) return false;
    HostBlock blk;
    if (g_host_pool.leased(v.data(), &blk) && blk.pinned) {  // built into a recycled page-locked block
        g->pinned_ranges.emplace_back(blk.p, blk.cap);
        return true;
    }
    auto try_reg = [&](void* p, size_t n) {
        const cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterMapped | cudaHostRegisterPortable);
        if (e == cudaSuccess) {
            g->pinned_ranges.emplace_back(p, n);
            return true;
        }
        (void)cudaGetLastError();
        if (verbose) std::fprintf(stderr, "librz_b200: cudaHostRegister(%zu bytes) refused: %s\n", n, cudaGetErrorString(e));
        return false;
    };
    if (try_reg(v.data(), bytes)) return true;
    {
        std::vector<T, A> fresh(v);
        v.swap(fresh);
    }
    if (try_reg(v.data(), bytes)) return true;
    // pieces that share no page: the first starts at the vector's first byte, the last ends at its last one
    const uintptr_t piece = (uintptr_t)256 << 20, page = 4096;
    uintptr_t a = (uintptr_t)v.data();
    const uintptr_t end = a + bytes;
    bool all = true;
    while (a < end) {
        uintptr_t b = (a + piece) & ~(page - 1);
        if (b > end || b <= a) b = end;
        all = try_reg((void*)a, b - a) && all;
        a = b;
    }
    return all;
}

static void pin_host(rz_geoms* g) {
    if (g->pinned) return;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return;
    }
    const bool verbose = std::getenv("RZ_VERBOSE") != nullptr;
    for (int k = 0; k < 3; k++) {
        const bool a = pin_vec(g, g->pool[k].x, verbose);
        const bool b = pin_vec(g, g->pool[k].y, verbose);
        const bool t = pin_vec(g, g->pool[k].tag, verbose);
        if (k == 0) g->pool0_mapped = a && b && t;
    }
    pin_vec(g, g->part_xlo, verbose);  // the parts table (45 B/part) is uploaded with the pools
    pin_vec(g, g->part_xhi, verbose);
    pin_vec(g, g->part_ylo, verbose);
    pin_vec(g, g->part_yhi, verbose);
    pin_vec(g, g->part_vbeg, verbose);
    pin_vec(g, g->part_vend, verbose);
    pin_vec(g, g->part_kind, verbose);
    g->pinned = true;
}

static void unpin_host(rz_geoms* g) {
    if (!g->pinned) return;
    for (auto& r : g->pinned_ranges)  // (pool blocks stay page-locked: they are recycled)
        if (!g_host_pool.leased(r.first, nullptr) && cudaHostUnregister(r.first) != cudaSuccess) (void)cudaGetLastError();
    g->pinned_ranges.clear();
    g->pinned = false;
    g->pool0_mapped = false;
}

// carve every device array of a geometry set out of one allocation (256-byte aligned; one spare element per array
// so that kernels may read index i+1 of the last vertex unconditionally)
static void alloc_device_geoms(const rz_geoms* g, DeviceGeoms* d) {
    const uint32_t n_parts = (uint32_t)g->part_kind.size();
    if (!d->block) {
        size_t total = 0;
        auto reserve = [&](size_t count, size_t elem) {
            const size_t at = total;
            total += ((count + 1) * elem + 255) & ~(size_t)255;
            return at;
        };
        size_t o_x[3], o_y[3], o_tag[3], o_se[3], o_sc[3];
        for (int k = 0; k < 3; k++) {
            o_x[k] = reserve(g->pool[k].size(), 8);
            o_y[k] = reserve(g->pool[k].size(), 8);
            o_tag[k] = reserve(g->pool[k].size(), 4);
            o_se[k] = reserve(g->pool[k].seq_end.size(), 4);
            o_sc[k] = reserve(g->pool[k].seq_closed.size(), 1);
        }
        const size_t o_kind = reserve(n_parts, 1), o_geom = reserve(n_parts, 4), o_xlo = reserve(n_parts, 8),
                     o_xhi = reserve(n_parts, 8), o_ylo = reserve(n_parts, 8), o_yhi = reserve(n_parts, 8),
                     o_vb = reserve(n_parts, 4), o_ve = reserve(n_parts, 4);
        d->block = g_device_pool.get(d->dev, total, &d->block_cap);
        if (!d->block) {
            if (cudaMalloc(&d->block, total) != cudaSuccess) {
                (void)cudaGetLastError();
                d->block = nullptr;
                g_device_pool.trim(d->dev);
                CUDA_TRY(cudaMalloc(&d->block, total));
            }
            d->block_cap = total;
        }
        char* base = (char*)d->block;
        for (int k = 0; k < 3; k++) {
            d->x[k] = (double*)(base + o_x[k]);
            d->y[k] = (double*)(base + o_y[k]);
            d->tag[k] = (uint32_t*)(base + o_tag[k]);
            d->seq_end[k] = (uint32_t*)(base + o_se[k]);
            d->seq_closed[k] = (uint8_t*)(base + o_sc[k]);
        }
        d->part_kind = (uint8_t*)(base + o_kind);
        d->part_geom = (uint32_t*)(base + o_geom);
        d->part_xlo = (double*)(base + o_xlo);
        d->part_xhi = (double*)(base + o_xhi);
        d->part_ylo = (double*)(base + o_ylo);
        d->part_yhi = (double*)(base + o_yhi);
        d->part_vbeg = (uint32_t*)(base + o_vb);
        d->part_vend = (uint32_t*)(base + o_ve);
    }
}

// pools_done: the vertex pools were already copied while the set was being flattened (rz_geoms_from_soa_to)
static DeviceGeoms* geoms_on_device(rz_geoms* g, DeviceCtx& c, cudaStream_t s, bool force, size_t* h2d_bytes,
                                    DeviceGeoms* prefilled = nullptr) {
    std::lock_guard<std::mutex> lk(g->mu);
    auto it = g->dev.find(c.dev);
    if (it != g->dev.end() && !force && !prefilled) return it->second;
    for (int k = 0; k < 3; k++)
        if (g->pool[k].size() >= 0xfffffff0ull) throw Error{RZ_RUNTIME_ERROR, "Too many vertices (limit 2^32 per pool)."};
    pin_host(g);
    std::unique_ptr<DeviceGeoms> fresh;
    DeviceGeoms* d = prefilled ? prefilled : (it != g->dev.end() ? it->second : nullptr);
    if (prefilled) fresh.reset(prefilled);
    if (!d) {
        fresh.reset(new DeviceGeoms());
        d = fresh.get();
    }
    d->dev = c.dev;
    const uint32_t n_parts = (uint32_t)g->part_kind.size();
    alloc_device_geoms(g, d);
    size_t bytes = 0;
    for (int k = 0; k < 3; k++) {
        if (prefilled) {
            bytes += g->pool[k].size() * 16;
            continue;
        }
        upload_vec(d->x[k], g->pool[k].x, s, bytes, g->pinned_ranges);
        upload_vec(d->y[k], g->pool[k].y, s, bytes, g->pinned_ranges);
    }
    upload_vec(d->part_kind, g->part_kind, s, bytes, g->pinned_ranges);
    std::vector<uint32_t> pg(g->part_geom.begin(), g->part_geom.end());
    upload_vec(d->part_geom, pg, s, bytes, g->pinned_ranges);
    upload_vec(d->part_xlo, g->part_xlo, s, bytes, g->pinned_ranges);
    upload_vec(d->part_xhi, g->part_xhi, s, bytes, g->pinned_ranges);
    upload_vec(d->part_ylo, g->part_ylo, s, bytes, g->pinned_ranges);
    upload_vec(d->part_yhi, g->part_yhi, s, bytes, g->pinned_ranges);
    upload_vec(d->part_vbeg, g->part_vbeg, s, bytes, g->pinned_ranges);
    upload_vec(d->part_vend, g->part_vend, s, bytes, g->pinned_ranges);
    // tag[] is not sent (4 bytes per vertex, a fifth of the upload): it is rebuilt from the parts table and the
    // sequence lists - tag = part id | TAG_SEQ_END on a sequence's last vertex | TAG_CLOSED on closed line strings
    if (n_parts) {
        for (int k = 0; k < 3; k++) {
            if (!g->pool[k].size()) continue;
            tag_parts_kernel<<<(n_parts + 7) / 8, 256, 0, s>>>(n_parts, (uint8_t)k, d->part_kind, d->part_vbeg, d->part_vend,
                                                             d->tag[k]);
            const uint32_t n_seq = (uint32_t)g->pool[k].seq_end.size();
            if (!n_seq) continue;
            upload_vec(d->seq_end[k], g->pool[k].seq_end, s, bytes, g->pinned_ranges);
            upload_vec(d->seq_closed[k], g->pool[k].seq_closed, s, bytes, g->pinned_ranges);
            tag_seqs_kernel<<<(n_seq + 255) / 256, 256, 0, s>>>(n_seq, d->seq_end[k], d->seq_closed[k], d->tag[k]);
        }
    }
    CUDA_TRY(cudaStreamSynchronize(s));  // `pg` is a temporary
    d->bytes = bytes;
    if (h2d_bytes) *h2d_bytes += bytes;
    if (fresh) g->dev[c.dev] = fresh.release();
    return d;
}

// ------------------------------------------------------------------------------------------------
// device-wide scan helper (kernels in rz_sparse.cuh)
// ------------------------------------------------------------------------------------------------
template <typename Op, typename In, typename Out>
static void device_scan(In in, uint32_t n, Out out, DevBuf& partial, cudaStream_t s, uint32_t& launches) {
    const uint32_t nb = (n + SC_TILE - 1) / SC_TILE;
    partial.ensure(((size_t)nb + 2) * 8);
    unsigned long long* p = partial.as<unsigned long long>();
    if (nb == 0) {
        CUDA_TRY(cudaMemsetAsync(p, 0, 8, s));
        return;
    }
    scan_reduce_kernel<Op, In><<<nb, SC_THREADS, 0, s>>>(in, n, p);
    scan_partials_kernel<Op><<<1, 1024, 0, s>>>(p, nb);
    scan_apply_kernel<Op, In, Out><<<nb, SC_THREADS, 0, s>>>(in, n, p, out);
    launches += 3;
}

static unsigned long long scan_total(DevBuf& partial, uint32_t n, cudaStream_t s) {
    const uint32_t nb = (n + SC_TILE - 1) / SC_TILE;
    unsigned long long t = 0;
    CUDA_TRY(cudaMemcpyAsync(&t, partial.as<unsigned long long>() + nb, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return t;
}

// ------------------------------------------------------------------------------------------------
// tile-binned engine dispatch
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// the dense pipeline
// ------------------------------------------------------------------------------------------------
static uint32_t bits_for(uint64_t n_values) {  // bits needed to represent values 0 .. n_values-1
    uint32_t b = 0;
    while (b < 64 && (1ull << b) < n_values) b++;
    return b;
}

struct Window {
    uint32_t r0, r1;
};

// 1/res if res is a positive, normal power of two whose reciprocal is normal too, else 0 (see px_x)
static double pow2_reciprocal(double res) {
    int e = 0;
    if (!(res > 0.0) || std::frexp(res, &e) != 0.5 || e < -1000 || e > 1000) return 0.0;
    return 1.0 / res;
}

struct Timer {
    DeviceCtx& c;
    cudaStream_t s;
    int next = 0;
    Timer(DeviceCtx& c_, cudaStream_t s_) : c(c_), s(s_) {}
};

static const uint64_t MAX_WINDOW_RECORDS = 1ull << 31;      // 32 GiB of ping-pong key buffers
// one of the two staging buffers when `out` is host memory (RZ_WINDOW_BYTES overrides it: tests use small windows)
static uint64_t max_window_out_bytes() {
    if (const char* e = std::getenv("RZ_WINDOW_BYTES")) {
        const unsigned long long v = std::strtoull(e, nullptr, 10);
        if (v) return v;
    }
    return 2ull << 30;
}

// PixelCache boxes of every polygon part (all_touched with sum / count): c.cache_box / c.cache_acc
static void build_cache_boxes(DeviceCtx& c, cudaStream_t s, const KParams& P, DeviceGeoms* dg, uint32_t nv_poly,
                              uint32_t n_parts, uint32_t& launches) {
    c.cache_acc.ensure(std::max<size_t>((size_t)n_parts * sizeof(CacheAcc), 64));
    c.cache_box.ensure(std::max<size_t>((size_t)n_parts * sizeof(CacheBox), 64));
    cache_box_init_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(n_parts, c.cache_acc.as<CacheAcc>());
    if (nv_poly)
        cache_box_kernel<<<(nv_poly + 255) / 256, 256, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly,
                                                              c.cache_acc.as<CacheAcc>());
    cache_box_finish_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(n_parts, c.cache_acc.as<CacheAcc>(),
                                                                 c.cache_box.as<CacheBox>());
    launches += 3;
}

// restores the caller's current device when an entry point returns (multi-GPU host processes)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            (void)cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0 && cudaSetDevice(prev) != cudaSuccess) (void)cudaGetLastError();
    }
};

// The per-device scratch (DeviceCtx) is shared by all calls on a device, whatever stream they run on: a call
// first waits for the previous call's last kernel (an event), so that two streams never use the scratch at once.
static void order_after_previous_call(DeviceCtx& c, cudaStream_t s) {
    if (c.ev_last_valid && c.last_stream != s) CUDA_TRY(cudaStreamWaitEvent(s, c.ev_last, 0));
}
static void mark_call_end(DeviceCtx& c, cudaStream_t s) {
    CUDA_TRY(cudaEventRecord(c.ev_last, s));
    c.ev_last_valid = true;
    c.last_stream = s;
}

// rust/src/rasterize.rs:208-229 + the band ids of rz_group_keys
static void validate_lengths(const rz_geoms* g, const rz_context* ctx) {
    if (!ctx->field_is_scalar && ctx->field_len != g->n_geoms)
        throw Error{RZ_VALUE_ERROR, "Geometry and field lengths must match"};
    if (ctx->band_of_geom && ctx->by_len != g->n_geoms)
        throw Error{RZ_VALUE_ERROR, "Geometry and by lengths must match"};
    if (ctx->band_of_geom && !(ctx->flags & RZ_FLAG_INPUTS_ON_DEVICE)) {  // negative = skipped; anything else must name a band
        const int32_t nb = std::max(ctx->n_bands, 0);
        bool bad = false;
        for (uint64_t i = 0; i < g->n_geoms; i++) bad |= ctx->band_of_geom[i] >= nb;
        if (bad) throw Error{RZ_VALUE_ERROR, "band_of_geom holds a band index >= n_bands"};
    }
}

// field / field_valid / band_of_geom of a call on the device: copied from the caller's host arrays, or used where they
// lie with RZ_FLAG_INPUTS_ON_DEVICE (a steady-state call then moves nothing over PCIe: at 8 GPUs the 4 MB field array
// of a 1M-geometry job, staged through pageable memory on every call, cost a quarter of the whole step)
struct CallInputs {
    const uint8_t* field = nullptr;
    const uint8_t* valid = nullptr;
    const int32_t* band = nullptr;
    size_t h2d = 0;
};
static CallInputs inputs_on_device(DeviceCtx& c, cudaStream_t s, const rz_geoms* g, const rz_context* ctx, size_t isz) {
    CallInputs in;
    const size_t n_field = ctx->field_is_scalar ? 1 : (size_t)g->n_geoms;
    if (ctx->flags & RZ_FLAG_INPUTS_ON_DEVICE) {
        in.field = (const uint8_t*)ctx->field;
        in.valid = g->n_geoms ? ctx->field_valid : nullptr;
        in.band = g->n_geoms ? ctx->band_of_geom : nullptr;
        if (!in.field) throw Error{RZ_VALUE_ERROR, "RZ_FLAG_INPUTS_ON_DEVICE needs a device `field` pointer"};
        return in;
    }
    c.field.ensure(std::max<size_t>(n_field * isz, 8));
    if (n_field) CUDA_TRY(cudaMemcpyAsync(c.field.p, ctx->field, n_field * isz, cudaMemcpyHostToDevice, s));
    in.field = c.field.as<uint8_t>();
    in.h2d += n_field * isz;
    if (ctx->field_valid && g->n_geoms) {
        c.valid.ensure(g->n_geoms);
        CUDA_TRY(cudaMemcpyAsync(c.valid.p, ctx->field_valid, g->n_geoms, cudaMemcpyHostToDevice, s));
        in.valid = c.valid.as<uint8_t>();
        in.h2d += g->n_geoms;
    }
    if (ctx->band_of_geom && g->n_geoms) {
        c.band.ensure(g->n_geoms * 4);
        CUDA_TRY(cudaMemcpyAsync(c.band.p, ctx->band_of_geom, g->n_geoms * 4, cudaMemcpyHostToDevice, s));
        in.band = c.band.as<int32_t>();
        in.h2d += g->n_geoms * 4;
    }
    return in;
}

// Where a row shard lands inside a larger host array (multi-device calls): band b of the shard starts at row
// `row_off` of band b of an array holding `band_rows` rows per band.
struct DenseExtra {
    uint64_t band_rows, row_off;
};

// R's array layout (R/rusterize/src/rust/src/encoding/rarrays.rs:9-17: `permuted_axes([0, 2, 1])` made contiguous,
// i.e. (row, col, band) column-major): element (band, r, col) of a rendered window goes to
// dst[(band * ncols + col) * band_rows + row_off + r].  32 x 32 tiles through shared memory, both sides coalesced.
template <typename T>
static __global__ void __launch_bounds__(256)
window_to_rcb_kernel(const T* __restrict__ src, T* __restrict__ dst, uint32_t rows, uint32_t ncols, uint64_t band_rows,
                     uint64_t row_off, uint64_t dst_band_cols /* columns per band in dst */) {
    __shared__ T tile[32][33];
    const uint32_t b = blockIdx.z, c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;  // 32 x 8
    const T* in = src + (size_t)b * rows * ncols;
    for (uint32_t j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < ncols) tile[j][tx] = in[(size_t)(r0 + j) * ncols + c0 + tx];
    __syncthreads();
    for (uint32_t j = ty; j < 32; j += 8)
        if (c0 + j < ncols && r0 + tx < rows)
            dst[((size_t)b * dst_band_cols + c0 + j) * band_rows + row_off + r0 + tx] = tile[tx][j];
}
static void launch_window_to_rcb(cudaStream_t s, size_t isz, const void* src, void* dst, uint32_t n_bands, uint32_t rows,
                                 uint32_t ncols, uint64_t band_rows, uint64_t row_off) {
    const dim3 grid((ncols + 31) / 32, (rows + 31) / 32, n_bands);
    switch (isz) {
        case 1: window_to_rcb_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)src, (uint8_t*)dst, rows, ncols, band_rows, row_off, ncols); break;
        case 2: window_to_rcb_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)src, (uint16_t*)dst, rows, ncols, band_rows, row_off, ncols); break;
        case 4: window_to_rcb_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)src, (uint32_t*)dst, rows, ncols, band_rows, row_off, ncols); break;
        default: window_to_rcb_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t*)src, (uint64_t*)dst, rows, ncols, band_rows, row_off, ncols); break;
    }
}

struct WallClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// Copy out of a page-locked bounce block into the caller's pageable array with several threads: the destination's
// pages are usually untouched (a fresh Vec / numpy array), so the copy is bound by page faults, which scale with
// threads; one thread reaches 3-5 GB/s, the link delivers 55.
static void parallel_copy(char* dst, const char* src, size_t bytes, unsigned threads) {
    const size_t grain = (size_t)4 << 20;
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, (bytes + grain - 1) / grain));
    if (nt == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) {
        const size_t lo = bytes / nt * t & ~(size_t)4095, hi = t + 1 == nt ? bytes : bytes / nt * (t + 1) & ~(size_t)4095;
        th.emplace_back([=]() { std::memcpy(dst + lo, src + lo, hi - lo); });
    }
    std::memcpy(dst, src, nt > 1 ? (bytes / nt & ~(size_t)4095) : bytes);
    for (auto& x : th) x.join();
}

// is `p` ordinary (pageable) host memory, i.e. neither page-locked / registered nor managed?
static bool is_pageable_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static void rasterize_dense(rz_geoms* g, const rz_context* ctx, void* out, rz_stats* st, const DenseExtra* ex = nullptr) {
    const WallClock wall;
    const rz_raster_info& ri = ctx->raster_info;
    validate_lengths(g, ctx);
    const size_t isz = dtype_size(ctx->dtype);
    if (!isz) throw Error{RZ_VALUE_ERROR, "Unsupported dtype"};
    if (ctx->pixel_fn < 0 || ctx->pixel_fn > RZ_ANY) throw Error{RZ_VALUE_ERROR, "Unknown pixel function"};
    if (ri.nrows == 0 || ri.ncols == 0) return;
    if (ri.nrows >= (1ull << 31) || ri.ncols >= (1ull << 31))
        throw Error{RZ_RUNTIME_ERROR, "Raster dimensions above 2^31 are not supported."};
    const uint32_t n_bands = ctx->band_of_geom ? (uint32_t)std::max(ctx->n_bands, 0) : 1u;
    if (n_bands == 0) return;
    uint32_t shard_r0 = (uint32_t)ctx->row_begin, shard_r1 = (uint32_t)ctx->row_end;
    if (ctx->row_begin == 0 && ctx->row_end == 0) shard_r1 = (uint32_t)ri.nrows;
    if (shard_r1 > ri.nrows || shard_r0 >= shard_r1) throw Error{RZ_VALUE_ERROR, "Invalid row shard"};
    const uint32_t shard_rows = shard_r1 - shard_r0;
    const bool out_dev = (ctx->flags & RZ_FLAG_OUT_ON_DEVICE) != 0;
    // R's (row, col, band) layout: windows are always rendered into the staging buffers and transposed out of them
    const bool rcb = (ctx->flags & RZ_FLAG_OUT_ROW_COL_BAND) != 0;
    const bool direct = out_dev && !rcb;  // the kernels write the caller's device array themselves

    DeviceGuard guard;
    DeviceCtx& c = device_ctx(ctx->device);
    std::lock_guard<std::mutex> lk(c.mu);
    CUDA_TRY(cudaSetDevice(c.dev));
    cudaStream_t s = ctx->stream ? (cudaStream_t)ctx->stream : c.stream;
    order_after_previous_call(c, s);
    rz_stats S;
    std::memset(&S, 0, sizeof S);
    enum { EV_START, EV_H2D, EV_END, EV_A0 };  // EV_A0..: one event pair per stage of a window (6 pairs)
    CUDA_TRY(cudaEventRecord(c.ev[EV_START], s));

    // ---- inputs to the device ------------------------------------------------------------------
    size_t h2d = 0;
    const uint64_t MAX_WINDOW_OUT_BYTES = max_window_out_bytes();
    DeviceGeoms* dg = geoms_on_device(g, c, s, (ctx->flags & RZ_FLAG_FORCE_H2D) != 0, &h2d);
    const uint32_t n_parts = (uint32_t)g->part_kind.size();
    const CallInputs in = inputs_on_device(c, s, g, ctx, isz);
    h2d += in.h2d;
    const uint8_t* d_valid = in.valid;
    const int32_t* d_band = in.band;
    CUDA_TRY(cudaEventRecord(c.ev[EV_H2D], s));

    // ---- tiling and key layout -----------------------------------------------------------------
    // One warp owns a row tile of tile_w pixels; the fill kernel keeps one toggle bit per pixel in a
    // 32-lane x 32-bit mask, hence tile_w <= 1024.
    uint32_t tile_w = FILL_MAX_TILE_W;
    if (ctx->tile_bytes) {
        tile_w = 1;
        while ((uint64_t)tile_w * 2 * isz <= ctx->tile_bytes && tile_w * 2 <= FILL_MAX_TILE_W) tile_w *= 2;
    }
    while (tile_w / 2 >= ri.ncols && tile_w > 1) tile_w /= 2;  // no wider than the raster needs
    if ((ri.ncols + tile_w - 1) / tile_w > 65535u)
        throw Error{RZ_RUNTIME_ERROR, "Raster too wide for the chosen tile width (more than 65535 column tiles)."};
    uint32_t tile_shift = bits_for(tile_w);
    const uint32_t n_tiles = (uint32_t)((ri.ncols + tile_w - 1) / tile_w);

    KParams P;
    std::memset(&P, 0, sizeof P);
    P.xmin = ri.xmin;
    P.ymax = ri.ymax;
    P.xres = ri.xres;
    P.yres = ri.yres;
    P.inv_xres = pow2_reciprocal(ri.xres);
    P.inv_yres = pow2_reciprocal(ri.yres);
    P.nrows = (uint32_t)ri.nrows;
    P.ncols = (uint32_t)ri.ncols;
    P.nrows_f = (double)ri.nrows;
    P.ncols_f = (double)ri.ncols;
    P.tile_w = tile_w;
    P.tile_shift = tile_shift;
    P.n_tiles = n_tiles;
    P.col_bits = tile_shift + 2;  // relative columns 0 .. tile_w inclusive + the boundary-walk flag bit
    P.part_bits = std::max(1u, bits_for(std::max<uint64_t>(n_parts, 1)));
    P.part_shift = P.col_bits;
    P.task_shift = P.col_bits + P.part_bits;
    P.n_bands = n_bands;
    const bool touched = ctx->all_touched != 0;
    // all_touched: every part writes each pixel of its pixel set once (PixelCache for sum/count; for the
    // other functions repeated writes of the same value are idempotent) - prelude.rs:116-118
    P.dedup_lines = ri.xres != ri.yres || touched;
    P.n_parts = n_parts;
    S.n_parts = n_parts;
    S.n_poly_vertices = g->pool[0].size();
    S.n_line_vertices = g->pool[1].size();
    S.n_points = g->pool[2].size();
    S.tile_width = tile_w;
    S.out_bytes = (uint64_t)n_bands * shard_rows * ri.ncols * isz;

    uint64_t bg_bits = 0;
    std::memcpy(&bg_bits, ctx->background, isz);

    // ---- per-part resolution -------------------------------------------------------------------
    c.part_info.ensure(std::max<size_t>((size_t)n_parts * sizeof(PartInfo), 16));
    c.last_kept.ensure(std::max<size_t>((size_t)n_parts * 4, 16));
    c.counters.ensure(sizeof(Counters));
    if (n_parts)
        part_prepare_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(
            P, dg->part_kind, dg->part_geom, dg->part_xlo, dg->part_xhi, in.field, (uint32_t)isz,
            ctx->field_is_scalar, d_valid, d_band, c.part_info.as<PartInfo>());
    const PartInfo* d_info = c.part_info.as<PartInfo>();
    Counters* d_ctr = c.counters.as<Counters>();

    // ---- windows ---------------------------------------------------------------------------------
    // Rows are processed in windows so that (a) the key fits 64 bits, (b) the record buffers and the
    // staging buffer stay bounded.  Windows that turn out too heavy are halved.
    auto window_fits = [&](uint32_t rows) {
        uint64_t tasks = (uint64_t)n_bands * rows * n_tiles;
        if (tasks >= (1ull << 31)) return false;
        if (bits_for(tasks) + P.task_shift > 64) return false;
        if (!direct && (uint64_t)n_bands * rows * ri.ncols * isz > MAX_WINDOW_OUT_BYTES && rows > 1) return false;
        return true;
    };
    uint32_t win_rows = shard_rows;
    while (!window_fits(win_rows)) {
        if (win_rows == 1) throw Error{RZ_RUNTIME_ERROR, "Problem too large for the 64-bit record key."};
        win_rows = (win_rows + 1) / 2;
    }
    std::vector<Window> todo;
    for (uint32_t r = shard_r0; r < shard_r1; r += win_rows) todo.push_back(Window{r, std::min(r + win_rows, shard_r1)});
    std::reverse(todo.begin(), todo.end());  // pop_back walks top to bottom

    uint32_t launches = n_parts ? 1u : 0u;  // part_prepare
    FillLaunch fill = fill_for(ctx->dtype, ctx->pixel_fn);
    float count_ms = 0, emit_ms = 0, sort_ms = 0, index_ms = 0, fill_ms = 0, d2h_ms = 0;
    const bool timed = (ctx->flags & RZ_FLAG_SYNC_STAGES) != 0;
    // Stage timings: every stage records its own event pair and the pairs are read once per window, after
    // the window's last kernel: no host synchronisation sits between the stages being timed.
    float* lap_acc[6];
    int n_laps = 0;
    auto lap = [&](float& acc, int, int) {
        if (timed) lap_acc[n_laps++] = &acc;
    };
    auto flush_laps = [&]() {
        if (!timed || !n_laps) return;
        CUDA_TRY(cudaEventSynchronize(c.ev[EV_A0 + 2 * n_laps - 1]));
        for (int i = 0; i < n_laps; i++) {
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, c.ev[EV_A0 + 2 * i], c.ev[EV_A0 + 2 * i + 1]));
            *lap_acc[i] += ms;
        }
        n_laps = 0;
    };
#define EV_A (EV_A0 + 2 * n_laps)
#define EV_B (EV_A0 + 2 * n_laps + 1)
    // Host output: windows are rendered into two staging buffers in turn; the copy of a finished window runs
    // on the copy stream while the next window is computed (PCIe D2H is the longest phase of an end-to-end call).
    uint32_t n_staged = 0;
    // A pageable destination (the Array3 / numpy array a binding allocated itself): cudaMemcpyAsync into it is a
    // synchronous, driver-staged copy at a fraction of the link rate.  Large rasters therefore go through two
    // page-locked bounce blocks - the DMA of piece i+1 runs while host threads copy piece i into the caller's array
    // (RZ_BOUNCE=0 turns it off, RZ_BOUNCE_MIN_BYTES moves the 64 MB threshold).
    const size_t bounce_min = std::getenv("RZ_BOUNCE_MIN_BYTES") ? (size_t)std::strtoull(std::getenv("RZ_BOUNCE_MIN_BYTES"), nullptr, 10) : ((size_t)64 << 20);
    const bool bounce = !out_dev && !rcb && !(std::getenv("RZ_BOUNCE") && std::atoi(std::getenv("RZ_BOUNCE")) == 0) &&
                        (uint64_t)n_bands * shard_rows * ri.ncols * isz >= bounce_min && is_pageable_host(out);
    const size_t BOUNCE_BYTES = std::getenv("RZ_BOUNCE_BYTES") ? std::max<size_t>(4096, (size_t)std::strtoull(std::getenv("RZ_BOUNCE_BYTES"), nullptr, 10)) : ((size_t)128 << 20);
    HostBlock bounce_blk[2];
    struct BounceGuard {
        HostBlock* b;
        ~BounceGuard() {
            for (int i = 0; i < 2; i++) g_host_pool.put(b[i]);
        }
    } bounce_guard{bounce_blk};
    unsigned bounce_threads = 8;
    if (bounce) {
        for (auto& b : bounce_blk) {
            b = g_host_pool.get(BOUNCE_BYTES);
            if (!b.pinned && cudaHostRegister(b.p, b.cap, cudaHostRegisterPortable) == cudaSuccess) b.pinned = true;
            else (void)cudaGetLastError();
        }
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        // the copy out is bound by page faults on the fresh destination, which scale with threads (config 4, 16
        // cores: 8 threads 730 ms, 16 threads 625 ms for the raster's 17.2 GB); shards of a multi-device call share
        bounce_threads = ex ? std::max(2u, std::min(16u, hw / 4)) : std::min(16u, hw);
        if (const char* e = std::getenv("RZ_BOUNCE_THREADS")) bounce_threads = (unsigned)std::max(1, std::atoi(e));
        // the destination is about to be written from end to end: ask for huge pages where it is not touched yet
        const uintptr_t huge = (uintptr_t)2 << 20;
        const size_t band_rows_all = ex ? ex->band_rows : shard_rows;
        const uintptr_t a = ((uintptr_t)out + huge - 1) & ~(huge - 1),
                        e = ((uintptr_t)out + (size_t)n_bands * band_rows_all * ri.ncols * isz) & ~(huge - 1);
        const bool want_huge = !(std::getenv("RZ_BOUNCE_HUGEPAGE") && std::atoi(std::getenv("RZ_BOUNCE_HUGEPAGE")) == 0);
        if (want_huge && e > a && (!ex || ex->row_off == 0)) (void)madvise((void*)a, e - a, MADV_HUGEPAGE);
    }
    const int copy_streams = std::getenv("RZ_COPY_STREAMS") ? std::atoi(std::getenv("RZ_COPY_STREAMS")) : 1;  // 2: ~2 % faster when it works, but the NEXT call's upload then often runs at half speed (measured)
    auto stage_begin = [&](uint32_t rows) -> void* {  // staging buffer the window's kernels may write now
        DevBuf& b = (n_staged & 1) ? c.win_out2 : c.win_out;
        if (n_staged >= 2) CUDA_TRY(cudaStreamWaitEvent(s, c.ev_copied[n_staged & 1], 0));
        else b.ensure((size_t)n_bands * std::max(rows, std::min(win_rows, shard_rows)) * ri.ncols * isz);  // largest window
        return b.p;
    };
    auto stage_copy = [&](const Window& w, void* d_out) {  // window finished on `s`: copy it back asynchronously
        const uint32_t k = n_staged & 1, rows = w.r1 - w.r0;
        if (rcb) {
            const uint64_t band_rows = ex ? ex->band_rows : shard_rows, row_off = (ex ? ex->row_off : 0) + (w.r0 - shard_r0);
            if (out_dev) {  // transposed straight into the caller's device array
                launch_window_to_rcb(s, isz, d_out, out, n_bands, rows, (uint32_t)ri.ncols, band_rows, row_off);
                CUDA_TRY(cudaEventRecord(c.ev_copied[k], s));
                n_staged++;
                return;
            }
            // host: transpose into a compact [band][col][rows] slab, then one strided copy per band
            if (n_staged) CUDA_TRY(cudaStreamWaitEvent(s, c.ev_copied[(n_staged - 1) & 1], 0));  // the slab is reused
            c.win_t.ensure((size_t)n_bands * rows * ri.ncols * isz);
            launch_window_to_rcb(s, isz, d_out, c.win_t.p, n_bands, rows, (uint32_t)ri.ncols, rows, 0);
            if (n_staged == 0) CUDA_TRY(cudaEventRecord(c.ev_first_fill, s));
            CUDA_TRY(cudaEventRecord(c.ev_filled[k], s));
            CUDA_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_filled[k], 0));
            if (n_staged == 0) CUDA_TRY(cudaEventRecord(c.ev_d2h[0], c.copy_stream));
            for (uint32_t b = 0; b < n_bands; b++)
                CUDA_TRY(cudaMemcpy2DAsync((char*)out + ((size_t)b * ri.ncols * band_rows + row_off) * isz, band_rows * isz,
                                           (const char*)c.win_t.p + (size_t)b * ri.ncols * rows * isz, (size_t)rows * isz,
                                           (size_t)rows * isz, ri.ncols, cudaMemcpyDeviceToHost, c.copy_stream));
            CUDA_TRY(cudaEventRecord(c.ev_copied[k], c.copy_stream));
            S.d2h_bytes += (size_t)rows * ri.ncols * isz * n_bands;
            n_staged++;
            return;
        }
        if (n_staged == 0) CUDA_TRY(cudaEventRecord(c.ev_first_fill, s));
        CUDA_TRY(cudaEventRecord(c.ev_filled[k], s));
        CUDA_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_filled[k], 0));
        if (n_staged == 0) CUDA_TRY(cudaEventRecord(c.ev_d2h[0], c.copy_stream));
        const size_t chunk = (size_t)rows * ri.ncols * isz;
        if (bounce) {  // drain the window now: DMA piece i+1 into one bounce block while piece i leaves the other
            struct Piece {
                char* dst;
                const char* src;
                size_t n;
            };
            std::vector<Piece> pieces;
            for (uint32_t b = 0; b < n_bands; b++) {
                char* dst = (char*)out + ((size_t)b * (ex ? ex->band_rows : shard_rows) + (ex ? ex->row_off : 0) +
                                          (w.r0 - shard_r0)) * ri.ncols * isz;
                const char* src = (const char*)d_out + (size_t)b * chunk;
                for (size_t o = 0; o < chunk; o += BOUNCE_BYTES) pieces.push_back(Piece{dst + o, src + o, std::min(BOUNCE_BYTES, chunk - o)});
            }
            auto issue = [&](size_t i) {
                CUDA_TRY(cudaMemcpyAsync(bounce_blk[i & 1].p, pieces[i].src, pieces[i].n, cudaMemcpyDeviceToHost, c.copy_stream));
                CUDA_TRY(cudaEventRecord(c.ev_bounce[i & 1], c.copy_stream));
            };
            if (!pieces.empty()) issue(0);
            for (size_t i = 0; i < pieces.size(); i++) {
                if (i + 1 < pieces.size()) issue(i + 1);  // (block (i+1)&1 was emptied by the host copy of piece i-1)
                CUDA_TRY(cudaEventSynchronize(c.ev_bounce[i & 1]));
                parallel_copy(pieces[i].dst, (const char*)bounce_blk[i & 1].p, pieces[i].n, bounce_threads);
            }
            CUDA_TRY(cudaEventRecord(c.ev_copied[k], c.copy_stream));
            S.d2h_bytes += chunk * n_bands;
            S.host_syncs += (uint32_t)pieces.size();
            n_staged++;
            return;
        }
        // large slabs go out as two halves on two streams: two copy engines in flight fill the link more evenly
        const bool split = copy_streams >= 2 && chunk >= ((size_t)64 << 20);
        const size_t half = split ? (size_t)(rows / 2) * ri.ncols * isz : chunk;
        if (split) CUDA_TRY(cudaStreamWaitEvent(c.copy_stream2, c.ev_filled[k], 0));
        for (uint32_t b = 0; b < n_bands; b++) {
            char* dst = (char*)out + ((size_t)b * (ex ? ex->band_rows : shard_rows) + (ex ? ex->row_off : 0) +
                                      (w.r0 - shard_r0)) * ri.ncols * isz;
            const char* src = (const char*)d_out + (size_t)b * chunk;
            CUDA_TRY(cudaMemcpyAsync(dst, src, half, cudaMemcpyDeviceToHost, c.copy_stream));
            if (split) CUDA_TRY(cudaMemcpyAsync(dst + half, src + half, chunk - half, cudaMemcpyDeviceToHost, c.copy_stream2));
        }
        if (split) {
            CUDA_TRY(cudaEventRecord(c.ev_half, c.copy_stream2));
            CUDA_TRY(cudaStreamWaitEvent(c.copy_stream, c.ev_half, 0));
        }
        CUDA_TRY(cudaEventRecord(c.ev_copied[k], c.copy_stream));
        S.d2h_bytes += chunk * n_bands;
        n_staged++;
    };
    const uint32_t nv_poly = (uint32_t)g->pool[0].size(), nv_line = (uint32_t)g->pool[1].size(),
                   nv_pt = (uint32_t)g->pool[2].size();
    const uint32_t per_block = SETUP_THREADS * SETUP_ITEMS;
    const bool all_poly = nv_line == 0 && nv_pt == 0 && !ctx->all_touched;

    // all_touched with sum / count (S::REQUIRES_DEDUP): the parts whose PixelCache box does not cover their fill
    AliasCtx alias;
    std::memset(&alias, 0, sizeof alias);
    if (touched && (ctx->pixel_fn == RZ_SUM || ctx->pixel_fn == RZ_COUNT) && nv_poly) {
        P.win_r0 = shard_r0;
        P.win_r1 = shard_r1;
        build_cache_boxes(c, s, P, dg, nv_poly, n_parts, launches);
        CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), s));
        touched_walk_kernel<<<(nv_poly + 255) / 256, 256, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly, d_info, d_ctr,
                                                                 nullptr, 2, c.cache_acc.as<CacheAcc>());
        launches++;
        readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_ctr, (volatile unsigned long long*)c.h_counters,
                                         (uint32_t)(sizeof(Counters) / 8));
        CUDA_TRY(cudaStreamSynchronize(s));
        S.host_syncs++;
        const unsigned long long walked = c.h_counters->cursor;
        if (walked) {  // some part has a dropped ring segment: remember its walked pixels
            unsigned long long cap = 1024;
            while (cap < 2 * walked) cap <<= 1;
            c.vs_keys.ensure(cap * 8);
            c.vs_first.ensure(cap * 8);
            CUDA_TRY(cudaMemsetAsync(c.vs_keys.p, 0xff, cap * 8, s));
            CUDA_TRY(cudaMemsetAsync(c.vs_first.p, 0xff, cap * 8, s));
            alias.vs.keys = c.vs_keys.as<unsigned long long>();
            alias.vs.first = c.vs_first.as<unsigned long long>();
            alias.vs.mask = cap - 1;
            alias.vs.col_bits = bits_for(ri.ncols + 1);
            alias.vs.row_bits = std::max(1u, bits_for(ri.nrows));
            if (alias.vs.col_bits + alias.vs.row_bits + P.part_bits > 64)
                throw Error{RZ_RUNTIME_ERROR, "Problem too large for the 64-bit pixel-cache key."};
            touched_walk_kernel<<<(nv_poly + 255) / 256, 256, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly, d_info,
                                                                     d_ctr, nullptr, 3, c.cache_acc.as<CacheAcc>(), alias.vs);
            launches++;
            alias.acc = c.cache_acc.as<CacheAcc>();
            alias.box = c.cache_box.as<CacheBox>();
        }
    }

    // The tile-binned engine takes polygon-only jobs without all_touched whose coordinates are all finite (its
    // odd-crossing rule relies on that; see tile_mask_kernel).
    // Lines and points ride along when the pixel function is order-free (rz_burn.cuh): `any`, or `count` / `sum` on
    // an integer dtype with background 0, square pixels: they are applied to the finished tiles by atomics.
    const bool int_dtype = ctx->dtype != RZ_F32 && ctx->dtype != RZ_F64;
    const bool order_free = !touched && ri.xres == ri.yres &&
                            (ctx->pixel_fn == RZ_ANY ||
                             ((ctx->pixel_fn == RZ_COUNT || ctx->pixel_fn == RZ_SUM) && int_dtype && bg_bits == 0));
    const bool tile_candidate = (order_free || (nv_line == 0 && nv_pt == 0)) && !touched && n_parts && !g->nonfinite &&
                                !(ctx->flags & RZ_FLAG_NO_TILE_ENGINE);
    bool tile_stats_pending = false;

    while (!todo.empty()) {
        Window w = todo.back();
        todo.pop_back();
        P.win_r0 = w.r0;
        P.win_r1 = w.r1;
        const uint32_t rows = w.r1 - w.r0;

        // ---- tile-binned engine ---------------------------------------------------------------------
        if (tile_candidate) {
            TileParams T;
            std::memset(&T, 0, sizeof T);
            T.tile_r = isz <= 4 ? 64u : 32u;
            T.n_tc = (uint32_t)((ri.ncols + TILE_C - 1) / TILE_C);
            T.n_tr = (rows + T.tile_r - 1) / T.tile_r;
            const uint64_t n_tiles64 = (uint64_t)n_bands * T.n_tr * T.n_tc;
            // tile_apply: up to 32 consecutive tiles per CTA amortise the pipeline start-up; small rasters take fewer so
            // that the grid still covers the machine several times over
            T.apply_tiles = APPLY_TILES;
            while (T.apply_tiles > 1 &&
                   (uint64_t)n_bands * T.n_tr * ((T.n_tc + T.apply_tiles - 1) / T.apply_tiles) < (uint64_t)c.sm_count * 12)
                T.apply_tiles /= 2;
            if (const char* e = std::getenv("RZ_APPLY_TILES")) T.apply_tiles = (uint32_t)std::max(1, std::atoi(e));  // experiments
            if (n_tiles64 < (1ull << 31)) {  // record = [tile | block], both below 2^31
                T.n_tiles = (uint32_t)n_tiles64;
                c.tile_cnt.ensure((size_t)n_parts * 12);
                c.tile_off.ensure((size_t)n_parts * 12);
                c.tile_ctr.ensure(sizeof(TileCounters));
                TileCounters* d_tc = c.tile_ctr.as<TileCounters>();
                const int float_bytes = ctx->dtype == RZ_F32 ? 4 : (ctx->dtype == RZ_F64 ? 8 : 0);
                const unsigned long long value_mask = isz >= 8 ? ~0ull : ((1ull << (8 * isz)) - 1ull);
                auto bin = [&](int mode, int ignore_band) {
                    tile_bin_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(
                        P, T, d_info, dg->part_kind, dg->part_xlo, dg->part_xhi, dg->part_ylo, dg->part_yhi, dg->part_vbeg,
                        dg->part_vend, c.tile_cnt.as<uint32_t>(), c.tile_off.as<uint32_t>(), c.tile_pt.as<PartTile>(),
                        c.tile_units.as<uint64_t>(), c.keys_a.as<uint64_t>(), c.tile_desc.as<BlockDesc>(), d_tc, mode,
                        ignore_band, float_bytes, bg_bits, value_mask);
                    launches++;
                };
                // Upper bounds of this (geometry set, grid, window): cached in the geometry handle.  The first call
                // counts on the device with every polygon part active and reads the totals back (the one host
                // synchronisation of the engine); later calls size their buffers from the cache and never wait.
                const TilePlanKey key{ri.nrows, ri.ncols, ri.xmin, ri.ymax, ri.xres, ri.yres, w.r0, w.r1, T.tile_r};
                TilePlan plan;
                bool have_plan = false;
                {
                    std::lock_guard<std::mutex> gl(g->mu);
                    auto it = g->tile_plans.find(key);
                    if (it != g->tile_plans.end()) {
                        plan = it->second;
                        have_plan = true;
                    }
                }
                if (!have_plan) {
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                    CUDA_TRY(cudaMemsetAsync(d_tc, 0, sizeof(TileCounters), s));
                    bin(0, 1);
                    static_assert(sizeof(TileCounters) % 8 == 0 && sizeof(TileCounters) <= 256, "readback layout");
                    readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_tc, (volatile unsigned long long*)c.h_tile_ctr,
                                                     (uint32_t)(sizeof(TileCounters) / 8));
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                    lap(emit_ms, EV_A, EV_B);
                    CUDA_TRY(cudaStreamSynchronize(s));
                    S.host_syncs++;
                    TileCounters h_tc;
                    std::memcpy(&h_tc, c.h_tile_ctr, sizeof h_tc);
                    plan.pairs = h_tc.pairs;
                    plan.units = h_tc.units;
                    plan.words = h_tc.words;
                    plan.edge_visits = h_tc.edge_visits;
                    plan.cross_lb = h_tc.cross_lb;
                    std::lock_guard<std::mutex> gl(g->mu);
                    if (g->tile_plans.size() >= 256) g->tile_plans.clear();
                    g->tile_plans[key] = plan;
                } else {
                    S.plan_cached = 1;
                }
                // Cost model.  tile_mask visits every ring vertex of a part once per mask unit and 512-column chunk
                // (a few instructions per visit) and then pays per crossing; the record pipeline pays ~10 passes
                // over HBM per crossing, several times more.  The tile engine therefore wins unless the vertex
                // visits dwarf the crossings: accepted when the visits are a small multiple of the vertex count
                // (small parts: config 4) or of the crossings' lower bound, two per part row (few vertices but
                // large extents: config 3 runs 4-5x faster here than through records).  A huge, vertex-rich part
                // cut into hundreds of units fails both tests and goes to the record pipeline.
                const bool wanted = (ctx->flags & RZ_FLAG_FORCE_TILE_ENGINE) ||
                                    plan.edge_visits <= 6ull * nv_poly + (1ull << 20) ||
                                    plan.edge_visits <= 6ull * plan.cross_lb;
                if (wanted && plan.pairs < (1ull << 31) && plan.units < (1ull << 31) && plan.words < (1ull << 32) - 64) {
                    T.cap_pairs = (uint32_t)std::max<uint64_t>(plan.pairs, 1);
                    T.cap_units = (uint32_t)std::max<uint64_t>(plan.units, 1);
                    T.block_bits = std::max(1u, bits_for(T.cap_pairs));
                    const uint32_t n_rec = T.cap_pairs;
                    const uint32_t n_lb = (n_parts + LB_TILE - 1) / LB_TILE;
                    c.keys_a.ensure((size_t)n_rec * 8);
                    c.keys_b.ensure((size_t)n_rec * 8);
                    c.tile_pt.ensure((size_t)n_parts * sizeof(PartTile));
                    c.tile_units.ensure((size_t)T.cap_units * 8);
                    c.tile_desc.ensure((size_t)n_rec * sizeof(BlockDesc));
                    c.tile_desc2.ensure((size_t)n_rec * sizeof(BlockDesc));
                    c.tile_masks.ensure(std::max<size_t>((size_t)plan.words * 4, 64));
                    c.lb_status.ensure((size_t)n_lb * 3 * 8);
                    uint64_t* ka = c.keys_a.as<uint64_t>();
                    uint64_t* kb = c.keys_b.as<uint64_t>();
                    // ---- binning: counts -> offsets -> units, block descriptors, tile records ---------------
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                    CUDA_TRY(cudaMemsetAsync(d_tc, 0, sizeof(TileCounters), s));
                    CUDA_TRY(cudaMemsetAsync(c.lb_status.p, 0, (size_t)n_lb * 3 * 8, s));
                    CUDA_TRY(cudaMemsetAsync(ka, 0xff, (size_t)n_rec * 8, s));  // fillers beyond the actual count sort last
                    bin(0, 0);
                    scan_lookback_kernel<3><<<n_lb, LB_THREADS, 0, s>>>(c.tile_cnt.as<uint32_t>(), c.tile_off.as<uint32_t>(),
                                                                        n_parts, c.lb_status.as<unsigned long long>(),
                                                                        &d_tc->scan_ticket);
                    launches++;
                    bin(1, 0);
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                    lap(emit_ms, EV_A, EV_B);
                    // ---- records are in part order: a stable sort on the tile bits keeps burn order ----------
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                    // Only the tile bits are sorted.  The 0xff.. fillers behind the actual records need no extra bit: their
                    // tile field is all ones, they come last in the input, and the sort is stable - they stay behind the
                    // records of the last tile (task_index compares the whole upper key, where a filler is huge).
                    // (65536 tiles per rank at 8 GPUs: 16 bits = two passes instead of three.)
                    const uint32_t tkey_bits = T.block_bits + std::max(1u, bits_for(n_tiles64));
                    if (n_rec > 1) {
                        const uint32_t n_blocks = (n_rec + RS_TILE - 1) / RS_TILE;
                        c.hist.ensure((size_t)n_blocks * RS_RADIX * 4);
                        c.digit_total.ensure(RS_RADIX * 4);
                        for (uint32_t shift = T.block_bits; shift < tkey_bits; shift += 8) {
                            radix_hist_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, n_rec, shift, n_blocks,
                                                                              c.hist.as<uint32_t>());
                            radix_scan_rows_kernel<<<RS_RADIX, 1024, 0, s>>>(c.hist.as<uint32_t>(), n_blocks,
                                                                            c.digit_total.as<uint32_t>());
                            radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, kb, n_rec, shift, n_blocks,
                                                                                 c.hist.as<uint32_t>(),
                                                                                 c.digit_total.as<uint32_t>());
                            std::swap(ka, kb);
                            launches += 3;
                            S.sort_passes++;
                        }
                    }
                    c.task_start.ensure(((size_t)T.n_tiles + 1) * 4);
                    task_index_kernel<<<(T.n_tiles + 1 + 255) / 256, 256, 0, s>>>(ka, n_rec, T.block_bits, T.n_tiles,
                                                                                c.task_start.as<uint32_t>());
                    block_pos_kernel<<<(n_rec + 255) / 256, 256, 0, s>>>(ka, d_tc, T.block_bits, c.tile_desc.as<BlockDesc>(),
                                                                        c.tile_desc2.as<BlockDesc>());
                    launches += 2;
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                    lap(sort_ms, EV_A, EV_B);
                    // ---- inside masks of every (part, tile) block ---------------------------------
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                    {
                        // persistent grid: 8 CTAs per SM (the kernel's launch bounds), fewer when there is less work
                        const uint32_t grid = std::min<uint32_t>((T.cap_units + MASK_WARPS * MASK_UNITS - 1) / (MASK_WARPS * MASK_UNITS),
                                                                 (uint32_t)c.sm_count * 8u);
                        if (T.tile_r == 64)
                            tile_mask_kernel<64><<<grid, MASK_WARPS * 32, 0, s>>>(
                                P, T, c.tile_units.as<uint64_t>(), d_tc, c.tile_pt.as<PartTile>(), dg->part_vbeg,
                                dg->part_vend, dg->x[0], dg->y[0], dg->tag[0], c.tile_desc.as<BlockDesc>(),
                                c.tile_masks.as<uint32_t>());
                        else
                            tile_mask_kernel<32><<<grid, MASK_WARPS * 32, 0, s>>>(
                                P, T, c.tile_units.as<uint64_t>(), d_tc, c.tile_pt.as<PartTile>(), dg->part_vbeg,
                                dg->part_vend, dg->x[0], dg->y[0], dg->tag[0], c.tile_desc.as<BlockDesc>(),
                                c.tile_masks.as<uint32_t>());
                        launches++;
                    }
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                    lap(count_ms, EV_A, EV_B);  // reported as the "count" stage slot: mask build
                    // ---- apply ------------------------------------------------------------------------
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                    void* d_out;
                    if (direct) {
                        d_out = out;
                        T.out_rows = shard_rows;
                        T.win_row_off = w.r0 - shard_r0;
                    } else {
                        d_out = stage_begin(rows);
                        T.out_rows = rows;
                        T.win_row_off = 0;
                    }
                    T.vec_ok = ((uintptr_t)d_out % 16 == 0) && ((ri.ncols * isz) % 16 == 0);
                    tile_for(ctx->dtype, ctx->pixel_fn)(s, P, T, c.task_start.as<uint32_t>(), c.tile_desc2.as<BlockDesc>(),
                                                        c.tile_masks.as<uint32_t>(), d_tc, bg_bits, d_out);
                    launches++;
                    CUDA_TRY(cudaGetLastError());
                    if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                    lap(fill_ms, EV_A, EV_B);
                    if (nv_line || nv_pt) {  // order-free job: line / point pixels onto the finished tiles
                        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
                        BurnTarget B;
                        B.out = d_out;
                        B.out_rows = T.out_rows;
                        B.win_row_off = T.win_row_off;
                        B.use_part_value = ctx->pixel_fn == RZ_SUM;
                        B.one = 1ull;
                        if (ctx->dtype == RZ_F32) B.one = 0x3f800000ull;
                        if (ctx->dtype == RZ_F64) B.one = 0x3ff0000000000000ull;
                        const bool add = ctx->pixel_fn != RZ_ANY;
                        CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), s));
                        auto burn = [&](auto sz, auto adding) {
                            constexpr int SZ = decltype(sz)::value;
                            constexpr bool ADD = decltype(adding)::value;
                            if (nv_line) {
                                CUDA_TRY(cudaMemsetAsync(c.last_kept.p, 0, (size_t)n_parts * 4, s));
                                line_last_kept_kernel<<<(nv_line + 255) / 256, 256, 0, s>>>(
                                    P, dg->x[1], dg->y[1], dg->tag[1], nv_line, d_info, c.last_kept.as<uint32_t>(), d_ctr);
                                line_burn_kernel<SZ, ADD><<<(nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1],
                                                                                              nv_line, d_info, d_ctr, B);
                                line_final_burn_kernel<SZ, ADD><<<(n_parts + 255) / 256, 256, 0, s>>>(
                                    P, dg->x[1], dg->y[1], dg->tag[1], dg->part_kind, d_info, c.last_kept.as<uint32_t>(), B);
                                launches += 3;
                            }
                            if (nv_pt) {
                                point_burn_kernel<SZ, ADD><<<(nv_pt + 255) / 256, 256, 0, s>>>(P, dg->x[2], dg->y[2], dg->tag[2],
                                                                                             nv_pt, d_info, B);
                                launches++;
                            }
                        };
                        auto by_size = [&](auto adding) {
                            switch (isz) {
                                case 1: burn(std::integral_constant<int, 1>{}, adding); break;
                                case 2: burn(std::integral_constant<int, 2>{}, adding); break;
                                case 4: burn(std::integral_constant<int, 4>{}, adding); break;
                                default: burn(std::integral_constant<int, 8>{}, adding); break;
                            }
                        };
                        if (add) by_size(std::true_type{});
                        else by_size(std::false_type{});
                        CUDA_TRY(cudaGetLastError());
                        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
                        lap(index_ms, EV_A, EV_B);  // reported in the "index" stage slot: line / point burn
                        if (nv_line) {  // a segment beyond the supported domain was skipped: report it like the record path
                            readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_ctr,
                                                             (volatile unsigned long long*)c.h_counters,
                                                             (uint32_t)(sizeof(Counters) / 8));
                            CUDA_TRY(cudaStreamSynchronize(s));
                            S.host_syncs++;
                            if (c.h_counters->bad_line)
                                throw Error{RZ_RUNTIME_ERROR,
                                            "A line segment extends more than 2^29 pixels from the raster origin; unsupported."};
                        }
                    }
                    S.engine = 1;
                    S.n_records += plan.pairs;
                    S.n_mask_words += plan.words;
                    S.n_tasks += T.n_tiles;
                    S.n_windows++;
                    S.key_bits = std::max(S.key_bits, tkey_bits);
                    S.tile_width = TILE_C;
                    tile_stats_pending = true;
                    if (!direct) stage_copy(w, d_out);
                    flush_laps();
                    continue;
                }
            }
        }

        // ---- count --------------------------------------------------------------------------
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
        CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), s));
        const uint32_t poly_blocks = (nv_poly + SETUP_THREADS - 1) / SETUP_THREADS;
        if (nv_poly) {
            c.block_total.ensure((size_t)poly_blocks * 4);
            poly_count_kernel<<<poly_blocks, SETUP_THREADS, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly, d_info,
                                                                   c.block_total.as<uint32_t>(), d_ctr);
            scan_u32_kernel<<<1, 1024, 0, s>>>(c.block_total.as<uint32_t>(), poly_blocks);
            launches += 2;
        }
        if (touched) {
            if (nv_poly) touched_walk_kernel<<<(nv_poly + 255) / 256, 256, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly,
                                                                                 d_info, d_ctr, nullptr, 0);
            if (nv_line) touched_walk_kernel<<<(nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1], nv_line,
                                                                                 d_info, d_ctr, nullptr, 0);
            launches += (nv_poly != 0) + (nv_line != 0);
        } else if (nv_line) {
            CUDA_TRY(cudaMemsetAsync(c.last_kept.p, 0, (size_t)n_parts * 4, s));
            line_count_kernel<<<(nv_line + per_block - 1) / per_block, SETUP_THREADS, 0, s>>>(
                P, dg->x[1], dg->y[1], dg->tag[1], nv_line, d_info, c.last_kept.as<uint32_t>(), d_ctr);
            line_final_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1], dg->part_kind,
                                                                     d_info, c.last_kept.as<uint32_t>(), d_ctr,
                                                                     nullptr, 0);
            launches += 2;
        }
        if (nv_pt) {
            point_kernel<<<(nv_pt + 255) / 256, 256, 0, s>>>(P, dg->x[2], dg->y[2], dg->tag[2], nv_pt, d_info, d_ctr,
                                                             nullptr, 0);
            launches++;
        }
        readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_ctr, (volatile unsigned long long*)c.h_counters,
                                         (uint32_t)(sizeof(Counters) / 8));
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
        CUDA_TRY(cudaStreamSynchronize(s));
        S.host_syncs++;
        lap(count_ms, EV_A, EV_B);
        if (c.h_counters->bad_line)
            throw Error{RZ_RUNTIME_ERROR,
                        "A line segment extends more than 2^29 pixels from the raster origin; unsupported."};
        const uint64_t n_rec = c.h_counters->records;
        if (n_rec > MAX_WINDOW_RECORDS && rows > 1) {  // too heavy: halve and retry
            uint32_t mid = w.r0 + rows / 2;
            todo.push_back(Window{mid, w.r1});
            todo.push_back(Window{w.r0, mid});
            flush_laps();
            continue;
        }
        if (n_rec >= (1ull << 32) - 4096) throw Error{RZ_RUNTIME_ERROR, "Too many records in a single raster row."};
        S.n_records += n_rec;
        S.n_crossings += c.h_counters->crossings;
        S.n_windows++;
        const uint32_t n = (uint32_t)n_rec;
        const uint32_t n_tasks = n_bands * rows * n_tiles;
        S.n_tasks += n_tasks;
        const uint32_t key_bits = P.task_shift + bits_for(n_tasks);
        S.key_bits = std::max(S.key_bits, key_bits);

        // ---- emit ---------------------------------------------------------------------------
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
        c.keys_a.ensure(std::max<size_t>((size_t)n * 8, 64));
        c.keys_b.ensure(std::max<size_t>((size_t)n * 8, 64));
        uint64_t* ka = c.keys_a.as<uint64_t>();
        uint64_t* kb = c.keys_b.as<uint64_t>();
        if (n) {
            if (nv_poly) {
                poly_emit_kernel<<<poly_blocks, SETUP_THREADS, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly,
                                                                      d_info, c.block_total.as<uint32_t>(), 0, ka);
                launches++;
            }
            // line / point records go behind the polygon records
            c.h_counters->cursor = c.h_counters->poly_records;
            CUDA_TRY(cudaMemcpyAsync(&d_ctr->cursor, &c.h_counters->cursor, 8, cudaMemcpyHostToDevice, s));
            if (touched) {
                if (nv_poly) touched_walk_kernel<<<(nv_poly + 255) / 256, 256, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0],
                                                                                     nv_poly, d_info, d_ctr, ka, 1);
                if (nv_line) touched_walk_kernel<<<(nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1],
                                                                                     nv_line, d_info, d_ctr, ka, 1);
                launches += (nv_poly != 0) + (nv_line != 0);
            } else if (nv_line) {
                line_emit_kernel<<<(nv_line + per_block - 1) / per_block, SETUP_THREADS, 0, s>>>(
                    P, dg->x[1], dg->y[1], dg->tag[1], nv_line, d_info, d_ctr, ka);
                line_final_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1],
                                                                         dg->part_kind, d_info,
                                                                         c.last_kept.as<uint32_t>(), d_ctr, ka, 1);
                launches += 2;
            }
            if (nv_pt) {
                point_kernel<<<(nv_pt + 255) / 256, 256, 0, s>>>(P, dg->x[2], dg->y[2], dg->tag[2], nv_pt, d_info,
                                                                 d_ctr, ka, 1);
                launches++;
            }
        }
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
        lap(emit_ms, EV_A, EV_B);

        // ---- sort ---------------------------------------------------------------------------
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
        if (n > 1) {
            const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
            c.hist.ensure((size_t)n_blocks * RS_RADIX * 4);
            c.digit_total.ensure(RS_RADIX * 4);
            // Polygon-only jobs are emitted in part order, so a stable sort on the task bits alone
            // leaves every task's records grouped by part in burn order; otherwise sort (task, part).
            // Column bits are never sorted: the fill kernel does not need column order.
            for (uint32_t shift = all_poly ? P.task_shift : P.part_shift; shift < key_bits; shift += 8) {
                radix_hist_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, n, shift, n_blocks, c.hist.as<uint32_t>());
                radix_scan_rows_kernel<<<RS_RADIX, 1024, 0, s>>>(c.hist.as<uint32_t>(), n_blocks,
                                                                c.digit_total.as<uint32_t>());
                radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, kb, n, shift, n_blocks,
                                                                     c.hist.as<uint32_t>(),
                                                                     c.digit_total.as<uint32_t>());
                std::swap(ka, kb);
                launches += 3;
                S.sort_passes++;
            }
        }
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
        lap(sort_ms, EV_A, EV_B);

        // ---- task index ---------------------------------------------------------------------
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
        c.task_start.ensure(((size_t)n_tasks + 1) * 4);
        task_index_kernel<<<(n_tasks + 1 + 255) / 256, 256, 0, s>>>(ka, n, P.task_shift, n_tasks,
                                                                    c.task_start.as<uint32_t>());
        launches++;
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
        lap(index_ms, EV_A, EV_B);

        // ---- fill ---------------------------------------------------------------------------
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_A], s));
        FillParams F;
        F.n_tasks = n_tasks;
        F.n_tiles = n_tiles;
        F.tile_w = tile_w;
        F.ncols = (uint32_t)ri.ncols;
        F.win_rows = rows;
        F.col_bits = P.col_bits;
        F.part_shift = P.part_shift;
        F.part_bits = P.part_bits;
        F.dedup_lines = P.dedup_lines;
        F.all_poly = nv_line == 0 && nv_pt == 0;
        F.all_touched = touched;
        F.nrows = (uint32_t)ri.nrows;
        F.win_r0 = w.r0;
        void* d_out;
        if (direct) {
            d_out = out;
            F.out_rows = shard_rows;
            F.win_row_off = w.r0 - shard_r0;
        } else {
            d_out = stage_begin(rows);
            F.out_rows = rows;
            F.win_row_off = 0;
        }
        F.vec_ok = ((uintptr_t)d_out % 16 == 0) && ((ri.ncols * isz) % 16 == 0) && (((size_t)tile_w * isz) % 16 == 0);
        const uint32_t grid = (n_tasks + FILL_WARPS - 1) / FILL_WARPS;
        fill(dim3(grid), (size_t)FILL_WARPS * FILL_MAX_TILE_W * isz, s, F, ka, c.task_start.as<uint32_t>(), d_info,
             dg->part_kind, bg_bits, d_out, alias);
        launches++;
        CUDA_TRY(cudaGetLastError());
        if (timed) CUDA_TRY(cudaEventRecord(c.ev[EV_B], s));
        lap(fill_ms, EV_A, EV_B);

        // ---- copy back ----------------------------------------------------------------------
        if (!direct) stage_copy(w, d_out);
        flush_laps();
    }
    if (!out_dev && n_staged) {  // the call returns when the last window has landed in host memory
        CUDA_TRY(cudaEventRecord(c.ev_d2h[1], c.copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(s, c.ev_d2h[1], 0));
    }
    const bool sync_end = !out_dev || timed;
    if (tile_stats_pending && sync_end)  // the host waits anyway: bring the last window's counters along
        readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)c.tile_ctr.p, (volatile unsigned long long*)c.h_tile_ctr,
                                         (uint32_t)(sizeof(TileCounters) / 8));
    CUDA_TRY(cudaEventRecord(c.ev[EV_END], s));
    mark_call_end(c, s);
    if (sync_end) {
        CUDA_TRY(cudaEventSynchronize(c.ev[EV_END]));
        S.host_syncs++;
        if (tile_stats_pending) {
            TileCounters h_tc;
            std::memcpy(&h_tc, c.h_tile_ctr, sizeof h_tc);
            if (h_tc.overflow) {  // cannot happen: the cached bounds count every polygon part
                std::lock_guard<std::mutex> gl(g->mu);
                g->tile_plans.clear();
                throw Error{RZ_RUNTIME_ERROR, "Internal error: tile plan overflow."};
            }
        }
        if (!out_dev && n_staged) {
            CUDA_TRY(cudaEventElapsedTime(&d2h_ms, c.ev_d2h[0], c.ev_d2h[1]));
            if (std::getenv("RZ_VERBOSE")) {
                float t0 = 0, t1 = 0, tf = 0;
                CUDA_TRY(cudaEventElapsedTime(&tf, c.ev[EV_START], c.ev_first_fill));
                std::fprintf(stderr, "librz_b200: first window rendered at %.2f ms\n", tf);
                CUDA_TRY(cudaEventElapsedTime(&t0, c.ev[EV_START], c.ev_d2h[0]));
                CUDA_TRY(cudaEventElapsedTime(&t1, c.ev_d2h[1], c.ev[EV_END]));
                std::fprintf(stderr, "librz_b200: first window copy starts at %.2f ms, copies span %.2f ms, %.2f ms after the last\n", t0, d2h_ms, t1);
            }
        }
        CUDA_TRY(cudaEventElapsedTime(&S.total_ms, c.ev[EV_START], c.ev[EV_END]));
        CUDA_TRY(cudaEventElapsedTime(&S.h2d_ms, c.ev[EV_START], c.ev[EV_H2D]));
    }
    S.h2d_bytes = h2d;
    S.count_ms = count_ms;
    S.emit_ms = emit_ms;
    S.sort_ms = sort_ms;
    S.index_ms = index_ms;
    S.fill_ms = fill_ms;
    S.d2h_ms = d2h_ms;
    S.kernel_launches = launches;
    S.wall_ms = wall.ms();
    if (st) *st = S;
#undef EV_A
#undef EV_B
}


// ------------------------------------------------------------------------------------------------
// the sparse pipeline
// ------------------------------------------------------------------------------------------------
}  // namespace rz

struct rz_sparse {
    rz::HostBlock rows, cols, data;  // u64[len], u64[len], dtype[len]
    uint64_t len = 0;
    std::vector<uint64_t> counts;
    ~rz_sparse() {
        rz::g_host_pool.put(rows);
        rz::g_host_pool.put(cols);
        rz::g_host_pool.put(data);
    }
};

namespace rz {

struct SparseJob {
    uint32_t n_rec, nv_poly, nv_line, nv_pt;
    const uint64_t* keys;
    VisitSet vs;
    bool touched, line_dedup, poly_dedup;
};

template <typename N>
static void sparse_expand(cudaStream_t s, const KParams& P, SparseLayout L, DeviceGeoms* dg, DeviceCtx& c,
                          const SparseJob& J, uint32_t& launches) {
    unsigned long long* rows = c.sp_rows.as<unsigned long long>();
    unsigned long long* cols = c.sp_cols.as<unsigned long long>();
    N* data = c.sp_data.as<N>();
    const PartInfo* info = c.part_info.as<PartInfo>();
    const unsigned long long* base = c.sp_f.as<unsigned long long>();
    const unsigned long long* start = c.sp_e.as<unsigned long long>();
    if (J.n_rec) {
        if (J.poly_dedup)
            poly_expand_dedup_kernel<N><<<(J.n_rec + 255) / 256, 256, 0, s>>>(
                P, J.keys, c.sp_a.as<uint32_t>(), c.sp_b.as<unsigned long long>(), J.n_rec, L, J.vs,
                c.cache_box.as<CacheBox>(), info, base, start, rows, cols, data);
        else
            poly_expand_kernel<N><<<(J.n_rec + 255) / 256, 256, 0, s>>>(J.keys, c.sp_a.as<uint32_t>(),
                                                                       c.sp_b.as<unsigned long long>(), J.n_rec, L, info,
                                                                       base, start, rows, cols, data);
        launches++;
    }
    if (J.touched && J.nv_poly) {  // pass 1 of every polygon part: its rings' boundary walk
        line_expand_kernel<N, true><<<(J.nv_poly + 255) / 256, 256, 0, s>>>(
            P, dg->x[0], dg->y[0], dg->tag[0], J.nv_poly, info, c.last_kept.as<uint32_t>(), c.counters.as<Counters>(),
            c.sp_w.as<unsigned long long>(), c.sp_wraw.as<unsigned long long>(), J.vs, J.poly_dedup ? 1 : 0, base,
            c.sp_ws.as<unsigned long long>(), rows, cols, data);
        launches++;
    }
    if (J.nv_line) {
        auto k = J.touched ? line_expand_kernel<N, true> : line_expand_kernel<N, false>;
        k<<<(J.nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1], J.nv_line, info,
                                                  c.last_kept.as<uint32_t>(), c.counters.as<Counters>(),
                                                  c.sp_c.as<unsigned long long>(), c.sp_raw.as<unsigned long long>(), J.vs,
                                                  J.line_dedup ? 1 : 0, base, start, rows, cols, data);
        launches++;
    }
    if (J.nv_pt) {
        point_expand_kernel<N><<<(J.nv_pt + 255) / 256, 256, 0, s>>>(P, dg->x[2], dg->y[2], dg->tag[2], J.nv_pt, info,
                                                                    c.sp_d.as<unsigned long long>(), base, start, rows,
                                                                    cols, data);
        launches++;
    }
}

// Multi-device calls: where a device's share of the triplet stream lands.  Once a device knows its per-band counts
// it asks the sink for the element offset of each of its bands inside the final host arrays (the sink waits for
// the counts of all devices, sizes the arrays once, and hands every device disjoint slices), then copies its
// bands there straight from device memory.
struct SparseSink {
    virtual ~SparseSink() {}
    // counts[b]: this device's triplets of band b -> elem_off[b]; the final arrays are then allocated
    virtual void place(const std::vector<uint64_t>& counts, std::vector<uint64_t>& elem_off, char*& rows, char*& cols,
                       char*& data) = 0;
};

static void rasterize_sparse(rz_geoms* g, const rz_context* ctx, rz_sparse* out, rz_stats* st, SparseSink* sink = nullptr) {
    const WallClock wall;
    const rz_raster_info& ri = ctx->raster_info;
    validate_lengths(g, ctx);
    const size_t isz = dtype_size(ctx->dtype);
    if (!isz) throw Error{RZ_VALUE_ERROR, "Unsupported dtype"};
    if (ctx->pixel_fn < 0 || ctx->pixel_fn > RZ_ANY) throw Error{RZ_VALUE_ERROR, "Unknown pixel function"};
    const uint32_t n_bands = ctx->band_of_geom ? (uint32_t)std::max(ctx->n_bands, 0) : 1u;
    out->counts.assign(n_bands, 0);
    // (a device of a multi-device call always reports its counts: the others wait for them)
    auto place_nothing = [&]() {
        if (!sink) return;
        std::vector<uint64_t> off;
        char *r = nullptr, *cc = nullptr, *d = nullptr;
        sink->place(out->counts, off, r, cc, d);
    };
    if (ri.nrows == 0 || ri.ncols == 0 || n_bands == 0) return place_nothing();
    if (ri.nrows >= (1ull << 31) || ri.ncols >= (1ull << 31))
        throw Error{RZ_RUNTIME_ERROR, "Raster dimensions above 2^31 are not supported."};
    const uint32_t nv_poly = (uint32_t)g->pool[0].size(), nv_line = (uint32_t)g->pool[1].size(),
                   nv_pt = (uint32_t)g->pool[2].size();
    // all_touched (burners.rs:94-247): line parts and, as pass 1, polygon rings are walked GDAL-style; with
    // sum / count (S::REQUIRES_DEDUP, prelude.rs:116-118) every part writes a pixel once (PixelCache)
    const bool touched = ctx->all_touched != 0;
    const bool fn_dedup = touched && (ctx->pixel_fn == RZ_SUM || ctx->pixel_fn == RZ_COUNT);
    const bool line_dedup = (ri.xres != ri.yres || fn_dedup) && nv_line;  // burn_geometry.rs:179, 202
    const bool poly_dedup = fn_dedup && nv_poly;                          // burn_geometry.rs:99-103

    DeviceGuard guard;
    DeviceCtx& c = device_ctx(ctx->device);
    std::lock_guard<std::mutex> lk(c.mu);
    CUDA_TRY(cudaSetDevice(c.dev));
    cudaStream_t s = ctx->stream ? (cudaStream_t)ctx->stream : c.stream;
    order_after_previous_call(c, s);
    rz_stats S;
    std::memset(&S, 0, sizeof S);
    enum { EV_START, EV_END, EV_SORTED, EV_COUNTED, EV_EXPANDED, EV_EXPAND };  // stage marks (rz_stats: emit+sort, index, fill, d2h)
    CUDA_TRY(cudaEventRecord(c.ev[EV_START], s));
    size_t h2d = 0;
    DeviceGeoms* dg = geoms_on_device(g, c, s, (ctx->flags & RZ_FLAG_FORCE_H2D) != 0, &h2d);
    const uint32_t n_parts = (uint32_t)g->part_kind.size();
    const CallInputs in = inputs_on_device(c, s, g, ctx, isz);
    const uint8_t* d_valid = in.valid;
    const int32_t* d_band = in.band;
    S.h2d_bytes = h2d + in.h2d;

    KParams P;
    std::memset(&P, 0, sizeof P);
    P.xmin = ri.xmin;
    P.ymax = ri.ymax;
    P.xres = ri.xres;
    P.yres = ri.yres;
    P.inv_xres = pow2_reciprocal(ri.xres);
    P.inv_yres = pow2_reciprocal(ri.yres);
    P.nrows = (uint32_t)ri.nrows;
    P.ncols = (uint32_t)ri.ncols;
    P.nrows_f = (double)ri.nrows;
    P.ncols_f = (double)ri.ncols;
    P.tile_w = 1u << 31;  // a single column tile: crossings keep absolute columns
    P.tile_shift = 31;
    P.n_tiles = 1;
    P.win_r0 = 0;
    P.win_r1 = P.nrows;
    P.n_bands = n_bands;
    P.n_parts = n_parts;
    SparseLayout L;
    L.col_bits = bits_for(ri.ncols + 1);
    L.row_bits = std::max(1u, bits_for(ri.nrows));
    const uint32_t part_bits = std::max(1u, bits_for(std::max<uint64_t>(n_parts, 1)));
    const uint32_t key_bits = L.col_bits + L.row_bits + part_bits;
    if (key_bits > 64) throw Error{RZ_RUNTIME_ERROR, "Problem too large for the 64-bit sparse record key."};
    S.key_bits = key_bits;
    S.n_parts = n_parts;
    S.n_poly_vertices = nv_poly;
    S.n_line_vertices = nv_line;
    S.n_points = nv_pt;
    uint32_t launches = 0;
    if (n_parts == 0) return place_nothing();

    c.part_info.ensure((size_t)n_parts * sizeof(PartInfo));
    c.last_kept.ensure((size_t)n_parts * 4);
    c.counters.ensure(sizeof(Counters));
    part_prepare_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(P, dg->part_kind, dg->part_geom, dg->part_xlo,
                                                              dg->part_xhi, in.field, (uint32_t)isz,
                                                              ctx->field_is_scalar, d_valid, d_band,
                                                              c.part_info.as<PartInfo>());
    launches++;
    const PartInfo* d_info = c.part_info.as<PartInfo>();
    Counters* d_ctr = c.counters.as<Counters>();
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(Counters), s));

    // ---- polygon crossings: count, emit in part order, sort by (part,row,col) ---------------------
    uint32_t n_rec = 0;
    uint64_t* keys = nullptr;
    if (nv_poly) {
        const uint32_t poly_blocks = (nv_poly + SETUP_THREADS - 1) / SETUP_THREADS;
        c.block_total.ensure((size_t)poly_blocks * 4);
        poly_count_kernel<<<poly_blocks, SETUP_THREADS, 0, s>>>(P, dg->x[0], dg->y[0], dg->tag[0], nv_poly, d_info,
                                                               c.block_total.as<uint32_t>(), d_ctr);
        scan_u32_kernel<<<1, 1024, 0, s>>>(c.block_total.as<uint32_t>(), poly_blocks);
        launches += 2;
        readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_ctr, (volatile unsigned long long*)c.h_counters,
                                         (uint32_t)(sizeof(Counters) / 8));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (c.h_counters->records >= (1ull << 32) - 4096)
            throw Error{RZ_RUNTIME_ERROR, "Too many polygon crossings for one sparse call (limit 2^32)."};
        n_rec = (uint32_t)c.h_counters->records;
        S.n_crossings = n_rec;
        S.n_records = n_rec;
        if (n_rec) {
            c.keys_a.ensure((size_t)n_rec * 8);
            c.keys_b.ensure((size_t)n_rec * 8);
            uint64_t* ka = c.keys_a.as<uint64_t>();
            uint64_t* kb = c.keys_b.as<uint64_t>();
            poly_emit_sparse_kernel<<<poly_blocks, SETUP_THREADS, 0, s>>>(P, L, dg->x[0], dg->y[0], dg->tag[0], nv_poly,
                                                                         d_info, c.block_total.as<uint32_t>(), ka);
            launches++;
            if (n_rec > 1) {
                const uint32_t n_blocks = (n_rec + RS_TILE - 1) / RS_TILE;
                c.hist.ensure((size_t)n_blocks * RS_RADIX * 4);
                c.digit_total.ensure(RS_RADIX * 4);
                // records are emitted in part order: a stable sort on (row, col) would interleave parts, so
                // all bits are sorted
                for (uint32_t shift = 0; shift < key_bits; shift += 8) {
                    radix_hist_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, n_rec, shift, n_blocks, c.hist.as<uint32_t>());
                    radix_scan_rows_kernel<<<RS_RADIX, 1024, 0, s>>>(c.hist.as<uint32_t>(), n_blocks,
                                                                    c.digit_total.as<uint32_t>());
                    radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, kb, n_rec, shift, n_blocks,
                                                                         c.hist.as<uint32_t>(),
                                                                         c.digit_total.as<uint32_t>());
                    std::swap(ka, kb);
                    launches += 3;
                    S.sort_passes++;
                }
            }
            keys = ka;
        }
    }
    CUDA_TRY(cudaEventRecord(c.ev[EV_SORTED], s));
    // ---- write counts of every unit -------------------------------------------------------------------
    unsigned long long poly_total = 0, line_total = 0, pt_total = 0, walk_total = 0;
    VisitSet vs;
    std::memset(&vs, 0, sizeof vs);
    if (n_rec) {  // (part,row) segments of the sorted crossings
        c.sp_a.ensure((size_t)n_rec * 4);  // seg_start
        c.sp_b.ensure((size_t)n_rec * 8);  // poly_off
        device_scan<OpMax>(InSegHead{keys, L.col_bits}, n_rec, OutSegStart{c.sp_a.as<uint32_t>()}, c.sp_partial, s,
                           launches);
    }
    // raw write counts of the line parts and of the polygon boundary walks
    const InLineLen<false> line_len{P, dg->x[1], dg->y[1], dg->tag[1], d_info, c.last_kept.as<uint32_t>(), d_ctr, nv_line};
    const InLineLen<true> line_walk{P, dg->x[1], dg->y[1], dg->tag[1], d_info, c.last_kept.as<uint32_t>(), d_ctr, nv_line};
    const InLineLen<true> poly_walk{P, dg->x[0], dg->y[0], dg->tag[0], d_info, c.last_kept.as<uint32_t>(), d_ctr, nv_poly};
    if (nv_line) {
        c.sp_c.ensure((size_t)nv_line * 8);
        if (touched) {
            device_scan<OpAdd>(line_walk, nv_line, OutPrefix64{c.sp_c.as<unsigned long long>()}, c.sp_partial, s, launches);
        } else {
            CUDA_TRY(cudaMemsetAsync(c.last_kept.p, 0, (size_t)n_parts * 4, s));
            line_last_kept_kernel<<<(nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1], nv_line,
                                                                       d_info, c.last_kept.as<uint32_t>(), d_ctr);
            launches++;
            device_scan<OpAdd>(line_len, nv_line, OutPrefix64{c.sp_c.as<unsigned long long>()}, c.sp_partial, s, launches);
        }
        line_total = scan_total(c.sp_partial, nv_line, s);
        readback_kernel<<<1, 32, 0, s>>>((const unsigned long long*)d_ctr, (volatile unsigned long long*)c.h_counters,
                                         (uint32_t)(sizeof(Counters) / 8));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (c.h_counters->bad_line)
            throw Error{RZ_RUNTIME_ERROR,
                        "A line segment extends more than 2^29 pixels from the raster origin; unsupported."};
    }
    if (touched && nv_poly) {
        c.sp_w.ensure((size_t)nv_poly * 8);   // walk_off
        c.sp_ws.ensure((size_t)n_parts * 8);  // walk_start
        device_scan<OpAdd>(poly_walk, nv_poly, OutPrefix64{c.sp_w.as<unsigned long long>()}, c.sp_partial, s, launches);
        walk_total = scan_total(c.sp_partial, nv_poly, s);
    }
    // First-visit filter (PixelCache): burn indices come from the raw prefixes; a hash set keeps the smallest
    // one per (part,row,col); kept writes are then re-scanned.
    const bool ld = line_dedup && line_total, pd = poly_dedup && walk_total;
    if (ld || pd) {
        const unsigned long long entries = (ld ? line_total : 0) + (pd ? walk_total : 0);
        unsigned long long cap = 1024;
        while (cap < 2 * entries) cap <<= 1;
        c.vs_keys.ensure(cap * 8);
        c.vs_first.ensure(cap * 8);
        CUDA_TRY(cudaMemsetAsync(c.vs_keys.p, 0xff, cap * 8, s));
        CUDA_TRY(cudaMemsetAsync(c.vs_first.p, 0xff, cap * 8, s));
        vs.keys = c.vs_keys.as<unsigned long long>();
        vs.first = c.vs_first.as<unsigned long long>();
        vs.mask = cap - 1;
        vs.col_bits = L.col_bits;
        vs.row_bits = L.row_bits;
    }
    if (ld) {
        c.sp_raw.ensure((size_t)nv_line * 8);
        CUDA_TRY(cudaMemcpyAsync(c.sp_raw.p, c.sp_c.p, (size_t)nv_line * 8, cudaMemcpyDeviceToDevice, s));
        auto ins = touched ? line_visit_insert_kernel<true> : line_visit_insert_kernel<false>;
        ins<<<(nv_line + 255) / 256, 256, 0, s>>>(P, dg->x[1], dg->y[1], dg->tag[1], nv_line, d_info,
                                                  c.last_kept.as<uint32_t>(), d_ctr, c.sp_raw.as<unsigned long long>(), vs);
        launches++;
        if (touched)
            device_scan<OpAdd>(InLineKept<true>{line_walk, c.sp_raw.as<unsigned long long>(), vs}, nv_line,
                               OutPrefix64{c.sp_c.as<unsigned long long>()}, c.sp_partial, s, launches);
        else
            device_scan<OpAdd>(InLineKept<false>{line_len, c.sp_raw.as<unsigned long long>(), vs}, nv_line,
                               OutPrefix64{c.sp_c.as<unsigned long long>()}, c.sp_partial, s, launches);
        line_total = scan_total(c.sp_partial, nv_line, s);
    }
    if (pd) {
        c.sp_wraw.ensure((size_t)nv_poly * 8);
        CUDA_TRY(cudaMemcpyAsync(c.sp_wraw.p, c.sp_w.p, (size_t)nv_poly * 8, cudaMemcpyDeviceToDevice, s));
        line_visit_insert_kernel<true><<<(nv_poly + 255) / 256, 256, 0, s>>>(
            P, dg->x[0], dg->y[0], dg->tag[0], nv_poly, d_info, c.last_kept.as<uint32_t>(), d_ctr,
            c.sp_wraw.as<unsigned long long>(), vs);
        launches++;
        device_scan<OpAdd>(InLineKept<true>{poly_walk, c.sp_wraw.as<unsigned long long>(), vs}, nv_poly,
                           OutPrefix64{c.sp_w.as<unsigned long long>()}, c.sp_partial, s, launches);
        walk_total = scan_total(c.sp_partial, nv_poly, s);
    }
    if (n_rec) {  // spans: pair the sorted crossings, prefix-sum their (kept) lengths
        const InSpanLen span_len{keys, c.sp_a.as<uint32_t>(), n_rec, L.col_bits};
        if (pd) {
            build_cache_boxes(c, s, P, dg, nv_poly, n_parts, launches);
            device_scan<OpAdd>(InSpanKept{span_len, vs, P, c.cache_box.as<CacheBox>(), L.row_bits}, n_rec,
                               OutPrefix64{c.sp_b.as<unsigned long long>()}, c.sp_partial, s, launches);
        }
        else
            device_scan<OpAdd>(span_len, n_rec, OutPrefix64{c.sp_b.as<unsigned long long>()}, c.sp_partial, s, launches);
        poly_total = scan_total(c.sp_partial, n_rec, s);
    }
    if (nv_pt) {
        c.sp_d.ensure((size_t)nv_pt * 8);
        device_scan<OpAdd>(InPointHit{P, dg->x[2], dg->y[2], dg->tag[2], d_info}, nv_pt,
                           OutPrefix64{c.sp_d.as<unsigned long long>()}, c.sp_partial, s, launches);
        pt_total = scan_total(c.sp_partial, nv_pt, s);
    }
    // ---- per-part counts, band-major bases ------------------------------------------------------
    c.task_start.ensure(((size_t)n_parts + 1) * 4);  // rec_beg
    part_rec_range_kernel<<<(n_parts + 1 + 255) / 256, 256, 0, s>>>(keys, n_rec, L.col_bits + L.row_bits, n_parts,
                                                                    c.task_start.as<uint32_t>());
    c.sp_g.ensure((size_t)n_parts * 8);  // count
    c.sp_e.ensure((size_t)n_parts * 8);  // start
    c.sp_f.ensure((size_t)n_parts * 8);  // base
    part_count_kernel<<<(n_parts + 255) / 256, 256, 0, s>>>(
        n_parts, dg->part_kind, dg->part_vbeg, dg->part_vend, c.task_start.as<uint32_t>(),
        c.sp_b.as<unsigned long long>(), n_rec, poly_total, c.sp_c.as<unsigned long long>(), nv_line, line_total,
        c.sp_d.as<unsigned long long>(), nv_pt, pt_total,
        (touched && nv_poly) ? c.sp_w.as<unsigned long long>() : nullptr, nv_poly, walk_total,
        c.sp_g.as<unsigned long long>(), c.sp_e.as<unsigned long long>(), c.sp_ws.as<unsigned long long>());
    launches += 2;
    unsigned long long total = 0;
    for (uint32_t b = 0; b < n_bands; b++) {  // bands in sorted key order, parts ascending inside a band
        device_scan<OpAdd>(InBandCount{c.sp_g.as<unsigned long long>(), d_info, (int32_t)b}, n_parts,
                           OutBandBase{c.sp_f.as<unsigned long long>(), d_info, (int32_t)b, total}, c.sp_partial, s,
                           launches);
        const unsigned long long bt = scan_total(c.sp_partial, n_parts, s);
        out->counts[b] = bt;
        total += bt;
    }
    CUDA_TRY(cudaEventRecord(c.ev[EV_COUNTED], s));
    // ---- expand ------------------------------------------------------------------------------------
    out->len = total;
    std::vector<uint64_t> sink_off;
    char *h_rows = nullptr, *h_cols = nullptr, *h_data = nullptr;
    if (sink) {
        sink->place(out->counts, sink_off, h_rows, h_cols, h_data);
    } else {
        out->rows = g_host_pool.get(total * 8);
        out->cols = g_host_pool.get(total * 8);
        out->data = g_host_pool.get(total * isz);
        h_rows = (char*)out->rows.p;
        h_cols = (char*)out->cols.p;
        h_data = (char*)out->data.p;
    }
    if (total) {
        c.sp_rows.ensure(total * 8);
        c.sp_cols.ensure(total * 8);
        c.sp_data.ensure(total * isz);
        CUDA_TRY(cudaEventRecord(c.ev[EV_EXPAND], s));  // (host blocks and device arrays exist: the kernels start now)
        SparseJob J;
        J.n_rec = n_rec;
        J.nv_poly = nv_poly;
        J.nv_line = nv_line;
        J.nv_pt = nv_pt;
        J.keys = keys;
        J.vs = vs;
        J.touched = touched;
        J.line_dedup = ld;
        J.poly_dedup = pd;
        switch (isz) {  // triplet values are moved bit-wise: dispatch on the item size only
            case 1: sparse_expand<uint8_t>(s, P, L, dg, c, J, launches); break;
            case 2: sparse_expand<uint16_t>(s, P, L, dg, c, J, launches); break;
            case 4: sparse_expand<uint32_t>(s, P, L, dg, c, J, launches); break;
            default: sparse_expand<uint64_t>(s, P, L, dg, c, J, launches); break;
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c.ev[EV_EXPANDED], s));
        if (!sink) {
            CUDA_TRY(cudaMemcpyAsync(h_rows, c.sp_rows.p, total * 8, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(h_cols, c.sp_cols.p, total * 8, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(h_data, c.sp_data.p, total * isz, cudaMemcpyDeviceToHost, s));
        } else {  // this device's stream is band-major; band b goes to its slice of band b of the final arrays
            uint64_t at = 0;
            for (uint32_t b = 0; b < n_bands; b++) {
                const uint64_t n = out->counts[b], o = sink_off[b];
                if (n) {
                    CUDA_TRY(cudaMemcpyAsync(h_rows + o * 8, (char*)c.sp_rows.p + at * 8, n * 8, cudaMemcpyDeviceToHost, s));
                    CUDA_TRY(cudaMemcpyAsync(h_cols + o * 8, (char*)c.sp_cols.p + at * 8, n * 8, cudaMemcpyDeviceToHost, s));
                    CUDA_TRY(cudaMemcpyAsync(h_data + o * isz, (char*)c.sp_data.p + at * isz, n * isz, cudaMemcpyDeviceToHost, s));
                }
                at += n;
            }
        }
        S.d2h_bytes = total * (16 + isz);
    }
    CUDA_TRY(cudaEventRecord(c.ev[EV_END], s));
    mark_call_end(c, s);
    CUDA_TRY(cudaEventSynchronize(c.ev[EV_END]));
    CUDA_TRY(cudaEventElapsedTime(&S.total_ms, c.ev[EV_START], c.ev[EV_END]));
    CUDA_TRY(cudaEventElapsedTime(&S.sort_ms, c.ev[EV_START], c.ev[EV_SORTED]));    // upload + crossings + sort
    CUDA_TRY(cudaEventElapsedTime(&S.index_ms, c.ev[EV_SORTED], c.ev[EV_COUNTED])); // unit scans, bases
    if (total) {
        CUDA_TRY(cudaEventElapsedTime(&S.fill_ms, c.ev[EV_EXPAND], c.ev[EV_EXPANDED]));  // expand kernels
        CUDA_TRY(cudaEventElapsedTime(&S.d2h_ms, c.ev[EV_EXPANDED], c.ev[EV_END]));       // pin + copy back
    }
    S.out_bytes = total * (16 + isz);
    S.kernel_launches = launches;
    S.wall_ms = wall.ms();
    if (st) *st = S;
}


// ------------------------------------------------------------------------------------------------
// sparse replay (SparseArray::build_array)
// ------------------------------------------------------------------------------------------------
static void sparse_build_array(const rz_context* ctx, uint64_t n_bands, const uint64_t* counts, const uint64_t* rows,
                               const uint64_t* cols, const void* data, void* out, rz_stats* st) {
    const rz_raster_info& ri = ctx->raster_info;
    const size_t isz = dtype_size(ctx->dtype);
    if (!isz) throw Error{RZ_VALUE_ERROR, "Unsupported dtype"};
    if (ctx->pixel_fn < 0 || ctx->pixel_fn > RZ_ANY) throw Error{RZ_VALUE_ERROR, "Unknown pixel function"};
    if (ri.nrows == 0 || ri.ncols == 0 || n_bands == 0) return;
    if (ri.nrows >= (1ull << 31) || ri.ncols >= (1ull << 31))
        throw Error{RZ_RUNTIME_ERROR, "Raster dimensions above 2^31 are not supported."};
    std::vector<unsigned long long> band_off(n_bands + 1, 0);
    for (uint64_t b = 0; b < n_bands; b++) band_off[b + 1] = band_off[b] + counts[b];
    const uint64_t n64 = band_off[n_bands];
    if (n64 >= (1ull << 32) - 4096) throw Error{RZ_RUNTIME_ERROR, "Too many triplets for one replay call (limit 2^32)."};
    const uint32_t n = (uint32_t)n64;
    const bool out_dev = (ctx->flags & RZ_FLAG_OUT_ON_DEVICE) != 0;

    DeviceGuard guard;
    DeviceCtx& c = device_ctx(ctx->device);
    std::lock_guard<std::mutex> lk(c.mu);
    CUDA_TRY(cudaSetDevice(c.dev));
    cudaStream_t s = ctx->stream ? (cudaStream_t)ctx->stream : c.stream;
    order_after_previous_call(c, s);
    uint32_t tile_w = FILL_MAX_TILE_W;
    while (tile_w / 2 >= ri.ncols && tile_w > 1) tile_w /= 2;
    const uint32_t tile_shift = bits_for(tile_w);
    const uint32_t n_tiles = (uint32_t)((ri.ncols + tile_w - 1) / tile_w);
    const uint64_t n_tasks64 = n_bands * ri.nrows * n_tiles;
    const uint32_t idx_bits = std::max(1u, bits_for(std::max<uint64_t>(n, 1)));
    if (n_tasks64 >= (1ull << 31) || bits_for(n_tasks64 + 1) + idx_bits > 64)
        throw Error{RZ_RUNTIME_ERROR, "Raster too large for a single replay call."};
    const uint32_t n_tasks = (uint32_t)n_tasks64;
    uint32_t launches = 0;
    const size_t out_bytes = (size_t)n_tasks64 / n_tiles * ri.ncols * isz;
    void* d_out = out;
    if (!out_dev) {
        c.win_out.ensure(out_bytes);
        d_out = c.win_out.p;
    }
    c.sp_rows.ensure(std::max<size_t>((size_t)n * 8, 8));
    c.sp_cols.ensure(std::max<size_t>((size_t)n * 8, 8));
    c.sp_data.ensure(std::max<size_t>((size_t)n * isz, 8));
    c.sp_g.ensure((n_bands + 1) * 8);
    c.keys_a.ensure(std::max<size_t>((size_t)n * 8, 64));
    c.keys_b.ensure(std::max<size_t>((size_t)n * 8, 64));
    uint64_t* ka = c.keys_a.as<uint64_t>();
    uint64_t* kb = c.keys_b.as<uint64_t>();
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(c.sp_rows.p, rows, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c.sp_cols.p, cols, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c.sp_data.p, data, (size_t)n * isz, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c.sp_g.p, band_off.data(), (n_bands + 1) * 8, cudaMemcpyHostToDevice, s));
        replay_emit_kernel<<<(n + 255) / 256, 256, 0, s>>>(c.sp_rows.as<unsigned long long>(),
                                                           c.sp_cols.as<unsigned long long>(), n,
                                                           c.sp_g.as<unsigned long long>(), (uint32_t)n_bands,
                                                           (uint32_t)ri.nrows, (uint32_t)ri.ncols, n_tiles, tile_shift,
                                                           idx_bits, n_tasks, ka);
        launches++;
        const uint32_t key_bits = idx_bits + bits_for((uint64_t)n_tasks + 1);
        if (n > 1) {
            const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
            c.hist.ensure((size_t)n_blocks * RS_RADIX * 4);
            c.digit_total.ensure(RS_RADIX * 4);
            for (uint32_t shift = idx_bits; shift < key_bits; shift += 8) {  // stable: burn order survives
                radix_hist_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, n, shift, n_blocks, c.hist.as<uint32_t>());
                radix_scan_rows_kernel<<<RS_RADIX, 1024, 0, s>>>(c.hist.as<uint32_t>(), n_blocks,
                                                                c.digit_total.as<uint32_t>());
                radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, s>>>(ka, kb, n, shift, n_blocks, c.hist.as<uint32_t>(),
                                                                     c.digit_total.as<uint32_t>());
                std::swap(ka, kb);
                launches += 3;
            }
        }
    }
    c.task_start.ensure(((size_t)n_tasks + 1) * 4);
    task_index_kernel<<<(n_tasks + 1 + 255) / 256, 256, 0, s>>>(ka, n, idx_bits, n_tasks, c.task_start.as<uint32_t>());
    FillParams F;
    std::memset(&F, 0, sizeof F);
    F.n_tasks = n_tasks;
    F.n_tiles = n_tiles;
    F.tile_w = tile_w;
    F.ncols = (uint32_t)ri.ncols;
    uint64_t bg_bits = 0;
    std::memcpy(&bg_bits, ctx->background, isz);
    replay_for(ctx->dtype, ctx->pixel_fn)(dim3((n_tasks + FILL_WARPS - 1) / FILL_WARPS),
                                          (size_t)FILL_WARPS * FILL_MAX_TILE_W * isz, s, F, ka,
                                          c.task_start.as<uint32_t>(), c.sp_cols.as<unsigned long long>(), c.sp_data.p,
                                          idx_bits, bg_bits, d_out);
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (!out_dev) CUDA_TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    mark_call_end(c, s);
    CUDA_TRY(cudaStreamSynchronize(s));
    if (st) {
        std::memset(st, 0, sizeof *st);
        st->n_records = n;
        st->kernel_launches = launches;
        st->out_bytes = out_bytes;
        st->h2d_bytes = (uint64_t)n * (16 + isz);
        st->d2h_bytes = out_dev ? 0 : out_bytes;
    }
}

// ------------------------------------------------------------------------------------------------
// multi-device calls (one host thread per device, no data-path collective)
// ------------------------------------------------------------------------------------------------
// Dense: device d of D owns a band of raster rows of every band of the output; it is given the parts
// that can write those rows - a part subset of the geometry set, in the same order, so every pixel sees its parts
// in burn order and the result is bit-identical to the single-device one - uploads only those, and copies its rows
// straight into the caller's [B][R][C] array.  Sparse: the stream is ordered band -> geometry -> burn order, so
// devices take contiguous geometry ranges (balanced by a work estimate) and their streams are concatenated by
// offset inside one set of host arrays (SparseSink).  Subsets are cached in the geometry handle.
static void accumulate_stats(rz_stats& a, const rz_stats& b) {
    a.n_parts += b.n_parts;
    a.n_poly_vertices += b.n_poly_vertices;
    a.n_line_vertices += b.n_line_vertices;
    a.n_points += b.n_points;
    a.n_records += b.n_records;
    a.n_crossings += b.n_crossings;
    a.n_tasks += b.n_tasks;
    a.n_mask_words += b.n_mask_words;
    a.h2d_bytes += b.h2d_bytes;
    a.d2h_bytes += b.d2h_bytes;
    a.out_bytes += b.out_bytes;
    a.kernel_launches += b.kernel_launches;
    a.host_syncs += b.host_syncs;
    a.n_windows += b.n_windows;
    a.key_bits = std::max(a.key_bits, b.key_bits);
    a.sort_passes = std::max(a.sort_passes, b.sort_passes);
    a.tile_width = std::max(a.tile_width, b.tile_width);
    a.engine = std::max(a.engine, b.engine);
    a.plan_cached = std::min(a.plan_cached, b.plan_cached);
    // stage times: the slowest device (devices run concurrently)
    a.h2d_ms = std::max(a.h2d_ms, b.h2d_ms);
    a.count_ms = std::max(a.count_ms, b.count_ms);
    a.emit_ms = std::max(a.emit_ms, b.emit_ms);
    a.sort_ms = std::max(a.sort_ms, b.sort_ms);
    a.index_ms = std::max(a.index_ms, b.index_ms);
    a.fill_ms = std::max(a.fill_ms, b.fill_ms);
    a.d2h_ms = std::max(a.d2h_ms, b.d2h_ms);
    a.total_ms = std::max(a.total_ms, b.total_ms);
    a.wall_ms = std::max(a.wall_ms, b.wall_ms);
    a.shard_ms = std::max(a.shard_ms, b.shard_ms);
}

// b ran after a on the same device (the row runs of a one-shot call): counts and times add
static void append_stats(rz_stats& a, const rz_stats& b) {
    const rz_stats first = a;
    a.h2d_ms = a.count_ms = a.emit_ms = a.sort_ms = a.index_ms = a.fill_ms = a.d2h_ms = a.total_ms = a.wall_ms = a.shard_ms = 0.f;
    accumulate_stats(a, b);  // counts add, flags combine
    a.h2d_ms = first.h2d_ms + b.h2d_ms;
    a.count_ms = first.count_ms + b.count_ms;
    a.emit_ms = first.emit_ms + b.emit_ms;
    a.sort_ms = first.sort_ms + b.sort_ms;
    a.index_ms = first.index_ms + b.index_ms;
    a.fill_ms = first.fill_ms + b.fill_ms;
    a.d2h_ms = first.d2h_ms + b.d2h_ms;
    a.total_ms = first.total_ms + b.total_ms;
    a.wall_ms = first.wall_ms + b.wall_ms;
    a.shard_ms = first.shard_ms + b.shard_ms;
}

static uint64_t f64_bits(double v) {
    uint64_t u;
    std::memcpy(&u, &v, 8);
    return u;
}

static std::shared_ptr<rz_geoms> cached_subset(rz_geoms* g, const std::vector<uint64_t>& key, unsigned threads,
                                               const std::function<void(std::vector<uint32_t>&)>& select) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        auto it = g->shards.find(key);
        if (it != g->shards.end()) return it->second;
    }
    std::vector<uint32_t> keep;
    select(keep);
    std::shared_ptr<rz_geoms> sub(subset_parts(g, keep.data(), keep.size(), threads));
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->shards.size() >= 64) g->shards.clear();
    g->shards[key] = sub;
    return sub;
}

// world-y extent of a part: the parts table holds it for polygon parts, other kinds are scanned
static void part_y_extent(const rz_geoms* g, uint32_t p, double& ylo, double& yhi) {
    const int k = g->part_kind[p];
    if (k == RZ_PART_POLYGON) {
        ylo = g->part_ylo[p];
        yhi = g->part_yhi[p];
        return;
    }
    const double inf = std::numeric_limits<double>::infinity();
    ylo = inf;
    yhi = -inf;
    const double* y = g->pool[k].y.data();
    bool odd = false;
    for (uint32_t v = g->part_vbeg[p]; v < g->part_vend[p]; v++) {
        const double a = y[v];
        if (a < ylo) ylo = a;
        if (a > yhi) yhi = a;
        odd |= !(a == a);
    }
    if (odd) {  // NaN ordinates: keep the part everywhere
        ylo = -inf;
        yhi = inf;
    }
}

// Run the calling thread (and the threads it starts) on the CPUs next to a device: its part subset is then first
// touched, and its copies are staged, on the NUMA node the device's PCIe link hangs off.  The CPU list comes from
// sysfs (/sys/bus/pci/devices/<bus id>/local_cpulist); best effort, RZ_NO_AFFINITY=1 turns it off.
static void bind_thread_near_device(int dev) {
    if (std::getenv("RZ_NO_AFFINITY")) return;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return;
    }
    for (char* q = bus; *q; q++) *q = (char)std::tolower((unsigned char)*q);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    std::FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return;
    char line[4096] = {0};
    const bool ok = std::fgets(line, (int)sizeof line, f) != nullptr;
    std::fclose(f);
    if (!ok) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_set = 0;
    for (char* q = line; *q && *q != '\n';) {  // "0-15,64-79"
        char* e = nullptr;
        const long a = std::strtol(q, &e, 10);
        if (e == q) break;
        long b = a;
        q = e;
        if (*q == '-') {
            b = std::strtol(q + 1, &e, 10);
            q = e;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, &set);
            n_set++;
        }
        if (*q == ',') q++;
    }
    if (n_set) (void)pthread_setaffinity_np(pthread_self(), sizeof set, &set);
}

template <typename F> static void run_per_device(int n, F&& body, Error& first_error) {
    std::vector<Error> errors((size_t)n, Error{RZ_OK, ""});
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++)
        th.emplace_back([&, d]() {
            try {
                body(d);
            } catch (const Error& e) {
                errors[d] = e;
            } catch (const std::bad_alloc&) {
                errors[d] = Error{RZ_RUNTIME_ERROR, "Out of host memory."};
            } catch (const std::exception& e) {
                errors[d] = Error{RZ_RUNTIME_ERROR, e.what()};
            }
        });
    for (auto& t : th) t.join();
    first_error = Error{RZ_OK, ""};
    for (auto& e : errors)
        if (e.code != RZ_OK) {
            first_error = e;
            break;
        }
}

static void check_devices(const int32_t* devices, int32_t n_devices, bool allow_repeats = false) {
    if (!devices || n_devices <= 0) throw Error{RZ_VALUE_ERROR, "Empty device list"};
    if (allow_repeats) return;
    for (int32_t i = 0; i < n_devices; i++)
        for (int32_t j = 0; j < i; j++)
            if (devices[i] == devices[j]) throw Error{RZ_VALUE_ERROR, "A device appears twice in the device list"};
}

// parts of g that can write raster rows [b0, b1) of the grid `ri` (margin: pixel rows of slack around a part's extent)
static void select_row_parts(const rz_geoms* g, const rz_raster_info& ri, uint64_t b0, uint64_t b1, double margin,
                             std::vector<uint32_t>& keep) {
    const size_t np = g->part_kind.size();
    keep.reserve(np / 4 + 16);
    for (size_t p = 0; p < np; p++) {
        double ylo, yhi;
        part_y_extent(g, (uint32_t)p, ylo, yhi);
        const double top = (ri.ymax - yhi) / ri.yres, bot = (ri.ymax - ylo) / ri.yres;
        // kept unless certainly outside (comparisons with NaN are false: kept)
        if (bot < (double)b0 - margin || top > (double)b1 + margin) continue;
        keep.push_back((uint32_t)p);
    }
}

static void rasterize_dense_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices, void* out,
                                  rz_stats* st, rz_stats* per_device) {
    const WallClock call_clock;
    // (a repeated device only serialises its shards: single-GPU machines can exercise the sharding that way)
    check_devices(devices, n_devices, std::getenv("RZ_ALLOW_REPEATED_DEVICES") != nullptr);
    if (ctx->flags & (RZ_FLAG_OUT_ON_DEVICE | RZ_FLAG_INPUTS_ON_DEVICE))
        throw Error{RZ_VALUE_ERROR, "Multi-device calls take host inputs and write host memory; use rz_rasterize_dense per device for device buffers"};
    const rz_raster_info& ri = ctx->raster_info;
    validate_lengths(g, ctx);
    uint64_t r0 = ctx->row_begin, r1 = ctx->row_end;
    if (r0 == 0 && r1 == 0) r1 = ri.nrows;
    if (r1 > ri.nrows || (r0 >= r1 && ri.nrows)) throw Error{RZ_VALUE_ERROR, "Invalid row shard"};
    const uint64_t rows = r1 - r0;
    const int D = (int)std::min<uint64_t>((uint64_t)n_devices, std::max<uint64_t>(rows, 1));
    // the devices' host threads cut their subsets at the same time: they share the machine's cores
    const unsigned sub_threads = std::max(2u, std::min(16u, std::max(1u, std::thread::hardware_concurrency()) / (unsigned)D));
    std::vector<rz_stats> S((size_t)D);
    for (auto& x : S) std::memset(&x, 0, sizeof x);
    const double margin = ctx->all_touched ? 3.0 : 2.0;  // pixel rows of slack around a part's extent
    Error err;
    run_per_device(D, [&](int d) {
        const uint64_t b0 = r0 + rows * (uint64_t)d / (uint64_t)D, b1 = r0 + rows * (uint64_t)(d + 1) / (uint64_t)D;
        rz_context c = *ctx;
        c.device = devices[d];
        c.stream = nullptr;  // the caller's stream belongs to one device
        c.row_begin = b0;
        c.row_end = b1;
        if (b1 <= b0) return;
        if (D > 1) bind_thread_near_device(devices[d]);
        rz_geoms* use = g;
        std::shared_ptr<rz_geoms> sub;
        const WallClock shard_clock;
        if (D > 1) {
            const std::vector<uint64_t> key{0, f64_bits(ri.ymax), f64_bits(ri.yres), b0, b1, (uint64_t)margin};
            sub = cached_subset(g, key, sub_threads, [&](std::vector<uint32_t>& keep) { select_row_parts(g, ri, b0, b1, margin, keep); });
            use = sub.get();
        }
        const float shard_ms = shard_clock.ms();
        const DenseExtra ex{rows, b0 - r0};
        rasterize_dense(use, &c, out, &S[d], &ex);
        S[d].shard_ms = shard_ms;
        S[d].wall_ms += shard_ms;
    }, err);
    if (err.code != RZ_OK) throw err;
    rz_stats A = S[0];
    for (int d = 1; d < D; d++) accumulate_stats(A, S[d]);
    A.wall_ms = call_clock.ms();
    if (st) *st = A;
    if (per_device)
        for (int d = 0; d < n_devices; d++) {
            if (d < D) per_device[d] = S[d];
            else std::memset(&per_device[d], 0, sizeof(rz_stats));
        }
}

struct SparseGather : SparseSink {
    std::mutex mu;
    std::condition_variable cv;
    int n_dev = 0, arrived = 0;
    bool failed = false, ready = false;
    size_t isz = 0;
    std::vector<std::vector<uint64_t>> counts;  // [device][band]
    std::vector<uint64_t> band_base;
    rz_sparse* out = nullptr;
    struct View : SparseSink {
        SparseGather* g;
        int d;
        void place(const std::vector<uint64_t>& c, std::vector<uint64_t>& off, char*& r, char*& cc, char*& da) override {
            g->place_from(d, c, off, r, cc, da);
        }
    };
    void place(const std::vector<uint64_t>&, std::vector<uint64_t>&, char*&, char*&, char*&) override {}
    void place_from(int d, const std::vector<uint64_t>& c, std::vector<uint64_t>& off, char*& r, char*& cc, char*& da) {
        std::unique_lock<std::mutex> lk(mu);
        counts[d] = c;
        arrived++;
        if (arrived == n_dev && !failed) {
            const size_t nb = c.size();
            band_base.assign(nb + 1, 0);
            for (size_t b = 0; b < nb; b++) {
                uint64_t t = 0;
                for (int e = 0; e < n_dev; e++) t += b < counts[e].size() ? counts[e][b] : 0;
                out->counts[b] = t;
                band_base[b + 1] = band_base[b] + t;
            }
            const uint64_t total = band_base[nb];
            try {
                out->len = total;
                // (written by every device of the call: pages on the memory nodes the devices hang off)
                out->rows = g_host_pool.get(total * 8, true);
                out->cols = g_host_pool.get(total * 8, true);
                out->data = g_host_pool.get(total * isz, true);
                ready = true;
            } catch (const std::bad_alloc&) {
                failed = true;
            }
            cv.notify_all();
        } else {
            cv.wait(lk, [&]() { return ready || failed; });
        }
        if (failed) throw Error{RZ_RUNTIME_ERROR, "Multi-device sparse call aborted: a device failed or host memory ran out."};
        off.assign(c.size(), 0);
        for (size_t b = 0; b < c.size(); b++) {
            uint64_t o = band_base[b];
            for (int e = 0; e < d; e++) o += counts[e][b];
            off[b] = o;
        }
        r = (char*)out->rows.p;
        cc = (char*)out->cols.p;
        da = (char*)out->data.p;
    }
    void fail() {
        std::lock_guard<std::mutex> lk(mu);
        failed = true;
        cv.notify_all();
    }
};

static void rasterize_sparse_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices,
                                   rz_sparse* out, rz_stats* st, rz_stats* per_device) {
    const WallClock call_clock;
    check_devices(devices, n_devices);
    if (ctx->flags & RZ_FLAG_INPUTS_ON_DEVICE)
        throw Error{RZ_VALUE_ERROR, "Multi-device calls take host inputs; use rz_rasterize_sparse per device for device buffers"};
    validate_lengths(g, ctx);
    const rz_raster_info& ri = ctx->raster_info;
    const size_t isz = dtype_size(ctx->dtype);
    if (!isz) throw Error{RZ_VALUE_ERROR, "Unsupported dtype"};
    const uint32_t n_bands = ctx->band_of_geom ? (uint32_t)std::max(ctx->n_bands, 0) : 1u;
    const size_t np = g->part_kind.size();
    const int D = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_devices, g->n_geoms));
    const unsigned sub_threads = std::max(2u, std::min(16u, std::max(1u, std::thread::hardware_concurrency()) / (unsigned)D));
    // ---- contiguous geometry ranges of equal estimated work (a prefix sum of per-geometry cost) ----
    // cost of a part: its vertices (setup) + for polygons the pixels of its box (fill, ~half of them burned) and two
    // crossings per row; for lines the longer side of each part's box is unknown without a scan: vertices x 16
    std::vector<double> cost(np);
    for (size_t p = 0; p < np; p++) {
        const double nv = (double)(g->part_vend[p] - g->part_vbeg[p]);
        double c = 4.0 * nv + 8.0;
        if (g->part_kind[p] == RZ_PART_POLYGON) {
            double w = (g->part_xhi[p] - g->part_xlo[p]) / ri.xres, h = (g->part_yhi[p] - g->part_ylo[p]) / ri.yres;
            w = std::min(std::max(w, 0.0), (double)ri.ncols);
            h = std::min(std::max(h, 0.0), (double)ri.nrows);
            if (w == w && h == h) c += 0.5 * w * h + 4.0 * h;
        } else if (g->part_kind[p] == RZ_PART_LINE) {
            c += 16.0 * nv;
        }
        cost[p] = c;
    }
    double total_cost = 0;
    for (double c : cost) total_cost += c;
    // cut at geometry borders: part p starts a geometry when part_geom changes
    std::vector<size_t> cut((size_t)D + 1, np);
    cut[0] = 0;
    {
        double acc = 0;
        int k = 1;
        for (size_t p = 0; p < np && k < D; p++) {
            acc += cost[p];
            const bool geom_end = p + 1 == np || g->part_geom[p + 1] != g->part_geom[p];
            if (geom_end && acc >= total_cost * (double)k / (double)D) cut[(size_t)k++] = p + 1;
        }
    }
    out->counts.assign(n_bands, 0);
    SparseGather gather;
    gather.n_dev = D;
    gather.isz = isz;
    gather.counts.assign((size_t)D, std::vector<uint64_t>(n_bands, 0));
    gather.out = out;
    std::vector<rz_stats> S((size_t)D);
    for (auto& x : S) std::memset(&x, 0, sizeof x);
    Error err;
    run_per_device(D, [&](int d) {
        SparseGather::View view;
        view.g = &gather;
        view.d = d;
        try {
            rz_context c = *ctx;
            c.device = devices[d];
            c.stream = nullptr;
            if (D > 1) bind_thread_near_device(devices[d]);
            rz_geoms* use = g;
            std::shared_ptr<rz_geoms> sub;
            const WallClock shard_clock;
            if (D > 1) {
                const std::vector<uint64_t> key{1, (uint64_t)cut[(size_t)d], (uint64_t)cut[(size_t)d + 1]};
                sub = cached_subset(g, key, sub_threads, [&](std::vector<uint32_t>& keep) {
                    keep.resize(cut[(size_t)d + 1] - cut[(size_t)d]);
                    for (size_t i = 0; i < keep.size(); i++) keep[i] = (uint32_t)(cut[(size_t)d] + i);
                });
                use = sub.get();
            }
            const float shard_ms = shard_clock.ms();
            rz_sparse local;  // counts only: the triplets go to the gathered arrays
            rasterize_sparse(use, &c, &local, &S[d], &view);
            S[d].shard_ms = shard_ms;
            S[d].wall_ms += shard_ms;
        } catch (...) {
            gather.fail();
            throw;
        }
    }, err);
    if (err.code != RZ_OK) throw err;
    rz_stats A = S[0];
    for (int d = 1; d < D; d++) accumulate_stats(A, S[d]);
    A.wall_ms = call_clock.ms();
    if (st) *st = A;
    if (per_device)
        for (int d = 0; d < n_devices; d++) {
            if (d < D) per_device[d] = S[d];
            else std::memset(&per_device[d], 0, sizeof(rz_stats));
        }
}

}  // namespace rz

// ================================================================================================
// C ABI
// ================================================================================================
using rz::Error;
using rz::set_err;

rz_geoms::~rz_geoms() {
    for (auto& kv : dev) delete kv.second;
    dev.clear();
    rz::unpin_host(this);
}

template <typename F> static int guarded(char* err, size_t errlen, F f) {
    try {
        f();
        return RZ_OK;
    } catch (const Error& e) {
        set_err(err, errlen, e.msg);
        return e.code;
    } catch (const std::bad_alloc&) {
        set_err(err, errlen, "Out of host memory.");
        return RZ_RUNTIME_ERROR;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return RZ_RUNTIME_ERROR;
    }
}

extern "C" {

// Parse geometries [i0, i1) into g (serial).
static void parse_wkb_range(rz_geoms* g, const uint8_t* const* bufs, const uint64_t* lens, uint64_t i0, uint64_t i1) {
    rz::Flattener f(g);
    // A vertex takes at least 16 WKB bytes: reserve the pool of the first geometry's kind once instead of growing
    // it by doubling (untouched reserve costs address space only).
    if (i1 > i0 && lens[i0] >= 5) {
        const uint8_t* b0 = bufs[i0];
        uint32_t t = 0;
        for (int k = 0; k < 4; k++) t |= (uint32_t)b0[1 + (b0[0] ? k : 3 - k)] << (8 * k);
        t = (t & 0x0fffffffu) % 1000;
        const int kind = (t == 3 || t == 6) ? RZ_PART_POLYGON : (t == 2 || t == 5) ? RZ_PART_LINE : (t == 1 || t == 4) ? RZ_PART_POINT : -1;
        uint64_t total = 0;
        for (uint64_t i = i0; i < i1; i++) total += lens[i];
        if (kind >= 0 && total / 16 < 0xfffffff0ull) {
            g->pool[kind].x.reserve(total / 16 + 16);
            g->pool[kind].y.reserve(total / 16 + 16);
            g->pool[kind].tag.reserve(total / 16 + 16);
        }
    }
    for (uint64_t i = i0; i < i1; i++) {
        bool keep = false;
        f.begin_geometry();
        if (!rz::read_wkb(bufs[i], (size_t)lens[i], f, &keep))
            throw Error{RZ_RUNTIME_ERROR, f.ok() ? "Cannot parse geometry. Check that the WKB bytes are valid." : f.error()};
        f.end_geometry(keep);
    }
}

// Where chunk c lands in the merged geometry set.
struct ChunkPlace {
    size_t part_off, geom_off, pool_off[3];
};
// Copy chunk c's pools into their place in g (called concurrently for different chunks: disjoint ranges);
// part ids inside the tags are shifted by the chunk's part offset.
static void place_pools(rz_geoms* g, rz_geoms* c, const ChunkPlace& at) {
    for (int k = 0; k < 3; k++) {
        rz::Pool& d = g->pool[k];
        rz::Pool& s = c->pool[k];
        const size_t o = at.pool_off[k], n = s.size();
        if (!n) continue;
        std::memcpy(d.x.data() + o, s.x.data(), n * 8);
        std::memcpy(d.y.data() + o, s.y.data(), n * 8);
        for (size_t i = 0; i < n; i++)
            d.tag[o + i] = (s.tag[i] & ~rz::TAG_PART_MASK) | (uint32_t)((s.tag[i] & rz::TAG_PART_MASK) + at.part_off);
        rz::Pool().x.swap(s.x);  // release the chunk's copy right away
        rz::Pool().y.swap(s.y);
        rz::Pool().tag.swap(s.tag);
    }
}
// ... and its sequence lists, parts table and bounds (serial, in chunk order)
static void place_parts(rz_geoms* g, rz_geoms* c, const ChunkPlace& at) {
    for (int k = 0; k < 3; k++) {
        for (uint32_t e : c->pool[k].seq_end) g->pool[k].seq_end.push_back((uint32_t)(e + at.pool_off[k]));
        g->pool[k].seq_closed.insert(g->pool[k].seq_closed.end(), c->pool[k].seq_closed.begin(), c->pool[k].seq_closed.end());
    }
    for (size_t p = 0; p < c->part_kind.size(); p++) {
        const int k = c->part_kind[p];
        g->part_kind.push_back(c->part_kind[p]);
        g->part_geom.push_back(c->part_geom[p] + at.geom_off);
        g->part_xlo.push_back(c->part_xlo[p]);
        g->part_xhi.push_back(c->part_xhi[p]);
        g->part_ylo.push_back(c->part_ylo[p]);
        g->part_yhi.push_back(c->part_yhi[p]);
        g->part_vbeg.push_back((uint32_t)(c->part_vbeg[p] + at.pool_off[k]));
        g->part_vend.push_back((uint32_t)(c->part_vend[p] + at.pool_off[k]));
    }
    if (c->has_bounds) {  // same fold as Flattener::end_geometry (rust/src/geo/raster.rs:81-84)
        if (!g->has_bounds) {
            std::memcpy(g->bounds, c->bounds, sizeof g->bounds);
            g->has_bounds = true;
        } else {
            g->bounds[0] = std::fmin(g->bounds[0], c->bounds[0]);
            g->bounds[1] = std::fmin(g->bounds[1], c->bounds[1]);
            g->bounds[2] = std::fmax(g->bounds[2], c->bounds[2]);
            g->bounds[3] = std::fmax(g->bounds[3], c->bounds[3]);
        }
    }
    g->n_geoms += c->n_geoms;
}

rz_geoms* rz_geoms_from_wkb(const uint8_t* const* bufs, const uint64_t* lens, uint64_t n, char* err, size_t errlen) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    int rc = guarded(err, errlen, [&]() {
        // Large inputs are parsed by several threads, each flattening a contiguous range of geometries (ranges of
        // equal byte counts) into its own pools; the chunks are then appended in order, which gives exactly the
        // serial result.  (The reference parses on one core; at config-4 scale the 3.2 GB of WKB would otherwise
        // take longer to ingest than the whole rasterisation takes end to end.)
        uint64_t total = 0;
        for (uint64_t i = 0; i < n; i++) total += lens[i];
        unsigned threads = std::min<unsigned>({std::max(1u, std::thread::hardware_concurrency()), 16u,
                                               (unsigned)(total >> 24) + 1u, (unsigned)(n / 1024) + 1u});
        if (const char* e = std::getenv("RZ_PARSE_THREADS")) threads = std::max(1, std::atoi(e));
        if (threads <= 1) {
            parse_wkb_range(g.get(), bufs, lens, 0, n);
        } else {
            std::vector<uint64_t> cut(threads + 1, n);
            cut[0] = 0;
            uint64_t acc = 0;
            unsigned k = 1;
            for (uint64_t i = 0; i < n && k < threads; i++) {
                acc += lens[i];
                if (acc >= total / threads * k) cut[k++] = i + 1;
            }
            std::vector<std::unique_ptr<rz_geoms>> chunk(threads);
            std::vector<Error> errors(threads, Error{RZ_OK, ""});
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < threads; t++) {
                chunk[t].reset(new rz_geoms());
                pool.emplace_back([&, t]() {
                    try {
                        parse_wkb_range(chunk[t].get(), bufs, lens, cut[t], cut[t + 1]);
                    } catch (const Error& e) {
                        errors[t] = e;
                    } catch (const std::bad_alloc&) {
                        errors[t] = Error{RZ_RUNTIME_ERROR, "Out of host memory."};
                    }
                });
            }
            for (auto& th : pool) th.join();
            for (unsigned t = 0; t < threads; t++)
                if (errors[t].code != RZ_OK) throw errors[t];  // the first failing range, like the serial walk
            std::vector<ChunkPlace> at(threads);
            ChunkPlace run{0, 0, {0, 0, 0}};
            for (unsigned t = 0; t < threads; t++) {
                at[t] = run;
                run.part_off += chunk[t]->part_kind.size();
                run.geom_off += chunk[t]->n_geoms;
                for (int p = 0; p < 3; p++) run.pool_off[p] += chunk[t]->pool[p].size();
            }
            if (run.part_off >= rz::TAG_PART_MASK) throw Error{RZ_RUNTIME_ERROR, "Too many geometry parts (limit 2^30 - 1)."};
            for (int p = 0; p < 3; p++) {
                if (run.pool_off[p] >= 0xfffffff0ull) throw Error{RZ_RUNTIME_ERROR, "Too many vertices (limit 2^32 per pool)."};
                g->pool[p].x.resize(run.pool_off[p]);  // default-initialised: first touched by the copies below
                g->pool[p].y.resize(run.pool_off[p]);
                g->pool[p].tag.resize(run.pool_off[p]);
            }
            pool.clear();
            for (unsigned t = 0; t < threads; t++)
                pool.emplace_back([&, t]() { place_pools(g.get(), chunk[t].get(), at[t]); });
            for (auto& th : pool) th.join();
            for (unsigned t = 0; t < threads; t++) {
                place_parts(g.get(), chunk[t].get(), at[t]);
                chunk[t].reset();
            }
        }
        // python/src/geo/parse_geometry.rs:20-28 (bail_if_empty_geoms)
        if (g->n_geoms == 0)
            throw Error{RZ_VALUE_ERROR, "Could not parse geometry. Only WKT or WKB formats are supported."};
        rz::finish_geoms(g.get());
    });
    return rc == RZ_OK ? g.release() : nullptr;
}

rz_geoms* rz_geoms_from_wkt(const char* const* strs, uint64_t n, char* err, size_t errlen) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    int rc = guarded(err, errlen, [&]() {
        rz::Flattener f(g.get());
        for (uint64_t i = 0; i < n; i++) {
            bool keep = false;
            f.begin_geometry();
            if (!rz::read_wkt(strs[i], f, &keep))
                throw Error{RZ_RUNTIME_ERROR, f.ok() ? "Cannot parse geometry. Check that the WKT is valid." : f.error()};
            f.end_geometry(keep);
        }
        if (g->n_geoms == 0)
            throw Error{RZ_VALUE_ERROR, "Could not parse geometry. Only WKT or WKB formats are supported."};
        rz::finish_geoms(g.get());
    });
    return rc == RZ_OK ? g.release() : nullptr;
}

rz_geoms* rz_geoms_from_soa(const rz_geom_soa* soa, char* err, size_t errlen) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    int rc = guarded(err, errlen, [&]() {
        // every size is known from the offsets: a counting pass gives each thread its exact place in the final
        // (page-locked, recycled) pools, then copy + extents run in one sweep - memcpy speed, no intermediate copy
        unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* e = std::getenv("RZ_PARSE_THREADS")) threads = std::max(1, std::atoi(e));
        std::string msg;
        const int code = rz::flatten_soa(soa, g.get(), threads, msg);
        if (code != RZ_OK) throw Error{code, msg};
    });
    return rc == RZ_OK ? g.release() : nullptr;
}

// Flatten and upload at the same time: the worker threads hand every few megabytes of finished pool to the copy engine
// (the pools are page-locked blocks), so the H2D transfer of a 3.3 GB geometry set (60 ms) hides behind its flattening
// (75 ms) instead of following it.
namespace {
struct UploadWhileFlattening {
    rz::DeviceCtx* c = nullptr;
    rz::DeviceGeoms* d = nullptr;
    rz_geoms* g = nullptr;
    std::atomic<int> failed{0};
    static void on_sized(void* p, rz_geoms* g) {
        auto* u = static_cast<UploadWhileFlattening*>(p);
        u->g = g;
        for (int k = 0; k < 3; k++)
            if (g->pool[k].size() >= 0xfffffff0ull) return;  // reported by the normal path
        u->d = new rz::DeviceGeoms();
        u->d->dev = u->c->dev;
        try {
            rz::alloc_device_geoms(g, u->d);
        } catch (const Error&) {
            delete u->d;
            u->d = nullptr;  // the first rasterize call uploads (and reports the problem)
        }
    }
    static void on_range(void* p, int kind, uint64_t v0, uint64_t v1) {
        auto* u = static_cast<UploadWhileFlattening*>(p);
        if (!u->d || u->failed.load(std::memory_order_relaxed)) return;
        const size_t n = (size_t)(v1 - v0) * 8;
        if (cudaSetDevice(u->c->dev) != cudaSuccess ||
            cudaMemcpyAsync(u->d->x[kind] + v0, u->g->pool[kind].x.data() + v0, n, cudaMemcpyHostToDevice, u->c->upload_stream) != cudaSuccess ||
            cudaMemcpyAsync(u->d->y[kind] + v0, u->g->pool[kind].y.data() + v0, n, cudaMemcpyHostToDevice, u->c->upload_stream) != cudaSuccess) {
            (void)cudaGetLastError();
            u->failed.store(1);
        }
    }
};
}  // namespace

// flatten `soa` (all parts, or those flagged in keep_part) with the pools going to `device` while they are written
static std::unique_ptr<rz_geoms> flatten_to_device(const rz_geom_soa* soa, const uint8_t* keep_part, int device,
                                                   unsigned threads) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    rz::DeviceGuard guard;
    rz::DeviceCtx& c = rz::device_ctx(device);
    // no context mutex: nothing of the context's scratch is touched, the copies go to their own stream - a call in
    // flight on this device (the previous rows of a one-shot call) keeps running
    CUDA_TRY(cudaSetDevice(c.dev));
    UploadWhileFlattening up;
    up.c = &c;
    rz::FlattenHooks hooks;
    hooks.ctx = &up;
    hooks.on_sized = UploadWhileFlattening::on_sized;
    hooks.on_range = UploadWhileFlattening::on_range;
    std::string msg;
    const int code = rz::flatten_soa(soa, g.get(), threads, msg, &hooks, keep_part);
    if (code != RZ_OK) {
        if (up.d) {
            cudaStreamSynchronize(c.upload_stream);
            delete up.d;
        }
        throw Error{code, msg};
    }
    if (up.d && !up.failed.load()) {  // parts table, sequence lists, vertex tags; then the set is resident
        rz::geoms_on_device(g.get(), c, c.upload_stream, false, nullptr, up.d);
    } else if (up.d) {
        cudaStreamSynchronize(c.upload_stream);
        delete up.d;
    }
    return g;
}

static unsigned host_threads() {
    unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = std::getenv("RZ_PARSE_THREADS")) threads = std::max(1, std::atoi(e));
    return threads;
}

rz_geoms* rz_geoms_from_soa_to(const rz_geom_soa* soa, int device, char* err, size_t errlen) {
    std::unique_ptr<rz_geoms> g;
    int rc = guarded(err, errlen, [&]() { g = flatten_to_device(soa, nullptr, device, host_threads()); });
    return rc == RZ_OK ? g.release() : nullptr;
}

// keep[p] = part p can write raster rows [b0, b1) of grid `ri` (same rule as select_row_parts, from raw extents)
static void keep_row_parts(const rz_raster_info& ri, const double* ylo, const double* yhi, uint64_t n_parts, uint64_t b0,
                           uint64_t b1, double margin, uint8_t* keep) {
    for (uint64_t p = 0; p < n_parts; p++) {
        const double top = (ri.ymax - yhi[p]) / ri.yres, bot = (ri.ymax - ylo[p]) / ri.yres;
        keep[p] = !(bot < (double)b0 - margin || top > (double)b1 + margin);  // comparisons with NaN are false: kept
    }
}

rz_geoms* rz_geoms_from_soa_rows(const rz_geom_soa* soa, const rz_raster_info* ri, uint64_t row_begin, uint64_t row_end,
                                 int all_touched, char* err, size_t errlen) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    int rc = guarded(err, errlen, [&]() {
        if (row_end > ri->nrows || row_begin >= row_end) throw Error{RZ_VALUE_ERROR, "Invalid row shard"};
        const uint64_t NP = soa->n_parts;
        std::vector<double> ylo(NP), yhi(NP);
        std::vector<uint8_t> keep(NP);
        std::string msg;
        int code = rz::soa_part_y_extents(soa, host_threads(), ylo.data(), yhi.data(), msg);
        if (code != RZ_OK) throw Error{code, msg};
        keep_row_parts(*ri, ylo.data(), yhi.data(), NP, row_begin, row_end, all_touched ? 3.0 : 2.0, keep.data());
        code = rz::flatten_soa(soa, g.get(), host_threads(), msg, nullptr, keep.data());
        if (code != RZ_OK) throw Error{code, msg};
    });
    return rc == RZ_OK ? g.release() : nullptr;
}

// Rows [b0, b1) of a one-shot call on one device, as a two-stage pipeline over runs of rows: while run k burns and
// copies back (PCIe device -> host, the long stage), the calling thread flattens and uploads the parts of run k+1
// (host -> device: the other direction of the link).  The first copy-back starts after 1/n_runs of the flattening.
static void one_shot_rows(const rz_geom_soa* soa, const rz_context* ctx, int device, uint64_t b0, uint64_t b1, uint64_t r0,
                          uint64_t rows, const double* ylo, const double* yhi, int n_runs, unsigned threads, void* out,
                          rz_stats& S) {
    const rz_raster_info& ri = ctx->raster_info;
    const uint64_t NP = soa->n_parts;
    const double margin = ctx->all_touched ? 3.0 : 2.0;
    std::vector<uint8_t> keep;
    std::future<rz_stats> burn;
    bool have = false;
    float flatten_ms = 0.f;
    auto collect = [&]() {
        const rz_stats st = burn.get();  // rethrows what the burn threw
        if (have) rz::append_stats(S, st);
        else S = st;
        have = true;
    };
    try {
        for (int k = 0; k < n_runs; k++) {
            const uint64_t a0 = b0 + (b1 - b0) * (uint64_t)k / (uint64_t)n_runs, a1 = b0 + (b1 - b0) * (uint64_t)(k + 1) / (uint64_t)n_runs;
            if (a1 <= a0) continue;
            const rz::WallClock clock;
            const uint8_t* keep_ptr = nullptr;
            if (ylo) {  // (no extents: one run over every row of the call - every part is kept)
                keep.resize(NP);
                keep_row_parts(ri, ylo, yhi, NP, a0, a1, margin, keep.data());
                keep_ptr = keep.data();
            }
            std::shared_ptr<rz_geoms> g(flatten_to_device(soa, keep_ptr, device, threads).release());
            flatten_ms += clock.ms();
            if (burn.valid()) collect();
            rz_context c = *ctx;
            c.device = device;
            c.stream = nullptr;
            c.row_begin = a0;
            c.row_end = a1;
            const rz::DenseExtra ex{rows, a0 - r0};
            if (n_runs == 1) {  // nothing to overlap with: burn on the calling thread
                rz::rasterize_dense(g.get(), &c, out, &S, &ex);
                have = true;
                break;
            }
            burn = std::async(std::launch::async, [g, c, ex, out, device]() {
                rz::bind_thread_near_device(device);
                rz_stats st;
                std::memset(&st, 0, sizeof st);
                rz::rasterize_dense(g.get(), &c, out, &st, &ex);
                return st;
            });
        }
        if (burn.valid()) collect();
    } catch (...) {
        if (burn.valid()) burn.wait();  // the burn in flight reads the caller's arrays: let it finish
        throw;
    }
    S.shard_ms = flatten_ms;
}

// how many runs of rows a device's share of a one-shot call is cut into.  ONE unless RZ_ONE_SHOT_RUNS says otherwise:
// measured on config 4 / 1 B200 (profiles/r2_one_shot_runs.txt) the flattening of run k+1 next to the copy-back of
// run k takes 2.3x as long and slows the copy by 6 % (both are bound by the host's memory system), so 4 runs end at
// 398 ms against 403 ms for one - within run-to-run noise - and 8 runs are slower.
static int one_shot_runs(const rz_geom_soa*, const rz_context*, uint64_t) {
    if (const char* e = std::getenv("RZ_ONE_SHOT_RUNS")) return std::max(1, std::min(64, std::atoi(e)));
    return 1;
}

// DenseArray::build in one call (rust/src/rasterize.rs:77-115): the caller's geometries (SoA), the context, the
// devices, the caller's array.  One parallel read of the y ordinates gives every part's extent; every device's host
// thread then flattens ONLY the parts of its rows straight out of the caller's arrays (no full flattened copy, no
// second subset copy: 8.2 GB of host traffic instead of 13 GB for config 4 on several devices) with the upload
// overlapped, burns them and copies them into `out` (one_shot_rows; optionally a run of rows at a time).
int rz_rasterize_dense_soa(const rz_geom_soa* soa, const rz_context* ctx, const int32_t* devices, int32_t n_devices,
                           void* out, rz_stats* stats, rz_stats* per_device, char* err, size_t errlen) {
    return guarded(err, errlen, [&]() {
        const rz::WallClock call_clock;
        rz::check_devices(devices, n_devices, std::getenv("RZ_ALLOW_REPEATED_DEVICES") != nullptr);
        if (ctx->flags & (RZ_FLAG_OUT_ON_DEVICE | RZ_FLAG_INPUTS_ON_DEVICE))
            throw Error{RZ_VALUE_ERROR, "One-shot calls take host inputs and write host memory"};
        const rz_raster_info& ri = ctx->raster_info;
        uint64_t r0 = ctx->row_begin, r1 = ctx->row_end;
        if (r0 == 0 && r1 == 0) r1 = ri.nrows;
        if (r1 > ri.nrows || (r0 >= r1 && ri.nrows)) throw Error{RZ_VALUE_ERROR, "Invalid row shard"};
        const uint64_t rows = r1 - r0;
        const int D = (int)std::min<uint64_t>((uint64_t)n_devices, std::max<uint64_t>(rows, 1));
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        std::vector<rz_stats> S((size_t)D);
        for (auto& x : S) std::memset(&x, 0, sizeof x);
        const int n_runs = one_shot_runs(soa, ctx, (rows + (uint64_t)D - 1) / (uint64_t)D);
        const uint64_t NP = soa->n_parts;
        std::vector<double> ylo, yhi;
        float extents_ms = 0.f;
        if (D > 1 || n_runs > 1) {
            ylo.resize(NP);
            yhi.resize(NP);
            std::string msg;
            const int code = rz::soa_part_y_extents(soa, std::min(32u, hw), ylo.data(), yhi.data(), msg);
            if (code != RZ_OK) throw Error{code, msg};
            extents_ms = call_clock.ms();
        }
        const double* pylo = ylo.empty() ? nullptr : ylo.data();
        const double* pyhi = yhi.empty() ? nullptr : yhi.data();
        if (D == 1) {
            one_shot_rows(soa, ctx, devices[0], r0, r1, r0, rows, (n_runs > 1 ? pylo : nullptr), pyhi, n_runs, host_threads(), out, S[0]);
            S[0].shard_ms += extents_ms;
        } else {
            const unsigned sub_threads = std::max(2u, std::min(16u, hw / (unsigned)D));
            Error first;
            rz::run_per_device(D, [&](int d) {
                const uint64_t b0 = r0 + rows * (uint64_t)d / (uint64_t)D, b1 = r0 + rows * (uint64_t)(d + 1) / (uint64_t)D;
                if (b1 <= b0) return;
                rz::bind_thread_near_device(devices[d]);
                const rz::WallClock shard_clock;
                one_shot_rows(soa, ctx, devices[d], b0, b1, r0, rows, pylo, pyhi, n_runs, sub_threads, out, S[d]);
                S[d].shard_ms += extents_ms;
                S[d].wall_ms = shard_clock.ms();
            }, first);
            if (first.code != RZ_OK) throw first;
        }
        rz_stats A = S[0];
        for (int d = 1; d < D; d++) rz::accumulate_stats(A, S[d]);
        A.wall_ms = call_clock.ms();
        if (stats) *stats = A;
        if (per_device)
            for (int d = 0; d < n_devices; d++) {
                if (d < D) per_device[d] = S[d];
                else std::memset(&per_device[d], 0, sizeof(rz_stats));
            }
    });
}

uint64_t rz_geoms_len(const rz_geoms* g) { return g->n_geoms; }
uint64_t rz_geoms_n_parts(const rz_geoms* g) { return g->part_kind.size(); }
uint64_t rz_geoms_n_coords(const rz_geoms* g) { return g->pool[0].size() + g->pool[1].size() + g->pool[2].size(); }

int rz_geoms_bounds(const rz_geoms* g, double out[4]) {
    if (!g->has_bounds) return RZ_RUNTIME_ERROR;
    std::memcpy(out, g->bounds, sizeof g->bounds);
    return RZ_OK;
}

int rz_geoms_upload(rz_geoms* g, int device, char* err, size_t errlen) {
    return guarded(err, errlen, [&]() {
        rz::DeviceGuard guard;
        rz::DeviceCtx& c = rz::device_ctx(device);
        std::lock_guard<std::mutex> lk(c.mu);
        CUDA_TRY(cudaSetDevice(c.dev));
        rz::geoms_on_device(g, c, c.stream, false, nullptr);
    });
}

void rz_geoms_evict(rz_geoms* g) {
    // a call in flight on a device holds that device's context mutex while it uses the cached copy: take it
    // before deleting the copy
    rz::DeviceGuard guard;
    std::vector<int> devs;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        for (auto& kv : g->dev) devs.push_back(kv.first);
    }
    for (int d : devs) {
        rz::DeviceCtx* c = nullptr;
        try {
            c = &rz::device_ctx(d);
        } catch (const Error&) {
            continue;
        }
        std::lock_guard<std::mutex> lc(c->mu);
        std::lock_guard<std::mutex> lk(g->mu);
        auto it = g->dev.find(d);
        if (it != g->dev.end()) {
            delete it->second;
            g->dev.erase(it);
        }
    }
}

void rz_geoms_free(rz_geoms* g) { delete g; }

rz_geoms* rz_geoms_row_shard(const rz_geoms* g, const rz_raster_info* ri, uint64_t row_begin, uint64_t row_end,
                             int all_touched, char* err, size_t errlen) {
    rz_geoms* out = nullptr;
    int rc = guarded(err, errlen, [&]() {
        if (row_end > ri->nrows || row_begin >= row_end) throw Error{RZ_VALUE_ERROR, "Invalid row shard"};
        std::vector<uint32_t> keep;
        rz::select_row_parts(g, *ri, row_begin, row_end, all_touched ? 3.0 : 2.0, keep);
        unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* e = std::getenv("RZ_PARSE_THREADS")) threads = std::max(1, std::atoi(e));
        out = rz::subset_parts(g, keep.data(), keep.size(), threads);
    });
    return rc == RZ_OK ? out : nullptr;
}

const uint8_t* rz_geoms_part_kind(const rz_geoms* g) { return g->part_kind.data(); }
const uint64_t* rz_geoms_part_geom(const rz_geoms* g) { return g->part_geom.data(); }
uint64_t rz_geoms_pool_len(const rz_geoms* g, int kind) { return kind >= 0 && kind < 3 ? g->pool[kind].size() : 0; }
const double* rz_geoms_pool_x(const rz_geoms* g, int kind) { return g->pool[kind].x.data(); }
const double* rz_geoms_pool_y(const rz_geoms* g, int kind) { return g->pool[kind].y.data(); }
const uint32_t* rz_geoms_pool_tag(const rz_geoms* g, int kind) {
    rz::ensure_tags(const_cast<rz_geoms*>(g), kind);  // not kept by every flattening path: the device never needs them
    return g->pool[kind].tag.data();
}

int rz_raster_info_build(const rz_raw_raster_info* raw, const rz_geoms* g, rz_raster_info* out, char* err,
                         size_t errlen) {
    std::string msg;
    int rc = rz::build_raster_info(raw, g, out, msg);
    if (rc) set_err(err, errlen, msg);
    return rc;
}

int64_t rz_group_keys(const char* const* keys, uint64_t n, int32_t* band_of_geom, uint64_t* band_first) {
    return rz::group_keys(keys, n, band_of_geom, band_first);
}

int rz_rasterize_dense(rz_geoms* g, const rz_context* ctx, void* out, rz_stats* stats, char* err, size_t errlen) {
    return guarded(err, errlen, [&]() { rz::rasterize_dense(g, ctx, out, stats); });
}

int rz_rasterize_sparse(rz_geoms* g, const rz_context* ctx, rz_sparse** out, rz_stats* stats, char* err,
                        size_t errlen) {
    std::unique_ptr<rz_sparse> sp(new rz_sparse());
    int rc = guarded(err, errlen, [&]() { rz::rasterize_sparse(g, ctx, sp.get(), stats); });
    if (rc == RZ_OK) *out = sp.release();
    return rc;
}
int rz_rasterize_dense_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices, void* out,
                             rz_stats* stats, rz_stats* per_device, char* err, size_t errlen) {
    return guarded(err, errlen, [&]() { rz::rasterize_dense_multi(g, ctx, devices, n_devices, out, stats, per_device); });
}
int rz_rasterize_sparse_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices,
                              rz_sparse** out, rz_stats* stats, rz_stats* per_device, char* err, size_t errlen) {
    std::unique_ptr<rz_sparse> sp(new rz_sparse());
    int rc = guarded(err, errlen,
                     [&]() { rz::rasterize_sparse_multi(g, ctx, devices, n_devices, sp.get(), stats, per_device); });
    if (rc == RZ_OK) *out = sp.release();
    return rc;
}
uint64_t rz_sparse_len(const rz_sparse* s) { return s->len; }
uint64_t rz_sparse_n_bands(const rz_sparse* s) { return s->counts.size(); }
const uint64_t* rz_sparse_rows(const rz_sparse* s) { return (const uint64_t*)s->rows.p; }
const uint64_t* rz_sparse_cols(const rz_sparse* s) { return (const uint64_t*)s->cols.p; }
const void* rz_sparse_data(const rz_sparse* s) { return s->data.p; }
const uint64_t* rz_sparse_counts(const rz_sparse* s) { return s->counts.data(); }
void rz_sparse_free(rz_sparse* s) { delete s; }
int rz_sparse_build_array(const rz_context* ctx, uint64_t n_bands, const uint64_t* counts, const uint64_t* rows,
                          const uint64_t* cols, const void* data, void* out, rz_stats* stats, char* err, size_t errlen) {
    return guarded(err, errlen, [&]() { rz::sparse_build_array(ctx, n_bands, counts, rows, cols, data, out, stats); });
}

void* rz_host_alloc(size_t bytes, char* err, size_t errlen) {
    void* p = nullptr;
    const int rc = guarded(err, errlen, [&]() {
        if (bytes == 0) throw Error{RZ_VALUE_ERROR, "rz_host_alloc: zero bytes"};
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
            (void)cudaGetLastError();
            throw Error{RZ_RUNTIME_ERROR, "No CUDA device available: librz_b200 has no CPU fallback."};
        }
        p = rz::g_host_pool.lease(std::max<size_t>(bytes, (size_t)2 << 20), true);
    });
    return rc == RZ_OK ? p : nullptr;
}

void rz_host_free(void* p) {
    if (p) rz::g_host_pool.unlease(p);
}

uint64_t rz_host_trim(uint64_t keep_bytes) {
    rz::DeviceGuard guard;
    rz::g_host_pool.trim((size_t)keep_bytes);
    return (uint64_t)rz::g_host_pool.pooled();
}

int rz_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

const char* rz_version(void) { return "rz_b200 0.2.0 (sm_100a)"; }

// sizes and selected field offsets of the header's structs, for bindings that mirror them by hand
int rz_abi_layout(uint64_t* out, int n) {
    const uint64_t v[16] = {
        sizeof(rz_raster_info), sizeof(rz_raw_raster_info), sizeof(rz_geom_soa), sizeof(rz_context), sizeof(rz_stats),
        offsetof(rz_context, field), offsetof(rz_context, band_of_geom), offsetof(rz_context, background),
        offsetof(rz_context, row_begin), offsetof(rz_context, stream), offsetof(rz_context, flags),
        offsetof(rz_stats, h2d_ms), offsetof(rz_stats, h2d_bytes), offsetof(rz_stats, kernel_launches),
        offsetof(rz_stats, n_mask_words), offsetof(rz_stats, wall_ms)};
    for (int i = 0; i < n && i < 16; i++) out[i] = v[i];
    return 16;
}

}  // extern "C"
