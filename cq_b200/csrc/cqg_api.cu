// cqg_api.cu — host side of libcqgpu: the C-ABI of include/cq_gpu.h over the sm_100a kernels
// of cqg_scan.cuh. Plain CUDA runtime; no torch, no CPU operator fallback: every query either
// runs on the device or fails with an error code.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <strings.h>

#include <algorithm>
#include <cerrno>
#include <cstdarg>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cq_gpu.h"
#include "cqg_lean.cuh"
#include "cqg_lean2.cuh"
#include "cqg_lean2g.cuh"
#include "cqg_lean2k.cuh"
#include "cqg_leanhc.cuh"

using namespace cqg;

#define CQG_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024];
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(CQG_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

CQG_API const char* cqg_last_error(void) { return g_err; }
CQG_API int cqg_abi_version(void) { return 1; }
CQG_API int64_t cqg_total_kernel_launches(void) { return g_launches.load(); }
// launches per scan kernel family (tests assert that the kernel a plan is meant for is the one that ran)
enum { KF_SCAN = 0, KF_LEAN, KF_LEAN2, KF_LEAN2G, KF_LEAN2K, KF_LEANHC, KF_COUNT };
static std::atomic<int64_t> g_family[KF_COUNT];
static const char* const kFamilyName[KF_COUNT] = {"scan", "lean", "lean2", "lean2g", "lean2k", "leanhc"};
// the last lean scan of this thread: tiles it covered, tiles and rows it handed to the general kernel
static thread_local int64_t g_last_scan[3];
CQG_API void cqg_last_scan_stats(int64_t* tiles, int64_t* handed_tiles, int64_t* handed_rows) {
    if (tiles) *tiles = g_last_scan[0];
    if (handed_tiles) *handed_tiles = g_last_scan[1];
    if (handed_rows) *handed_rows = g_last_scan[2];
}
CQG_API int64_t cqg_kernel_launches_named(const char* family) {
    for (int k = 0; k < KF_COUNT; k++)
        if (family && !strcmp(family, kFamilyName[k])) return g_family[k].load();
    return -1;
}

CQG_API int cqg_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(CQG_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return -1;
    }
    return n;
}

static bool g_pool_ready[64];
CQG_API int cqg_set_device(int device) {
    CU(cudaSetDevice(device));
    if (device >= 0 && device < 64 && !g_pool_ready[device]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        g_pool_ready[device] = true;
    }
    return CQG_OK;
}

static int ensure_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(CQG_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64 && !g_pool_ready[dev]) return cqg_set_device(dev);
    return CQG_OK;
}

constexpr size_t kDevPad = 4096;
CQG_API size_t cqg_device_padding(void) { return kDevPad; }

// ------------------------------------------------------------------------------------------
// device scratch with RAII
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t n = 0;
    cudaStream_t s = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t bytes, cudaStream_t st) {
        release();
        s = st;
        n = bytes ? bytes : 8;
        return cudaMallocAsync(&p, n, st);
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
struct cqg_table {
    uint8_t* d_data = nullptr;
    bool owns_device = false;
    size_t size = 0;
    const uint8_t* h_data = nullptr;  // host view of the bytes when there is one
    void* map = nullptr;              // mmap'ed region (portable_mmap, src/mmap.c:78)
    size_t map_len = 0;
    cqg_csv_config_t cfg{};
    std::vector<std::string> names;
    size_t data_start = 0;  // where data rows may begin (end of the header line)
    int shard_index = 0, shard_count = 1;
    uint64_t global_base = 0;
    int64_t row_count = -1;
    mutable std::vector<uint8_t> sample;  // first bytes after the header (layout guesses only), fetched on first use
    mutable bool sample_ready = false;
    bool src_pinned = false;  // h_data is page-locked: DMA straight from it
    int fd = -1;              // the mapped file, kept open: staging reads it with pread (no page faults on the mapping)
    // explicit ownership range of a scan (overrides the equal shards): the bytes one GPU of a multi-GPU table holds
    uint64_t range_lo = 0, range_hi = 0;
    bool has_range = false;
    // multi-GPU residency (CQ_GPUS, cqg_multi below): one virtual address range, slice d physically on device d
    int ngpu = 0;
    uint64_t va_size = 0;
    std::vector<unsigned long long> vmm_handles;
    std::vector<uint64_t> cuts;  // ngpu + 1 byte offsets: device d holds [cuts[d], cuts[d + 1])
    std::vector<int> devices;    // CUDA device of slice d
    std::vector<cqg_table*> views;  // per slice: the same bytes (same address), scans own [cuts[d], cuts[d + 1])
};

static inline bool host_is_space(unsigned c) { return c == 32u || (c - 9u) <= 4u; }

// parse_line (src/csv_reader.c:285-338) on the host, for the header line only
static void split_header(const uint8_t* ls, const uint8_t* le, char delim, char quote,
                         std::vector<std::pair<const uint8_t*, size_t>>& out) {
    const uint8_t* ptr = ls;
    while (ptr < le) {
        while (ptr < le && host_is_space(*ptr)) ptr++;
        if (ptr >= le) break;
        const uint8_t* fs = ptr;
        size_t flen = 0;
        if (*ptr == (uint8_t)quote) {
            ptr++;
            fs = ptr;
            while (ptr < le) {
                if (*ptr == (uint8_t)quote) {
                    if (ptr + 1 < le && ptr[1] == (uint8_t)quote) {
                        ptr += 2;
                        flen += 2;
                    } else {
                        flen = (size_t)(ptr - fs);
                        ptr++;
                        break;
                    }
                } else {
                    ptr++;
                }
            }
            while (ptr < le && *ptr != (uint8_t)delim) ptr++;
        } else {
            while (ptr < le && *ptr != (uint8_t)delim) ptr++;
            flen = (size_t)(ptr - fs);
        }
        out.emplace_back(fs, flen);
        if (ptr < le && *ptr == (uint8_t)delim) ptr++;
    }
}

// header handling of csv_load (src/csv_reader.c:341-357, 404-427) given the first bytes of the file
static bool parse_header_bytes(cqg_table* t, const uint8_t* p, size_t n, bool complete) {
    size_t i = 0;
    while (i < n) {
        size_t ls = i;
        while (i < n && p[i] != '\n' && p[i] != '\r') i++;
        size_t le = i;
        if (le > ls) {
            if (i == n && !complete) return false;  // line not finished in this prefix
            std::vector<std::pair<const uint8_t*, size_t>> f;
            split_header(p + ls, p + le, t->cfg.delimiter, t->cfg.quote, f);
            t->names.clear();
            for (size_t c = 0; c < f.size(); c++) {
                if (t->cfg.has_header && f[c].second > 0) {
                    std::string s((const char*)f[c].first, f[c].second);
                    size_t z = s.find('\0');
                    if (z != std::string::npos) s.resize(z);
                    size_t a = 0;
                    while (a < s.size() && host_is_space((unsigned char)s[a])) a++;
                    s = s.substr(a);
                    while (s.size() > 1 && host_is_space((unsigned char)s.back())) s.pop_back();
                    if (s.size() == 1 && host_is_space((unsigned char)s[0]) && false) s.clear();
                    t->names.push_back(s);
                } else {
                    t->names.push_back("$" + std::to_string(c));
                }
            }
            t->data_start = t->cfg.has_header ? le : ls;
            return true;
        }
        while (i < n && (p[i] == '\n' || p[i] == '\r')) i++;
    }
    if (!complete) return false;
    t->names.clear();
    t->data_start = n;
    return true;
}

static int finish_open(cqg_table* t) {
    if (t->h_data) {
        parse_header_bytes(t, t->h_data, t->size, true);
        return CQG_OK;
    }
    // device-only bytes: pull prefixes until the header line is complete
    size_t want = 1 << 16;
    std::vector<uint8_t> tmp;
    for (;;) {
        size_t n = std::min(want, t->size);
        tmp.resize(n);
        CU(cudaMemcpy(tmp.data(), t->d_data, n, cudaMemcpyDeviceToHost));
        if (parse_header_bytes(t, tmp.data(), n, n == t->size)) return CQG_OK;
        want *= 16;
    }
}

static int check_dialect(cqg_csv_config_t cfg) {
    unsigned d = (unsigned char)cfg.delimiter, q = (unsigned char)cfg.quote;
    auto bad = [](unsigned c) {
        return c == 0 || c == '\n' || c == '\r' || (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') ||
               c == '.' || c == '+' || c == '-';
    };
    // strtod / strtoll in the reference read on past the field into a delimiter that looks
    // numeric (src/csv_reader.c:207-210); such dialects are not reproduced here
    if (bad(d) || bad(q)) return fail(CQG_ERR_UNSUPPORTED_PLAN, "delimiter/quote character 0x%02x/0x%02x not supported", d, q);
    return CQG_OK;
}

// Host bytes -> HBM. Page-locked sources are DMA'd in place. Pageable / mmap'ed sources (portable_mmap,
// src/mmap.c:78-108) go through page-locked bounce buffers: several host threads, each with two buffers and a
// stream of its own, fault their chunks in and hand them to the copy engines, so that the host memcpy of one
// chunk overlaps the DMA of others (one thread tops out far below what PCIe Gen5 takes).
static int env_int_early(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

struct StagePool {
    static constexpr int kThreads = 32;  // at most; copy_host_range uses min(kThreads, CQG_STAGE_THREADS (24), cores) of them
    static constexpr size_t kChunk = 8u << 20;
    uint8_t* bounce[kThreads][2] = {};
    cudaStream_t stream[kThreads] = {};
    cudaEvent_t ev[kThreads][2] = {};
    int ready = 0;  // threads whose buffers exist
    std::mutex mu;
    int prepare(int want) {
        for (; ready < want; ready++) {
            for (int k = 0; k < 2; k++)
                if (cudaMallocHost((void**)&bounce[ready][k], kChunk) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev[ready][k], cudaEventDisableTiming) != cudaSuccess) {
                    cudaGetLastError();
                    return ready;
                }
            if (cudaStreamCreateWithFlags(&stream[ready], cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                return ready;
            }
        }
        return ready;
    }
};
static StagePool g_stage[64];

// host bytes [off, off + n) of the table -> device memory at dst, on the CURRENT device (its own bounce buffers and
// streams): several threads, each with two buffers
static int copy_host_range(const cqg_table* t, const uint8_t* src, uint64_t off0, size_t n_total, uint8_t* dst) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    StagePool& sp = g_stage[dev & 63];
    std::lock_guard<std::mutex> lock(sp.mu);
    constexpr size_t chunk = StagePool::kChunk;
    const size_t nchunks = (n_total + chunk - 1) / chunk;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int cap_threads = std::max(1, std::min(StagePool::kThreads, env_int_early("CQG_STAGE_THREADS", 24)));
    int want = (int)std::min<size_t>(std::min<size_t>((size_t)cap_threads, hw), std::max<size_t>(1, nchunks / 2));
    const int T = sp.prepare(want);
    if (T < 1) return fail(CQG_ERR_CUDA, "page-locked staging buffers: allocation failed");
    std::atomic<int> bad{0};
    auto work = [&](int k) {
        cudaSetDevice(dev);
        int b = 0;
        for (size_t c = (size_t)k; c < nchunks && !bad.load(); c += (size_t)T, b ^= 1) {
            const size_t off = c * chunk, n = std::min(chunk, n_total - off);
            if (cudaEventSynchronize(sp.ev[k][b]) != cudaSuccess) bad = 1;  // the copy that last used this buffer
            bool have = false;
            if (t->fd >= 0) {  // page cache -> bounce buffer in the kernel: no per-page faults on the mapping
                size_t got = 0;
                while (got < n) {
                    const ssize_t r = pread(t->fd, sp.bounce[k][b] + got, n - got, (off_t)(off0 + off + got));
                    if (r <= 0) break;
                    got += (size_t)r;
                }
                have = got == n;
            }
            if (!have) memcpy(sp.bounce[k][b], src + off0 + off, n);
            if (cudaMemcpyAsync(dst + off, sp.bounce[k][b], n, cudaMemcpyHostToDevice, sp.stream[k]) != cudaSuccess) bad = 1;
            cudaEventRecord(sp.ev[k][b], sp.stream[k]);
        }
        if (cudaStreamSynchronize(sp.stream[k]) != cudaSuccess) bad = 1;
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < T; k++) th.emplace_back(work, k);
        for (auto& x : th) x.join();
    }
    if (bad.load()) return fail(CQG_ERR_CUDA, "host to device staging: %s", cudaGetErrorString(cudaGetLastError()));
    return CQG_OK;
}

static int stage_to_device(cqg_table* t, const uint8_t* src, size_t size, bool pinned) {
    // stream-ordered allocation from the device pool (kept warm: release threshold is unlimited)
    CU(cudaMallocAsync((void**)&t->d_data, size + kDevPad, 0));
    t->owns_device = true;
    CU(cudaMemsetAsync(t->d_data + size, '\n', kDevPad, 0));
    if (pinned) {
        CU(cudaMemcpyAsync(t->d_data, src, size, cudaMemcpyHostToDevice, 0));
        CU(cudaStreamSynchronize(0));
        return CQG_OK;
    }
    CU(cudaStreamSynchronize(0));  // the allocation is ordered on stream 0, the copies run on streams of their own
    return copy_host_range(t, src, 0, size, t->d_data);
}

// ------------------------------------------------------------------------------------------
// multi-GPU tables (CQ_GPUS=N, cqg_table_open_multi): ONE virtual address range for the file, its 2 MB pages
// physically on N devices (slice d = bytes [cuts[d], cuts[d + 1]) on device d), mapped readable and writable on
// all of them (CUDA virtual memory management). Every device scans the rows that start in its own slice out of
// its own HBM; a row reference (global offset) means the same on every device, so the merged result is finished
// by device 0 reading the groups' representative rows over NVLink, and a query shape that does not shard (joins,
// projections) simply runs on device 0 over the whole range. The driver entry points are looked up through the
// runtime (cudaGetDriverEntryPoint): the library does not link libcuda.
// ------------------------------------------------------------------------------------------
namespace vmm {
typedef CUresult (*GetGranularity_t)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*AddressReserve_t)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*AddressFree_t)(CUdeviceptr, size_t);
typedef CUresult (*Create_t)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*Release_t)(CUmemGenericAllocationHandle);
typedef CUresult (*Map_t)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*Unmap_t)(CUdeviceptr, size_t);
typedef CUresult (*SetAccess_t)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
struct Api {
    GetGranularity_t granularity = nullptr;
    AddressReserve_t reserve = nullptr;
    AddressFree_t addr_free = nullptr;
    Create_t create = nullptr;
    Release_t release = nullptr;
    Map_t map = nullptr;
    Unmap_t unmap = nullptr;
    SetAccess_t set_access = nullptr;
    bool ok = false;
};
static Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name) -> void* {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            return fn;
        };
        a.granularity = (GetGranularity_t)get("cuMemGetAllocationGranularity");
        a.reserve = (AddressReserve_t)get("cuMemAddressReserve");
        a.addr_free = (AddressFree_t)get("cuMemAddressFree");
        a.create = (Create_t)get("cuMemCreate");
        a.release = (Release_t)get("cuMemRelease");
        a.map = (Map_t)get("cuMemMap");
        a.unmap = (Unmap_t)get("cuMemUnmap");
        a.set_access = (SetAccess_t)get("cuMemSetAccess");
        a.ok = a.granularity && a.reserve && a.addr_free && a.create && a.release && a.map && a.unmap && a.set_access;
    });
    return a;
}
}  // namespace vmm

// One released multi-GPU allocation is kept for the next table of the same size and devices (creating and mapping 10 GB
// of physical memory costs 0.1-0.2 s; the single-GPU route keeps its stream-ordered pool warm in the same way).
struct MultiAlloc {
    uint8_t* va = nullptr;
    uint64_t total = 0;
    std::vector<int> devices;
    std::vector<uint64_t> cuts;
    std::vector<unsigned long long> handles;
};
static std::mutex g_multi_mu;
static MultiAlloc g_multi_spare;

static void release_multi(cqg_table* t, bool keep = true) {
    vmm::Api& v = vmm::api();
    for (cqg_table* w : t->views) delete w;
    t->views.clear();
    if (!t->va_size || !v.ok) return;
    int home = 0;
    cudaGetDevice(&home);
    for (int d : t->devices) {
        cudaSetDevice(d);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(home);
    MultiAlloc old;
    if (keep) {
        std::lock_guard<std::mutex> lock(g_multi_mu);
        old = g_multi_spare;
        g_multi_spare.va = t->d_data;
        g_multi_spare.total = t->va_size;
        g_multi_spare.devices = t->devices;
        g_multi_spare.cuts = t->cuts;
        g_multi_spare.handles = t->vmm_handles;
    } else {  // (a range that was never completely mapped: taken apart, not kept)
        old.va = t->d_data;
        old.total = t->va_size;
        old.handles = t->vmm_handles;
    }
    t->vmm_handles.clear();
    t->d_data = nullptr;
    t->va_size = 0;
    if (!old.va) return;
    v.unmap((CUdeviceptr)(uintptr_t)old.va, old.total);
    for (unsigned long long h : old.handles) v.release((CUmemGenericAllocationHandle)h);
    v.addr_free((CUdeviceptr)(uintptr_t)old.va, old.total);
}

// the host bytes of a multi-GPU table onto its devices: slices cut at 2 MB pages, one staging thread (with its own
// bounce buffers and copy streams, copy_host_range) per device, so that N host-to-device links run at once
static int stage_multi(cqg_table* t) {
    vmm::Api& v = vmm::api();
    if (!v.ok) return fail(CQG_ERR_CUDA, "multi-GPU tables need the CUDA virtual memory management entry points");
    const int N = t->ngpu;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    const bool same = getenv("CQG_MULTI_SAME_DEVICE") != nullptr;  // (tests on a one-GPU box: every slice on device 0)
    if (ndev < 1 || (!same && ndev < N)) return fail(CQG_ERR_CUDA, "CQ_GPUS=%d but %d CUDA device(s)", N, ndev);
    int home = 0;
    CU(cudaGetDevice(&home));
    t->devices.resize(N);
    for (int d = 0; d < N; d++) t->devices[d] = same ? home : d;
    for (int d : t->devices) {  // (a context on every device before any driver call names it)
        CU(cudaSetDevice(d));
        CU(cudaFree(nullptr));
        int rc = ensure_device();
        if (rc) return rc;
    }
    CU(cudaSetDevice(t->devices[0]));
    size_t gran = 2u << 20;
    for (int d = 0; d < N; d++) {
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = t->devices[d];
        size_t g = 0;
        if (v.granularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return fail(CQG_ERR_CUDA, "cuMemGetAllocationGranularity failed");
        gran = std::max(gran, g);
    }
    const bool timing = getenv("CQG_TIMING") != nullptr;
    auto now_ms = [] {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    double t_mark = now_ms();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const double t1 = now_ms();
        fprintf(stderr, "[cqg timing] %-28s %9.3f ms\n", what, t1 - t_mark);
        t_mark = t1;
    };
    const uint64_t total = ((uint64_t)t->size + kDevPad + gran - 1) / gran * gran;
    const uint64_t pages = total / gran;
    t->cuts.assign(N + 1, 0);
    for (int d = 0; d <= N; d++) t->cuts[d] = pages * (uint64_t)d / (uint64_t)N * gran;
    CUdeviceptr va = 0;
    bool reused = false;
    {
        std::lock_guard<std::mutex> lock(g_multi_mu);
        if (g_multi_spare.va && g_multi_spare.total == total && g_multi_spare.devices == t->devices && g_multi_spare.cuts == t->cuts) {
            va = (CUdeviceptr)(uintptr_t)g_multi_spare.va;
            t->vmm_handles = g_multi_spare.handles;
            g_multi_spare = MultiAlloc();
            reused = true;
        }
    }
    if (!reused && v.reserve(&va, total, gran, 0, 0) != CUDA_SUCCESS)
        return fail(CQG_ERR_NOMEM, "cuMemAddressReserve of %llu bytes failed", (unsigned long long)total);
    t->d_data = (uint8_t*)(uintptr_t)va;
    t->va_size = total;
    for (int d = 0; d < N && !reused; d++) {
        const uint64_t len = t->cuts[d + 1] - t->cuts[d];
        if (!len) continue;
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = t->devices[d];
        CUmemGenericAllocationHandle h = 0;
        if (v.create(&h, len, &prop, 0) != CUDA_SUCCESS) {
            release_multi(t, false);
            return fail(CQG_ERR_NOMEM, "cuMemCreate of %llu bytes on device %d failed", (unsigned long long)len, t->devices[d]);
        }
        t->vmm_handles.push_back((unsigned long long)h);
        if (v.map(va + t->cuts[d], len, 0, h, 0) != CUDA_SUCCESS) {
            release_multi(t, false);
            return fail(CQG_ERR_CUDA, "cuMemMap failed");
        }
    }
    if (!reused) {
        std::vector<CUmemAccessDesc> acc;
        for (int d = 0; d < N; d++) {
            bool seen = false;
            for (const CUmemAccessDesc& a : acc) seen = seen || a.location.id == t->devices[d];
            if (seen) continue;
            CUmemAccessDesc a{};
            a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
            a.location.id = t->devices[d];
            a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            acc.push_back(a);
        }
        if (v.set_access(va, total, acc.data(), acc.size()) != CUDA_SUCCESS) {
            release_multi(t, false);
            return fail(CQG_ERR_CUDA, "cuMemSetAccess failed (no peer access between the devices of CQ_GPUS?)");
        }
    }
    lap("multi: reserve, create, map");
    // the bytes: every device pulls its own slice
    std::vector<int> rcs(N, CQG_OK);
    std::vector<std::string> errs(N);
    std::vector<std::thread> th;
    for (int d = 0; d < N; d++) {
        th.emplace_back([&, d] {
            if (cudaSetDevice(t->devices[d]) != cudaSuccess) {
                rcs[d] = CQG_ERR_CUDA;
                return;
            }
            const uint64_t lo = std::min<uint64_t>(t->cuts[d], t->size), hi = std::min<uint64_t>(t->cuts[d + 1], t->size);
            int rc = CQG_OK;
            if (hi > lo) {
                if (t->src_pinned) {
                    if (cudaMemcpyAsync(t->d_data + lo, t->h_data + lo, hi - lo, cudaMemcpyHostToDevice, 0) != cudaSuccess) rc = CQG_ERR_CUDA;
                } else {
                    rc = copy_host_range(t, t->h_data, lo, hi - lo, t->d_data + lo);
                }
            }
            if (rc == CQG_OK && t->cuts[d + 1] > t->size) {  // what lies behind the file on this device: newlines
                const uint64_t from = std::max<uint64_t>(t->cuts[d], t->size);
                if (cudaMemsetAsync(t->d_data + from, '\n', std::min<uint64_t>(t->cuts[d + 1], t->size + kDevPad) - from, 0) != cudaSuccess) rc = CQG_ERR_CUDA;
            }
            if (rc == CQG_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = CQG_ERR_CUDA;
            if (rc != CQG_OK) errs[d] = cqg_last_error();
            rcs[d] = rc;
        });
    }
    for (std::thread& x : th) x.join();
    lap("multi: slices staged");
    CU(cudaSetDevice(home));
    for (int d = 0; d < N; d++)
        if (rcs[d] != CQG_OK) {
            release_multi(t, false);
            return fail(rcs[d], "staging slice %d: %s", d, errs[d].empty() ? "CUDA error" : errs[d].c_str());
        }
    // per slice a view of the same address range that owns the rows starting in the slice
    for (int d = 0; d < N; d++) {
        cqg_table* w = new cqg_table();
        w->d_data = t->d_data;
        w->size = t->size;
        w->cfg = t->cfg;
        w->names = t->names;
        w->data_start = t->data_start;
        w->global_base = t->global_base;
        w->range_lo = std::min<uint64_t>(t->cuts[d], t->size);
        w->range_hi = d + 1 == N ? (uint64_t)t->size : std::min<uint64_t>(t->cuts[d + 1], t->size);
        w->has_range = true;
        t->views.push_back(w);
    }
    return CQG_OK;
}

// tables opened from host bytes are uploaded when a query first needs them: a statement whose shape is then
// declined (plan-time checks) or routed elsewhere never pays for the copy
static int ensure_staged(const cqg_table* tc) {
    cqg_table* t = const_cast<cqg_table*>(tc);
    if (!t || t->d_data || !t->h_data) return CQG_OK;
    if (t->ngpu > 1) return stage_multi(t);
    return stage_to_device(t, t->h_data, t->size, t->src_pinned);
}

CQG_API int cqg_table_open_buffer(const void* data, size_t size, int pinned, cqg_csv_config_t cfg, cqg_table_t** out) {
    if (!out) return fail(CQG_ERR_ARG, "null out");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = check_dialect(cfg))) return rc;
    cqg_table* t = new cqg_table();
    t->cfg = cfg;
    t->size = size;
    t->h_data = (const uint8_t*)data;
    t->src_pinned = pinned != 0;
    rc = finish_open(t);
    if (rc != CQG_OK) {
        cqg_table_close(t);
        return rc;
    }
    *out = t;
    return CQG_OK;
}

CQG_API int cqg_table_open(const char* path, cqg_csv_config_t cfg, cqg_table_t** out) {
    if (!out || !path) return fail(CQG_ERR_ARG, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = check_dialect(cfg))) return rc;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(CQG_ERR_IO, "%s", strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) < 0 || sb.st_size == 0) {  // src/mmap.c:91-98: an empty file is an error too
        int e = errno;
        close(fd);
        return fail(CQG_ERR_IO, "%s", sb.st_size == 0 ? "empty file" : strerror(e));
    }
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
        int e = errno;
        close(fd);
        return fail(CQG_ERR_IO, "mmap: %s", strerror(e));
    }
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    cqg_table* t = new cqg_table();
    t->fd = fd;
    t->cfg = cfg;
    t->size = (size_t)sb.st_size;
    t->map = m;
    t->map_len = (size_t)sb.st_size;
    t->h_data = (const uint8_t*)m;
    rc = finish_open(t);
    if (rc != CQG_OK) {
        cqg_table_close(t);
        return rc;
    }
    if (const char* e = getenv("CQ_GPUS")) {  // the drop-in CLI: CQ_GPUS=8 cq -q "SELECT ... FROM 'big.csv' ..."
        const int n = atoi(e);
        if (n > 1 && (rc = cqg_table_set_gpus(t, n)) != CQG_OK) {
            cqg_table_close(t);
            return rc;
        }
    }
    *out = t;
    return CQG_OK;
}

// Spread a table that still lives in host memory (opened from a path or a host buffer, no query run yet) over the first
// `ngpu` devices. Files below CQG_MULTI_MIN_BYTES (default 64 MB) stay on one device: there is nothing to win.
CQG_API int cqg_table_set_gpus(cqg_table_t* t, int ngpu) {
    if (!t || ngpu < 1 || ngpu > 64) return fail(CQG_ERR_ARG, "bad argument");
    if (t->d_data || !t->h_data) return fail(CQG_ERR_ARG, "the table is already resident on a device");
    const char* e = getenv("CQG_MULTI_MIN_BYTES");
    const uint64_t min_bytes = e ? strtoull(e, nullptr, 10) : (64ull << 20);
    if (!getenv("CQG_MULTI_SAME_DEVICE")) {  // no more slices than devices (CQ_GPUS=8 on a smaller box: what is there)
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 1;
        ngpu = std::min(ngpu, std::max(ndev, 1));
    }
    t->ngpu = (ngpu > 1 && t->size >= min_bytes) ? ngpu : 0;
    return CQG_OK;
}

CQG_API int cqg_table_gpus(const cqg_table_t* t) { return t ? (t->ngpu > 1 ? t->ngpu : 1) : 0; }

CQG_API int cqg_table_open_device(uint64_t device_ptr, size_t size, cqg_csv_config_t cfg, cqg_table_t** out) {
    if (!out || !device_ptr) return fail(CQG_ERR_ARG, "null argument");
    if (device_ptr & 15ull) return fail(CQG_ERR_ARG, "device pointer must be 16-byte aligned");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = check_dialect(cfg))) return rc;
    cqg_table* t = new cqg_table();
    t->cfg = cfg;
    t->size = size;
    t->d_data = (uint8_t*)device_ptr;
    t->owns_device = false;
    rc = finish_open(t);
    if (rc != CQG_OK) {
        cqg_table_close(t);
        return rc;
    }
    *out = t;
    return CQG_OK;
}

CQG_API int cqg_table_set_shard(cqg_table_t* t, int index, int count) {
    if (!t || count < 1 || index < 0 || index >= count) return fail(CQG_ERR_ARG, "bad shard %d/%d", index, count);
    t->shard_index = index;
    t->shard_count = count;
    t->row_count = -1;
    return CQG_OK;
}

CQG_API int cqg_table_set_global_offset(cqg_table_t* t, uint64_t offset) {
    if (!t) return fail(CQG_ERR_ARG, "null table");
    t->global_base = offset;
    return CQG_OK;
}

CQG_API void cqg_table_close(cqg_table_t* t) {
    if (!t) return;
    if (t->ngpu > 1) release_multi(t);
    else if (t->owns_device && t->d_data) cudaFreeAsync(t->d_data, 0);
    if (t->map) munmap(t->map, t->map_len);
    if (t->fd >= 0) close(t->fd);
    delete t;
}

CQG_API int cqg_table_column_count(const cqg_table_t* t) { return t ? (int)t->names.size() : 0; }
CQG_API const char* cqg_table_column_name(const cqg_table_t* t, int col) {
    return (t && col >= 0 && col < (int)t->names.size()) ? t->names[col].c_str() : nullptr;
}
CQG_API int cqg_table_column_index(const cqg_table_t* t, const char* name) {
    if (!t || !name) return -1;
    for (size_t i = 0; i < t->names.size(); i++)
        if (strcasecmp(t->names[i].c_str(), name) == 0) return (int)i;
    return -1;
}
CQG_API size_t cqg_table_size(const cqg_table_t* t) { return t ? t->size : 0; }
CQG_API uint64_t cqg_table_device_ptr(const cqg_table_t* t) {
    if (!t || ensure_staged(t) != CQG_OK) return 0;
    return (uint64_t)t->d_data;
}

// ------------------------------------------------------------------------------------------
// kernel configuration
// ------------------------------------------------------------------------------------------
using ScanGeo = Geo<128, 16384, 2>;

struct LaunchCfg {
    int sms = 0;
    int ctas_per_sm = 0;
    bool ready = false;
};
static LaunchCfg g_cfg[64];

static int scan_smem_bytes(int table_bytes) { return ScanGeo::OFF_TABLE + table_bytes; }

template <int NW>
static int launch_scan_nw(const DevPlan& P, int smem, int dev, cudaStream_t st) {
    static bool attr_set[64];
    if (!attr_set[dev & 63]) {
        CU(cudaFuncSetAttribute(scan_kernel<ScanGeo, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev & 63] = true;
    }
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_kernel<ScanGeo, NW>, ScanGeo::THREADS, smem));
    if (per_sm < 1) return fail(CQG_ERR_CUDA, "scan kernel does not fit: %d bytes of shared memory", smem);
    int grid = std::min(P.n_tiles, c.sms * per_sm);
    scan_kernel<ScanGeo, NW><<<grid, ScanGeo::THREADS, smem, st>>>(P);
    g_launches++;
    g_family[KF_SCAN]++;
    CU(cudaGetLastError());
    return CQG_OK;
}

// ------------------------------------------------------------------------------------------
// kernels compiled per query (NVRTC): the lean kernels with the plan's shape as compile-time constants
// ------------------------------------------------------------------------------------------
// The ahead-of-time lean kernels interpret the plan's shape per row (which columns, how many delimiters between
// them, which leaf reads which slot, ...). cqg_jit compiles the SAME source once per distinct shape with the
// shape as macros (cqg_lean2.cuh: CQG_SPEC), caches the cubin in the process, and launches it through the
// runtime's library API. Anything that goes wrong (no libnvrtc, sources not next to the library, a compile
// error) falls back to the ahead-of-time kernel of the same source: same results, only slower. CQG_JIT=0 turns
// it off.
namespace cqg_jit {

typedef int (*nvrtcCreateProgram_t)(void**, const char*, const char*, int, const char* const*, const char* const*);
typedef int (*nvrtcProg1_t)(void**);
typedef int (*nvrtcAddName_t)(void*, const char*);
typedef int (*nvrtcCompile_t)(void*, int, const char* const*);
typedef int (*nvrtcLowered_t)(void*, const char*, const char**);
typedef int (*nvrtcSize_t)(void*, size_t*);
typedef int (*nvrtcGet_t)(void*, char*);

struct Api {
    void* h = nullptr;
    nvrtcCreateProgram_t create = nullptr;
    nvrtcProg1_t destroy = nullptr;
    nvrtcAddName_t add_name = nullptr;
    nvrtcCompile_t compile = nullptr;
    nvrtcLowered_t lowered = nullptr;
    nvrtcSize_t cubin_size = nullptr, log_size = nullptr;
    nvrtcGet_t cubin = nullptr, log = nullptr;
    bool ok = false;
};

static Api& api() {
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    // the newest run-time compiler of those present: a process that loaded another CUDA's libnvrtc.so.12 first (PyTorch
    // bundles its own) would otherwise compile with that one, and 256-bit loads (pk_load, cqg_lean.cuh) need 12.9
    int best = -1;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
        void* h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (!h) continue;
        int major = 0, minor = 0;
        typedef int (*nvrtcVersion_t)(int*, int*);
        nvrtcVersion_t ver = (nvrtcVersion_t)dlsym(h, "nvrtcVersion");
        if (ver) ver(&major, &minor);
        if (major * 100 + minor > best) {
            best = major * 100 + minor;
            a.h = h;
        }
    }
    if (!a.h) return a;
    a.create = (nvrtcCreateProgram_t)dlsym(a.h, "nvrtcCreateProgram");
    a.destroy = (nvrtcProg1_t)dlsym(a.h, "nvrtcDestroyProgram");
    a.add_name = (nvrtcAddName_t)dlsym(a.h, "nvrtcAddNameExpression");
    a.compile = (nvrtcCompile_t)dlsym(a.h, "nvrtcCompileProgram");
    a.lowered = (nvrtcLowered_t)dlsym(a.h, "nvrtcGetLoweredName");
    a.cubin_size = (nvrtcSize_t)dlsym(a.h, "nvrtcGetCUBINSize");
    a.cubin = (nvrtcGet_t)dlsym(a.h, "nvrtcGetCUBIN");
    a.log_size = (nvrtcSize_t)dlsym(a.h, "nvrtcGetProgramLogSize");
    a.log = (nvrtcGet_t)dlsym(a.h, "nvrtcGetProgramLog");
    a.ok = a.create && a.destroy && a.add_name && a.compile && a.lowered && a.cubin_size && a.cubin && a.log_size && a.log;
    return a;
}

// <dir of libcqgpu.so>/csrc and <that dir>/../include (the in-tree layout); CQG_JIT_SRC / CQG_JIT_INCLUDE override
static void source_dirs(std::string& csrc, std::string& inc) {
    const char* e1 = getenv("CQG_JIT_SRC");
    const char* e2 = getenv("CQG_JIT_INCLUDE");
    std::string base;
    Dl_info info;
    if (dladdr((void*)&source_dirs, &info) && info.dli_fname) {
        base = info.dli_fname;
        size_t k = base.rfind('/');
        base = k == std::string::npos ? "." : base.substr(0, k);
    }
    csrc = e1 ? e1 : base + "/csrc";
    inc = e2 ? e2 : base + "/../include";
}

struct Kernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    bool failed = false;
};

static std::mutex g_mu;
static std::map<std::string, Kernel> g_cache;

static bool enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("CQG_JIT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

// a compile costs about a second: only scans big enough to win it back within a few queries get one
// (CQG_JIT_MIN_BYTES, default 64 MB of owned bytes; read every time so that tests can change it)
static bool worth_it(uint64_t scan_bytes) {
    const char* e = getenv("CQG_JIT_MIN_BYTES");
    const uint64_t min_bytes = e ? strtoull(e, nullptr, 10) : (64ull << 20);
    return scan_bytes >= min_bytes;
}

// a hash of everything NVRTC will read (the kernel headers next to the library and cq_gpu.h): part of every cache key,
// so that editing a header never reuses a cubin compiled from the old text
static uint64_t sources_hash() {
    static uint64_t h = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        std::string csrc, inc;
        source_dirs(csrc, inc);
        uint64_t x = 1469598103934665603ull;
        const char* files[] = {"cqg_rtc.h", "cqg_device.cuh", "cqg_plan.cuh", "cqg_scan.cuh", "cqg_lean.cuh", "cqg_lean2.cuh",
                               "cqg_lean2g.cuh", "cqg_lean2k.cuh", "cqg_leanhc.cuh"};
        auto eat = [&](const std::string& path) {
            if (FILE* f = fopen(path.c_str(), "rb")) {
                char buf[65536];
                size_t n;
                while ((n = fread(buf, 1, sizeof buf, f)) > 0)
                    for (size_t i = 0; i < n; i++) x = (x ^ (unsigned char)buf[i]) * 1099511628211ull;
                fclose(f);
            }
        };
        for (const char* f : files) eat(csrc + "/" + f);
        eat(inc + "/cq_gpu.h");
        h = x;
    });
    return h;
}

// cubins persist across processes in $CQG_JIT_CACHE, else $HOME/.cache/cqg_jit; nowhere when neither is set. The
// directory must belong to this user and be closed to everybody else (a cubin found there is loaded and launched).
// "" : no disk cache.
static std::string cache_path(const std::string& key) {
    const char* e = getenv("CQG_JIT_CACHE");
    std::string dir;
    if (e && *e) dir = e;
    else if (const char* home = getenv("HOME")) {
        if (!*home) return "";
        dir = std::string(home) + "/.cache/cqg_jit";
        mkdir((std::string(home) + "/.cache").c_str(), 0700);
    } else return "";
    mkdir(dir.c_str(), 0700);
    struct stat sb;
    if (lstat(dir.c_str(), &sb) != 0 || !S_ISDIR(sb.st_mode) || sb.st_uid != geteuid() || (sb.st_mode & 077) != 0) return "";
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : key) h = (h ^ c) * 1099511628211ull;
    char name[40];
    snprintf(name, sizeof name, "/%016llx.cubin", (unsigned long long)h);
    return dir + name;
}

// defs: the #define block of the shape; header: the file that holds the kernel template; name: its instantiation
static cudaKernel_t get(const std::string& defs, const char* header, const char* name, uint64_t scan_bytes) {
    if (!enabled() || !worth_it(scan_bytes)) return nullptr;
    std::lock_guard<std::mutex> lock(g_mu);
    char sh[40];
    snprintf(sh, sizeof sh, "sm_100a|v2|%016llx\n", (unsigned long long)sources_hash());
    const std::string key = std::string(sh) + name + "\n" + defs;
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second.failed ? nullptr : it->second.fn;
    Kernel& k = g_cache[key];
    k.failed = true;
    // a cubin of this shape from an earlier process? (file = the whole key, NUL, lowered name, NUL, cubin: the key is
    // compared, not just its hash in the file name)
    const std::string path = cache_path(key);
    if (!path.empty()) {
        std::vector<char> blob;
        if (FILE* f = fopen(path.c_str(), "rb")) {
            char buf[65536];
            size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) blob.insert(blob.end(), buf, buf + n);
            fclose(f);
        }
        if (blob.size() > key.size() + 2 && memcmp(blob.data(), key.data(), key.size()) == 0 && blob[key.size()] == 0) {
            const char* low0 = blob.data() + key.size() + 1;
            const void* nul = memchr(low0, 0, blob.size() - key.size() - 1);
            if (nul) {
                const size_t off = (const char*)nul - blob.data() + 1;
                if (off < blob.size() && cudaLibraryLoadData(&k.lib, blob.data() + off, nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
                    cudaLibraryGetKernel(&k.fn, k.lib, low0) == cudaSuccess) {
                    k.failed = false;
                    return k.fn;
                }
                cudaGetLastError();
            }
        }
    }
    Api& a = api();
    if (!a.ok) return nullptr;
    std::string csrc, inc;
    source_dirs(csrc, inc);
    const std::string src = "#define CQG_JIT 1\n" + defs + "#include \"" + header + "\"\n";
    void* prog = nullptr;
    if (a.create(&prog, src.c_str(), "cqg_jit.cu", 0, nullptr, nullptr) != 0) return nullptr;
    const std::string i1 = "-I" + csrc, i2 = "-I" + inc;
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-lineinfo", i1.c_str(), i2.c_str(),
                          "-I/usr/local/cuda/include"};
    bool ok = a.add_name(prog, name) == 0 && a.compile(prog, (int)(sizeof opts / sizeof opts[0]), opts) == 0;
    if (!ok && getenv("CQG_JIT_VERBOSE")) {
        size_t n = 0;
        a.log_size(prog, &n);
        std::string log(n + 1, 0);
        a.log(prog, &log[0]);
        fprintf(stderr, "[cqg jit] compile failed:\n%s\n", log.c_str());
    }
    const char* low = nullptr;
    size_t n = 0;
    std::vector<char> cubin;
    if (ok) ok = a.lowered(prog, name, &low) == 0 && low && a.cubin_size(prog, &n) == 0 && n > 0;
    std::string lowered = ok ? low : "";
    if (ok) {
        cubin.resize(n);
        ok = a.cubin(prog, cubin.data()) == 0;
    }
    a.destroy(&prog);
    if (!ok) return nullptr;
    if (cudaLibraryLoadData(&k.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
        cudaLibraryGetKernel(&k.fn, k.lib, lowered.c_str()) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    k.failed = false;
    if (getenv("CQG_JIT_VERBOSE")) fprintf(stderr, "[cqg jit] compiled %s (%zu bytes)\n", name, n);
    if (!path.empty()) {
        const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
        if (FILE* f = fopen(tmp.c_str(), "wb")) {
            const bool w = fwrite(key.c_str(), 1, key.size() + 1, f) == key.size() + 1 &&
                           fwrite(lowered.c_str(), 1, lowered.size() + 1, f) == lowered.size() + 1 && fwrite(cubin.data(), 1, n, f) == n;
            fclose(f);
            if (!w || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
        }
    }
    return k.fn;
}

// the shape of a lean plan as the macros cqg_lean2.cuh / cqg_lean2g.cuh read
static std::string lean_shape_defs(const cqg::DevPlan& P) {
    char b[256];
    std::string d;
    auto def = [&](const char* name, long long v) {
        snprintf(b, sizeof b, "#define CQG_JIT_%s %lld\n", name, v);
        d += b;
    };
    auto def_at = [&](const char* name, int n, auto value) {
        std::string m = std::string("#define CQG_JIT_") + name + "(i) (";
        for (int i = 0; i < n; i++) {
            snprintf(b, sizeof b, "(i)==%d?%lld:", i, (long long)value(i));
            m += b;
        }
        d += m + "0)\n";
    };
    def("NWANT", P.nwantL);
    def("GAP0", P.gap[0]);
    def("GAP1", P.gap[1]);
    def("GAP2", P.gap[2]);
    def("GAP3", P.gap[3]);
    def("NPROG", P.l_nprog);
    def("NLEAF", P.l_nleaf);
    def("NGC", P.ngc);
    def("CRLF", P.crlf ? 1 : 0);  // (lean2k_kernel: CR as a terminator is part of the compiled shape)
    def("NAGG", P.l_nagg);
    def_at("PROG", P.l_nprog, [&](int i) { return (int)P.l_prog[i]; });
    def_at("LEAFSLOT", P.l_nleaf, [&](int i) { return P.l_leaf[i].slot; });
    def_at("LEAFKIND", P.l_nleaf, [&](int i) { return P.l_leaf[i].kind; });
    def_at("GSLOT", P.ngc, [&](int i) { return (int)P.gslot[i]; });
    def_at("ASLOT", P.l_nagg, [&](int i) { return P.aggs[P.l_agg[i]].slot; });
    def_at("AFUNC", P.l_nagg, [&](int i) { return P.aggs[P.l_agg[i]].func; });
    // packed lines of the lean GROUP BY in global mode (PackedLayout)
    def("PKIDW", P.pk.id_words);
    def("PKBYTES", P.pk.entry_bytes);
    def("PKCOUNT", P.pk.count_off);
    def_at("PKKEYWORD", P.ngc, [&](int i) { return (int)P.pk.key_word[i]; });
    def_at("PKKEYWIDE", P.ngc, [&](int i) { return (int)P.pk.key_wide[i]; });
    def_at("PKAGGOFF", P.l_nagg, [&](int i) { return (int)P.pk.agg_off[i]; });
    def_at("PKAGGKEY", P.l_nagg, [&](int i) { return (int)P.pk.agg_key[i]; });
    return d;
}

}  // namespace cqg_jit

template <class LG, int MINB, bool GROUPED, bool ONELEAF, bool MINMAX>
static int launch_lean_geo(const DevPlan& P0, cudaStream_t st) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevPlan P = P0;
    // tiles of this geometry covering the same ownership range
    if (P.own_hi > P.own_lo) {
        P.first_tile = (int32_t)(P.own_lo / LG::TILE);
        P.n_tiles = (int32_t)((P.own_hi - 1) / LG::TILE) - P.first_tile + 1;
    }
    if (P.n_tiles <= 0) return CQG_OK;
    const int smem = LeanLayout<LG>::OFF_TABLE + (GROUPED ? kLeanDictCap * kLeanDictEntry + 16 + LG::NWARPS * lean_warp_acc(MINMAX) : 0);
    static bool attr_set[64];
    if (!attr_set[dev & 63]) {
        CU(cudaFuncSetAttribute(lean_kernel<LG, MINB, GROUPED, ONELEAF, MINMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev & 63] = true;
    }
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lean_kernel<LG, MINB, GROUPED, ONELEAF, MINMAX>, LG::THREADS, smem));
    if (per_sm < 1) return fail(CQG_ERR_CUDA, "lean kernel does not fit");
    int grid = std::min(P.n_tiles, c.sms * per_sm);
    {
        // compiled for this query's shape when the run-time compiler is there (cqg_jit), else the generic kernel below
        char name[200];
        snprintf(name, sizeof name, "cqg::lean_kernel<cqg::Geo<%d, %d, %d, %d>, %d, %s, %s, %s>", LG::THREADS, LG::TILE, LG::STAGES,
                 LG::OVER, MINB, GROUPED ? "true" : "false", ONELEAF ? "true" : "false", MINMAX ? "true" : "false");
        if (cudaKernel_t jk = cqg_jit::get(cqg_jit::lean_shape_defs(P), "cqg_lean.cuh", name, P.own_hi - P.own_lo)) {
            void* args[] = {(void*)&P};
            if (cudaFuncSetAttribute((const void*)jk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                cudaLaunchKernel((const void*)jk, dim3(grid), dim3(LG::THREADS), args, (size_t)smem, st) == cudaSuccess) {
                g_launches++;
    g_family[KF_LEAN]++;
                return CQG_OK;
            }
            cudaGetLastError();
        }
    }
    lean_kernel<LG, MINB, GROUPED, ONELEAF, MINMAX><<<grid, LG::THREADS, smem, st>>>(P);
    g_launches++;
    g_family[KF_LEAN]++;
    CU(cudaGetLastError());
    return CQG_OK;
}


template <class LG, int MINB, bool ONELEAF, int GAP0, bool CRLF = false>
static int launch_lean2_geo(const DevPlan& P0, cudaStream_t st) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevPlan P = P0;
    if (P.own_hi > P.own_lo) {
        P.first_tile = (int32_t)(P.own_lo / LG::TILE);
        P.n_tiles = (int32_t)((P.own_hi - 1) / LG::TILE) - P.first_tile + 1;
    }
    if (P.n_tiles <= 0) return CQG_OK;
    const int smem = Lean2Layout<LG>::TOTAL;  // tile + padded masks + interval table
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lean2_kernel<LG, MINB, ONELEAF, GAP0, CRLF>, LG::THREADS, smem));
    if (per_sm < 1) return fail(CQG_ERR_CUDA, "lean2 kernel does not fit");
    int grid = std::min(P.n_tiles, c.sms * per_sm);
    if (!ONELEAF) {
        // the general instantiation interprets the plan's shape per row: compiled for this query when possible
        char name[160];
        snprintf(name, sizeof name, "cqg::lean2_kernel<cqg::Geo<%d, %d, %d, %d>, %d, false, -1, %s>", LG::THREADS, LG::TILE, LG::STAGES,
                 LG::OVER, MINB, CRLF ? "true" : "false");
        if (cudaKernel_t jk = cqg_jit::get(cqg_jit::lean_shape_defs(P), "cqg_lean2.cuh", name, P.own_hi - P.own_lo)) {
            void* args[] = {(void*)&P};
            if (cudaLaunchKernel((const void*)jk, dim3(grid), dim3(LG::THREADS), args, (size_t)smem, st) == cudaSuccess) {
                g_launches++;
    g_family[KF_LEAN2]++;
                return CQG_OK;
            }
            cudaGetLastError();
        }
    }
    lean2_kernel<LG, MINB, ONELEAF, GAP0, CRLF><<<grid, LG::THREADS, smem, st>>>(P);
    g_launches++;
    g_family[KF_LEAN2]++;
    CU(cudaGetLastError());
    return CQG_OK;
}


template <class LG, int MINB>
static int launch_lean2g_geo(const DevPlan& P0, cudaStream_t st) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevPlan P = P0;
    if (P.own_hi > P.own_lo) {
        P.first_tile = (int32_t)(P.own_lo / LG::TILE);
        P.n_tiles = (int32_t)((P.own_hi - 1) / LG::TILE) - P.first_tile + 1;
    }
    if (P.n_tiles <= 0) return CQG_OK;
    const int smem = Lean2GLayout<LG>::TOTAL;  // tile + masks + interval table + dictionary + per-warp accumulators
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lean2g_kernel<LG, MINB>, LG::THREADS, smem));
    if (per_sm < 1) return fail(CQG_ERR_CUDA, "lean2g kernel does not fit");
    int grid = std::min(P.n_tiles, c.sms * per_sm);
    // the same kernel compiled for this query's shape, when the run-time compiler is there (else the generic one)
    char name[160];
    snprintf(name, sizeof name, "cqg::lean2g_kernel<cqg::Geo<%d, %d, %d, %d>, %d>", LG::THREADS, LG::TILE, LG::STAGES, LG::OVER, MINB);
    if (cudaKernel_t jk = cqg_jit::get(cqg_jit::lean_shape_defs(P), "cqg_lean2g.cuh", name, P.own_hi - P.own_lo)) {
        void* args[] = {(void*)&P};
        if (cudaLaunchKernel((const void*)jk, dim3(grid), dim3(LG::THREADS), args, (size_t)smem, st) == cudaSuccess) {
            g_launches++;
    g_family[KF_LEAN2G]++;
            return CQG_OK;
        }
        cudaGetLastError();
    }
    lean2g_kernel<LG, MINB><<<grid, LG::THREADS, smem, st>>>(P);
    g_launches++;
    g_family[KF_LEAN2G]++;
    CU(cudaGetLastError());
    return CQG_OK;
}

// lean2k_kernel (cqg_lean2k.cuh) exists only compiled for the query's shape. *launched = 0: no run-time compiler (or the
// shape does not compile): the caller runs lean2g_kernel instead.
template <class LG, int MINB>
static int launch_lean2k_geo(const DevPlan& P0, cudaStream_t st, int* launched) {
    *launched = 0;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevPlan P = P0;
    if (P.own_hi > P.own_lo) {
        P.first_tile = (int32_t)(P.own_lo / LG::TILE);
        P.n_tiles = (int32_t)((P.own_hi - 1) / LG::TILE) - P.first_tile + 1;
    }
    if (P.n_tiles <= 0) {
        *launched = 1;
        return CQG_OK;
    }
    const int smem = Lean2KLayout<LG>::TOTAL;  // tile + masks + dictionary + thread-private accumulators
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    char name[160];
    snprintf(name, sizeof name, "cqg::lean2k_kernel<cqg::Geo<%d, %d, %d, %d>, %d>", LG::THREADS, LG::TILE, LG::STAGES, LG::OVER, MINB);
    cudaKernel_t jk = cqg_jit::get(cqg_jit::lean_shape_defs(P), "cqg_lean2k.cuh", name, P.own_hi - P.own_lo);
    if (!jk) return CQG_OK;
    int per_sm = 0;
    if (cudaFuncSetAttribute((const void*)jk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)jk, LG::THREADS, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return CQG_OK;
    }
    const int grid = std::min(P.n_tiles, c.sms * per_sm);
    void* args[] = {(void*)&P};
    if (cudaLaunchKernel((const void*)jk, dim3(grid), dim3(LG::THREADS), args, (size_t)smem, st) != cudaSuccess) {
        cudaGetLastError();
        return CQG_OK;
    }
    g_launches++;
    g_family[KF_LEAN2K]++;
    *launched = 1;
    return CQG_OK;
}

template <class LG, int MINB>
static int launch_leanhc_geo(const DevPlan& P0, cudaStream_t st) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    DevPlan P = P0;
    if (P.own_hi > P.own_lo) {
        P.first_tile = (int32_t)(P.own_lo / LG::TILE);
        P.n_tiles = (int32_t)((P.own_hi - 1) / LG::TILE) - P.first_tile + 1;
    }
    if (P.n_tiles <= 0) return CQG_OK;
    const int smem = LeanHCLayout<LG>::TOTAL;
    LaunchCfg& c = g_cfg[dev & 63];
    if (!c.ready) {
        CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
        c.ready = true;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, leanhc_kernel<LG, MINB>, LG::THREADS, smem));
    if (per_sm < 1) return fail(CQG_ERR_CUDA, "leanhc kernel does not fit");
    const int grid = std::min(P.n_tiles, c.sms * per_sm);
    // the same kernel compiled for this query's shape, when the run-time compiler is there (else the generic one)
    char name[160];
    snprintf(name, sizeof name, "cqg::leanhc_kernel<cqg::Geo<%d, %d, %d, %d>, %d>", LG::THREADS, LG::TILE, LG::STAGES, LG::OVER, MINB);
    if (cudaKernel_t jk = cqg_jit::get(cqg_jit::lean_shape_defs(P), "cqg_leanhc.cuh", name, P.own_hi - P.own_lo)) {
        void* args[] = {(void*)&P};
        if (cudaLaunchKernel((const void*)jk, dim3(grid), dim3(LG::THREADS), args, (size_t)smem, st) == cudaSuccess) {
            g_launches++;
    g_family[KF_LEANHC]++;
            return CQG_OK;
        }
        cudaGetLastError();
    }
    leanhc_kernel<LG, MINB><<<grid, LG::THREADS, smem, st>>>(P);
    g_launches++;
    g_family[KF_LEANHC]++;
    CU(cudaGetLastError());
    return CQG_OK;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// the lean kernel hands tiles over by index: its tile size must be the general kernel's
static int launch_lean(const DevPlan& P, cudaStream_t st) {
    if (P.simple == 2) {
        bool mm = false;
        for (int a = 0; a < P.l_nagg; a++) mm = mm || P.aggs[P.l_agg[a]].func == CQG_AGG_MIN || P.aggs[P.l_agg[a]].func == CQG_AGG_MAX;
        if (P.lean_global) return launch_leanhc_geo<Geo<128, 16384, 1, 224>, 6>(P, st);  // many groups: packed global table
        if (mm) return launch_lean_geo<Geo<128, 16384, 1>, 4, true, false, true>(P, st);
        if (env_int("CQG_LEAN2", 1)) {  // few groups, COUNT/SUM/AVG
            if (P.lean_k) {  // one text key, <= 16 groups: the written-out loop, compiled for the query
                int launched = 0;
                const int minb = env_int("CQG_L2K_MINB", 8);  // CTAs per SM the kernel is compiled for (A/B runs)
                const int rc = minb == 6   ? launch_lean2k_geo<Geo<128, 16384, 1, 224>, 6>(P, st, &launched)
                               : minb == 8 ? launch_lean2k_geo<Geo<128, 16384, 1, 224>, 8>(P, st, &launched)
                                           : launch_lean2k_geo<Geo<128, 16384, 1, 224>, 7>(P, st, &launched);
                if (rc || launched) return rc;
            }
            return launch_lean2g_geo<Geo<128, 16384, 1, 224>, 6>(P, st);
        }
        return launch_lean_geo<Geo<128, 16384, 1>, 5, true, false, false>(P, st);
    }
    const bool oneleaf = P.l_nprog == 1 && P.l_nleaf == 1 && P.l_leaf[0].kind == 0 && P.l_leaf[0].slot == 0 && P.l_nagg == 0 &&
                         P.nwantL == 1 && !P.signed_hint;
    bool mm0 = false;
    for (int a = 0; a < P.l_nagg; a++) mm0 = mm0 || P.aggs[P.l_agg[a]].func == CQG_AGG_MIN || P.aggs[P.l_agg[a]].func == CQG_AGG_MAX;
    const int lean2 = env_int("CQG_LEAN2", 1);  // 0: the first lean kernel (A/B runs)
    if (lean2 && !mm0) {
        using LS = Geo<128, 16384, 1, 224>;  // rows of 64 bytes and more are handed over anyway: a short overlap
        if (P.crlf) return oneleaf ? launch_lean2_geo<LS, 9, true, -1, true>(P, st) : launch_lean2_geo<LS, 8, false, -1, true>(P, st);
        if (oneleaf) {
            switch (P.gap[0]) {
                case 0: return launch_lean2_geo<LS, 9, true, 0>(P, st);
                case 1: return launch_lean2_geo<LS, 9, true, 1>(P, st);
                case 2: return launch_lean2_geo<LS, 9, true, 2>(P, st);
                case 3: return launch_lean2_geo<LS, 9, true, 3>(P, st);
                case 4: return launch_lean2_geo<LS, 9, true, 4>(P, st);
                case 5: return launch_lean2_geo<LS, 9, true, 5>(P, st);
                case 6: return launch_lean2_geo<LS, 9, true, 6>(P, st);
                case 7: return launch_lean2_geo<LS, 9, true, 7>(P, st);
                default: return launch_lean2_geo<LS, 9, true, -1>(P, st);
            }
        }
        return launch_lean2_geo<LS, 8, false, -1>(P, st);
    }
    if (oneleaf) {
        static int variant = -1;
        if (variant < 0) {
            const char* e = getenv("CQG_LEAN_GEO");
            variant = e ? atoi(e) : 0;
        }
        if (variant == 1) return launch_lean_geo<Geo<128, 16384, 2>, 6, false, true, false>(P, st);
        if (variant == 2) return launch_lean_geo<Geo<128, 16384, 1>, 12, false, true, false>(P, st);
        return launch_lean_geo<Geo<128, 16384, 1>, 9, false, true, false>(P, st);
    }
    bool mm = false;
    for (int a = 0; a < P.l_nagg; a++) mm = mm || P.aggs[P.l_agg[a]].func == CQG_AGG_MIN || P.aggs[P.l_agg[a]].func == CQG_AGG_MAX;
    if (mm) return launch_lean_geo<Geo<128, 16384, 1>, 6, false, false, true>(P, st);
    return launch_lean_geo<Geo<128, 16384, 1>, 8, false, false, false>(P, st);
}

// [4][10000] doubles: mant / 10^fd, correctly rounded (one IEEE division each), per device
static const double* decimal_table(int dev) {
    static double* tab[64];
    if (!tab[dev & 63]) {
        std::vector<double> h(4 * 10000);
        const double p10[4] = {1.0, 10.0, 100.0, 1000.0};
        for (int fd = 0; fd < 4; fd++)
            for (int m = 0; m < 10000; m++) h[(size_t)fd * 10000 + m] = (double)m / p10[fd];
        double* d = nullptr;
        if (cudaMalloc((void**)&d, h.size() * 8) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        tab[dev & 63] = d;
    }
    return tab[dev & 63];
}

static int launch_scan(const DevPlan& P, int table_bytes, cudaStream_t st) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (P.n_tiles <= 0) return CQG_OK;
    int smem = scan_smem_bytes(table_bytes);
    // wanted fields in registers when there are few of them; the 16-slot build otherwise
    if (P.nwantL <= 4) return launch_scan_nw<4>(P, smem, dev, st);
    return launch_scan_nw<kMaxSlots>(P, smem, dev, st);
}

// ------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------
struct HostPlan {
    DevPlan P{};
    std::vector<uint8_t> entry_init;
    std::vector<uint8_t> pool;  // host copy of the predicate's string pool
    DevBuf d_entry_init, d_code, d_consts, d_pool;
    DevBuf d_scalars;  // errflags, rows_scanned, gcount, sel_count, jrow_count, jclass[2]
    int table_smem_bytes = 0;
    std::vector<uint8_t> packed_init;  // image of an empty packed line (lean GROUP BY, global mode)
    DevBuf d_packed_init;
    bool start_global = false;  // the sampled rows already hold more distinct keys than a CTA dictionary numbers
    bool few_text_keys = false; // one key column, and the sampled rows hold <= 16 distinct keys, all of them plain text
};

struct ScalarBlock {
    unsigned errflags;
    unsigned jclass[2];
    unsigned pad;
    unsigned long long rows_scanned, gcount, sel_count, jrow_count, def_tile_count, def_row_count, pcount, n_dense;
};

static void shard_range(const cqg_table* t, uint64_t& lo, uint64_t& hi) {
    if (t->has_range) {
        lo = std::max<uint64_t>(t->range_lo, t->data_start);
        hi = std::min<uint64_t>(t->range_hi, t->size);
        if (hi < lo) hi = lo;
        return;
    }
    unsigned __int128 sz = t->size;
    lo = (uint64_t)(sz * (unsigned)t->shard_index / (unsigned)t->shard_count);
    hi = (uint64_t)(sz * (unsigned)(t->shard_index + 1) / (unsigned)t->shard_count);
    if (lo < t->data_start) lo = t->data_start;
    if (hi < lo) hi = lo;
}

static void set_file(DevPlan& P, const cqg_table* t, bool whole) {
    P.data = t->d_data;
    P.size = t->size;
    uint64_t lo, hi;
    if (whole) {
        lo = t->data_start;
        hi = t->size;
    } else {
        shard_range(t, lo, hi);
    }
    P.own_lo = lo;
    P.own_hi = hi;
    P.global_base = t->global_base;
    P.delim = (uint8_t)t->cfg.delimiter;
    P.quote = (uint8_t)t->cfg.quote;
    // the mask fast paths need: a quote that is '"' or a control byte (so that the "special byte" tests see it:
    // kLeanSpecialXor in cqg_lean.cuh, and the general kernel's test for bytes below 0x23), a delimiter that is
    // neither blank nor below 0x23, delimiter != quote
    unsigned d = P.delim, q = P.quote;
    P.exact_only = (!(q == 0x22u || q < 0x20u) || d < 0x23u || d == q || d >= 0x80u) ? 1 : 0;
    if (hi > lo) {
        P.first_tile = (int32_t)(lo / ScanGeo::TILE);
        P.n_tiles = (int32_t)((hi - 1) / ScanGeo::TILE) - P.first_tile + 1;
    } else {
        P.first_tile = 0;
        P.n_tiles = 0;
    }
}

// register column `col` of one side in the slot map
static int want_column(DevPlan& P, int col) {
    if (col < 0) return CQG_OK;
    if (col >= kMaxQueryCols) return fail(CQG_ERR_UNSUPPORTED, "column index %d beyond %d", col, kMaxQueryCols);
    if (col >= P.n_cols_total) return CQG_OK;  // reads as NULL
    if (P.colslot[col] != -1) return CQG_OK;
    P.colslot[col] = -2;  // marked, numbered later
    return CQG_OK;
}

// shortest decimal form of a double with <= 15 significant digits: value = mant / 10^fd
static bool exact_decimal(double x, long long& mant, int& fd) {
    if (!(x >= 0.0) || x > 1e9) return false;
    for (int k = 0; k <= 6; k++) {
        double scaled = x * std::pow(10.0, k);
        double r = std::nearbyint(scaled);
        if (r > 9e14) return false;
        char a[64], b[64];
        snprintf(a, sizeof a, "%.17g", (double)r / std::pow(10.0, k));
        snprintf(b, sizeof b, "%.17g", x);
        if (strcmp(a, b) == 0) {  // r / 10^k rounds to exactly x: x is the double of that decimal
            mant = (long long)r;
            fd = k;
            return true;
        }
    }
    return false;
}

// one F_CMP leaf of a fused program -> a lean leaf; false when the lean kernel cannot evaluate it
static bool make_lean_leaf(DevPlan& P, const FInsn& in, LeanLeaf& L) {
    int op = in.n;
    int col_ref = in.a, const_ref = in.b;
    if ((col_ref & kRefConst) && col_ref != kRefNull) {  // literal on the left: mirror the operator
        std::swap(col_ref, const_ref);
        op = op == CQG_OP_GT ? CQG_OP_LT : op == CQG_OP_LT ? CQG_OP_GT : op == CQG_OP_GE ? CQG_OP_LE
           : op == CQG_OP_LE ? CQG_OP_GE : op;
    }
    if (col_ref == kRefNull || (col_ref & kRefConst) || !(const_ref & kRefConst) || const_ref == kRefNull) return false;
    int slot = (col_ref >= 0 && col_ref < kMaxQueryCols) ? P.colslot[col_ref] : -1;
    if (slot < 0 || slot >= P.nwantL || slot >= 4) return false;
    const DConst& c = P.consts_inl[const_ref & 0x1fff];
    memset(&L, 0, sizeof L);
    L.slot = slot;
    if (c.type == CQG_TYPE_STRING) {
        if (op != CQG_OP_EQ && op != CQG_OP_NE) return false;
        if (c.len < 1 || c.len > 16) return false;
        return false;  // filled by the caller, which owns the host copy of the string pool
    }
    long long cm;
    int cfd;
    if (c.type == CQG_TYPE_INTEGER) {
        if (c.bits < 0 || c.bits > 1000000000ll) return false;
        cm = c.bits;
        cfd = 0;
    } else if (c.type == CQG_TYPE_DOUBLE) {
        double d;
        memcpy(&d, &c.bits, 8);
        if (!exact_decimal(d, cm, cfd) || cm > 1000000000ll) return false;
    } else {
        return false;
    }
    L.kind = 0;
    for (int fd = 0; fd < 4; fd++) {
        int K = std::max(fd, cfd);
        long long A = 1, B = cm;
        for (int i = 0; i < K - fd; i++) A *= 10;
        for (int i = 0; i < K - cfd; i++) B *= 10;
        L.A[fd] = (uint32_t)A;
        // integers on both sides: a >= b  <=>  a > b - 1
        L.LB[fd] = B + (op == CQG_OP_GE ? -1 : op == CQG_OP_LE ? 1 : 0);
    }
    L.lop = (op == CQG_OP_GT || op == CQG_OP_GE) ? 0 : (op == CQG_OP_LT || op == CQG_OP_LE) ? 1 : op == CQG_OP_EQ ? 2 : 3;
    return true;
}

// DevPlan::simple — see cqg_plan.cuh. Called before fused refs are rewritten to slots.
// `pool`: host copy of the predicate's string pool (DConst::bits are offsets into it).
static void plan_simple_route(DevPlan& P, const std::vector<uint8_t>& pool) {
    P.simple = 0;
    P.l_nleaf = P.l_nprog = 0;
    if (P.join || P.mode != SCAN_AGG || P.exact_only || P.nwantL > 4 || P.ngc > 4) return;
    P.l_nagg = 0;
    for (int a = 0; a < P.naggs; a++) {
        if (P.aggs[a].off < 0) continue;  // COUNT(*), COUNT(col), unknown column: no state
        if (P.aggs[a].slot < 0 || P.aggs[a].slot >= 4 || P.l_nagg >= 4) return;
        P.l_agg[P.l_nagg++] = a;
    }
    if (P.pred_kind == 2) return;
    if (P.pred_kind == 1) {
        for (int k = 0; k < P.n_fcode; k++) {
            const FInsn& in = P.fcode_inl[k];
            if (P.l_nprog >= 16) return;
            if (in.op == F_AND) P.l_prog[P.l_nprog++] = -1;
            else if (in.op == F_OR) P.l_prog[P.l_nprog++] = -2;
            else if (in.op == F_NOT) P.l_prog[P.l_nprog++] = -3;
            else if (in.op == F_CMP) {
                if (P.l_nleaf >= kMaxLeanLeaf) return;
                LeanLeaf& L = P.l_leaf[P.l_nleaf];
                if (!make_lean_leaf(P, in, L)) {
                    // text equality: pack the literal like a key part
                    int op = in.n, col_ref = in.a, const_ref = in.b;
                    if ((col_ref & kRefConst) && col_ref != kRefNull) std::swap(col_ref, const_ref);
                    if (col_ref == kRefNull || (col_ref & kRefConst) || !(const_ref & kRefConst) || const_ref == kRefNull) return;
                    int slot = (col_ref >= 0 && col_ref < kMaxQueryCols) ? P.colslot[col_ref] : -1;
                    const DConst& c = P.consts_inl[const_ref & 0x1fff];
                    if (slot < 0 || slot >= P.nwantL || slot >= 4 || c.type != CQG_TYPE_STRING) return;
                    if ((op != CQG_OP_EQ && op != CQG_OP_NE) || c.len < 1 || c.len > 16) return;
                    if ((size_t)c.bits + c.len > pool.size()) return;
                    const uint8_t* sp = pool.data() + c.bits;
                    // a literal that could be typed as a number/date is not a STRING const; blanks cannot
                    // occur in a clean tile's field, so such a literal simply never matches there
                    memset(&L, 0, sizeof L);
                    L.slot = slot;
                    L.kind = op == CQG_OP_EQ ? 1 : 2;
                    L.slen = (int32_t)c.len;
                    for (uint32_t i = 0; i < c.len; i++) {
                        if (i < 8) L.w0 |= (uint64_t)sp[i] << (8 * i);
                        else L.w1 |= (uint64_t)sp[i] << (8 * (i - 8));
                    }
                }
                P.l_prog[P.l_nprog++] = (int8_t)P.l_nleaf;
                P.l_nleaf++;
            } else {
                return;  // IN, LIKE, TRUE/FALSE: general kernel
            }
        }
    }
    P.simple = P.ngc == 0 ? 1 : 2;  // 2: lean GROUP BY (per-CTA dictionary, up to 64 groups per CTA)
}

static int number_slots(DevPlan& P, const std::vector<uint8_t>& pool) {
    P.nwantL = P.nwantR = 0;
    for (int c = 0; c < P.n_left_cols && c < kMaxQueryCols; c++)
        if (P.colslot[c] == -2) {
            if (P.nwantL >= kMaxSlots) return fail(CQG_ERR_UNSUPPORTED, "more than %d columns referenced", kMaxSlots);
            P.wantL[P.nwantL++] = (int16_t)c;
        }
    for (int c = P.n_left_cols; c < P.n_cols_total && c < kMaxQueryCols; c++)
        if (P.colslot[c] == -2) {
            if (P.nwantR >= kMaxSlots) return fail(CQG_ERR_UNSUPPORTED, "more than %d columns referenced", kMaxSlots);
            P.wantR[P.nwantR++] = (int16_t)(c - P.n_left_cols);
        }
    for (int k = 0; k < P.nwantL; k++) P.colslot[P.wantL[k]] = (int16_t)k;
    for (int k = 0; k < P.nwantR; k++) P.colslot[P.n_left_cols + P.wantR[k]] = (int16_t)(P.nwantL + k);
    for (int k = 0; k < P.nwantL; k++) P.gap[k] = (int16_t)(k == 0 ? P.wantL[0] : P.wantL[k] - P.wantL[k - 1]);
    // operators address fields by slot from here on
    auto slot_of = [&](int col) { return (col < 0 || col >= kMaxQueryCols) ? -1 : (int)P.colslot[col]; };
    for (int a = 0; a < P.naggs; a++) P.aggs[a].slot = slot_of(P.aggs[a].col);
    for (int g = 0; g < P.ngc; g++) P.gslot[g] = (int16_t)slot_of(P.gcol[g]);
    auto ref_slot = [&](int16_t ref) -> int16_t {
        if (ref == kRefNull || (ref & kRefConst)) return ref;
        int sl = slot_of(ref);
        return (int16_t)(sl < 0 ? kRefNull : (kRefSlot | sl));
    };
    plan_simple_route(P, pool);
    if (P.pred_kind == 1) {
        for (int k = 0; k < P.n_fcode; k++) {
            FInsn& in = P.fcode_inl[k];
            if (in.op == F_CMP || in.op == F_LIKE || in.op == F_ILIKE) {
                in.a = ref_slot(in.a);
                in.b = ref_slot(in.b);
            } else if (in.op == F_IN || in.op == F_NOT_IN) {
                in.a = ref_slot(in.a);
                for (int j = 0; j < in.n; j++) P.frefs_inl[in.b + j] = ref_slot(P.frefs_inl[in.b + j]);
            }
        }
    }
    return CQG_OK;
}

// cqg_insn_t postfix -> device program(s)
static int compile_predicate(HostPlan& hp, const cqg_predicate_t& w, cudaStream_t st) {
    DevPlan& P = hp.P;
    P.pred_kind = 0;
    if (!w.code || w.n_code <= 0) return CQG_OK;
    // constants + string pool
    std::vector<DConst> consts((size_t)std::max(w.n_consts, 1));
    std::vector<uint8_t>& pool = hp.pool;
    pool.assign(8, 0);
    for (int i = 0; i < w.n_consts; i++) {
        const cqg_value_t& c = w.consts[i];
        DConst d{};
        d.type = c.type;
        switch (c.type) {
            case CQG_TYPE_INTEGER: d.bits = c.int_value; break;
            case CQG_TYPE_DOUBLE: memcpy(&d.bits, &c.double_value, 8); break;
            case CQG_TYPE_DATE:
                d.bits = ((long long)c.date_value.year << 16) | ((long long)c.date_value.month << 8) | c.date_value.day;
                break;
            case CQG_TYPE_STRING: {
                const char* s = c.string_value ? c.string_value : "";
                d.len = (uint32_t)strlen(s);
                d.bits = (long long)pool.size();
                pool.insert(pool.end(), s, s + d.len);
                pool.push_back(0);
                break;
            }
            default: d.type = CQG_TYPE_NULL; break;
        }
        consts[(size_t)i] = d;
    }
    // validate + collect columns; try the fused form
    struct Item {
        bool is_value;
        int ref;  // simple operand reference, or -1 for a computed value
    };
    std::vector<Item> stack;
    std::vector<FInsn> fcode;
    std::vector<int16_t> frefs;
    bool fusable = true;
    int depth_max = 0;
    for (int pc = 0; pc < w.n_code; pc++) {
        const cqg_insn_t in = w.code[pc];
        auto need = [&](size_t n) { return stack.size() >= n; };
        switch (in.op) {
            case CQG_OP_COL: {
                int rc = want_column(P, in.a);
                if (rc) return rc;
                int ref = (in.a < 0 || in.a >= P.n_cols_total) ? kRefNull : in.a;
                stack.push_back({true, ref});
                break;
            }
            case CQG_OP_CONST:
                if (in.a < 0 || in.a >= w.n_consts) return fail(CQG_ERR_ARG, "constant index %d out of range", in.a);
                stack.push_back({true, kRefConst | in.a});
                break;
            case CQG_OP_ADD: case CQG_OP_SUB: case CQG_OP_MUL: case CQG_OP_DIV: case CQG_OP_MOD: case CQG_OP_BAND:
            case CQG_OP_BOR: case CQG_OP_BXOR: case CQG_OP_ARITH_NULL:
                if (!need(2)) return fail(CQG_ERR_ARG, "predicate stack underflow at %d", pc);
                stack.pop_back();
                stack.back() = {true, -1};
                fusable = false;
                break;
            case CQG_OP_NEG:
                if (!need(1)) return fail(CQG_ERR_ARG, "predicate stack underflow at %d", pc);
                stack.back() = {true, -1};
                fusable = false;
                break;
            case CQG_OP_POS:
                if (!need(1)) return fail(CQG_ERR_ARG, "predicate stack underflow at %d", pc);
                break;
            case CQG_OP_EQ: case CQG_OP_NE: case CQG_OP_GT: case CQG_OP_LT: case CQG_OP_GE: case CQG_OP_LE:
            case CQG_OP_LIKE: case CQG_OP_ILIKE: {
                if (!need(2)) return fail(CQG_ERR_ARG, "predicate stack underflow at %d", pc);
                Item r = stack.back();
                stack.pop_back();
                Item l = stack.back();
                stack.pop_back();
                if (!l.is_value || !r.is_value) return fail(CQG_ERR_ARG, "comparison of a condition at %d", pc);
                if (l.ref < 0 || r.ref < 0) fusable = false;
                int fop = in.op == CQG_OP_LIKE ? F_LIKE : in.op == CQG_OP_ILIKE ? F_ILIKE : F_CMP;
                fcode.push_back({(int16_t)fop, (int16_t)l.ref, (int16_t)r.ref, (int16_t)in.op});
                stack.push_back({false, -1});
                break;
            }
            case CQG_OP_IN: case CQG_OP_NOT_IN: {
                if (in.a < 0 || !need((size_t)in.a + 1)) return fail(CQG_ERR_ARG, "IN list underflow at %d", pc);
                int first = (int)frefs.size();
                for (int k = 0; k < in.a; k++) {
                    Item it = stack[stack.size() - (size_t)in.a + (size_t)k];
                    if (!it.is_value) return fail(CQG_ERR_ARG, "IN item is a condition at %d", pc);
                    if (it.ref < 0) fusable = false;
                    frefs.push_back((int16_t)it.ref);
                }
                stack.resize(stack.size() - (size_t)in.a);
                Item l = stack.back();
                stack.pop_back();
                if (!l.is_value) return fail(CQG_ERR_ARG, "IN operand is a condition at %d", pc);
                if (l.ref < 0) fusable = false;
                fcode.push_back({(int16_t)(in.op == CQG_OP_IN ? F_IN : F_NOT_IN), (int16_t)l.ref, (int16_t)first, (int16_t)in.a});
                stack.push_back({false, -1});
                break;
            }
            case CQG_OP_AND: case CQG_OP_OR:
                if (!need(2) || stack.back().is_value || stack[stack.size() - 2].is_value)
                    return fail(CQG_ERR_ARG, "AND/OR operands at %d", pc);
                stack.pop_back();
                fcode.push_back({(int16_t)(in.op == CQG_OP_AND ? F_AND : F_OR), 0, 0, 0});
                break;
            case CQG_OP_NOT:
                if (!need(1) || stack.back().is_value) return fail(CQG_ERR_ARG, "NOT operand at %d", pc);
                fcode.push_back({F_NOT, 0, 0, 0});
                break;
            case CQG_OP_TRUE: case CQG_OP_FALSE:
                stack.push_back({false, -1});
                fcode.push_back({(int16_t)(in.op == CQG_OP_TRUE ? F_TRUE : F_FALSE), 0, 0, 0});
                break;
            case CQG_OP_POP:
                if (!need(1)) return fail(CQG_ERR_ARG, "POP underflow at %d", pc);
                stack.pop_back();
                fusable = false;
                break;
            default: return fail(CQG_ERR_ARG, "unknown opcode %d at %d", in.op, pc);
        }
        depth_max = std::max(depth_max, (int)stack.size());
    }
    if (stack.size() != 1 || stack.back().is_value) return fail(CQG_ERR_ARG, "predicate does not reduce to one condition");
    if (depth_max >= kStackMax - 1) return fail(CQG_ERR_UNSUPPORTED, "predicate nesting deeper than %d", kStackMax - 2);
    if (depth_max > 30) fusable = false;
    if (fcode.size() > (size_t)kMaxFused || frefs.size() > (size_t)kMaxFusedRefs || consts.size() > (size_t)kMaxFusedConsts)
        fusable = false;

    CU(hp.d_consts.alloc(consts.size() * sizeof(DConst), st));
    CU(cudaMemcpyAsync(hp.d_consts.p, consts.data(), consts.size() * sizeof(DConst), cudaMemcpyHostToDevice, st));
    CU(hp.d_pool.alloc(pool.size(), st));
    CU(cudaMemcpyAsync(hp.d_pool.p, pool.data(), pool.size(), cudaMemcpyHostToDevice, st));
    P.pred.consts = hp.d_consts.as<DConst>();
    P.pred.pool = hp.d_pool.as<uint8_t>();
    if (fusable) {
        for (size_t k = 0; k < fcode.size(); k++) P.fcode_inl[k] = fcode[k];
        for (size_t k = 0; k < frefs.size(); k++) P.frefs_inl[k] = frefs[k];
        for (size_t k = 0; k < consts.size(); k++) P.consts_inl[k] = consts[k];
        P.n_fcode = (int)fcode.size();
        P.pred_kind = 1;
    } else {
        CU(hp.d_code.alloc((size_t)w.n_code * sizeof(cqg_insn_t), st));
        CU(cudaMemcpyAsync(hp.d_code.p, w.code, (size_t)w.n_code * sizeof(cqg_insn_t), cudaMemcpyHostToDevice, st));
        P.pred.code = hp.d_code.as<cqg_insn_t>();
        P.pred.n_code = w.n_code;
        P.pred_kind = 2;
    }
    // (sources are pageable host memory: cudaMemcpyAsync has staged them when it returns, so the vectors may go;
    //  the copies themselves are ordered before the scan on the same stream)
    return CQG_OK;
}

static void layout_entry(HostPlan& hp) {
    DevPlan& P = hp.P;
    int off = kOffKeys + 16 * P.ngc;
    for (int a = 0; a < P.naggs; a++) {
        AggSpec& s = P.aggs[a];
        if (s.func == CQG_AGG_COUNT_STAR || s.func == CQG_AGG_COUNT || s.col < 0) {
            s.off = -1;
            continue;
        }
        s.off = off;
        off += (s.func == CQG_AGG_SUM || s.func == CQG_AGG_AVG) ? 32 : 48;
    }
    P.entry_bytes = (off + 15) / 16 * 16;
    hp.entry_init.assign((size_t)P.entry_bytes, 0);
    uint64_t ones = ~0ull;
    memcpy(hp.entry_init.data() + kOffFirst, &ones, 8);
    for (int a = 0; a < P.naggs; a++) {
        const AggSpec& s = P.aggs[a];
        if (s.off < 0) continue;
        if (s.func == CQG_AGG_MIN || s.func == CQG_AGG_MAX) {
            uint64_t* st = (uint64_t*)(hp.entry_init.data() + s.off);
            uint64_t empty = s.func == CQG_AGG_MIN ? ~0ull : 0ull;
            st[0] = ~0ull;
            st[1] = empty;
            st[2] = empty;
            st[3] = ~0ull;
            st[4] = empty;
            st[5] = 0;
        }
    }
}

// the first bytes after the header line, for layout guesses (never for results)
static const std::vector<uint8_t>& table_sample(const cqg_table* t) {
    if (!t->sample_ready) {
        t->sample_ready = true;
        const size_t lo = t->data_start, n = t->size > lo ? std::min<size_t>(t->size - lo, 4096) : 0;
        t->sample.resize(n);
        if (n) {
            if (t->h_data) memcpy(t->sample.data(), t->h_data + lo, n);
            else if (cudaMemcpy(t->sample.data(), t->d_data + lo, n, cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                t->sample.clear();
            }
        }
    }
    return t->sample;
}

// PackedLayout (cqg_plan.cuh) of a lean GROUP BY plan. Key columns whose sampled fields all read as unsigned
// decimals (or are empty) get a narrow 8-byte slot; a text met there later is handed over, so the guess only
// costs speed, never a result.
static void layout_packed(HostPlan& hp, const cqg_table* t) {
    DevPlan& P = hp.P;
    memset(&P.pk, 0, sizeof P.pk);
    if (P.simple != 2 || P.ngc < 1 || P.ngc > 4 || P.l_nagg > 4) return;
    bool narrow[4] = {false, false, false, false};
    {
        const std::vector<uint8_t>& sm = table_sample(t);
        int seen[4] = {0, 0, 0, 0};
        bool numeric[4] = {true, true, true, true};
        size_t pos = 0;
        int rows = 0;
        std::vector<std::string> keys;
        while (pos < sm.size() && rows < 256) {
            size_t eol = pos;
            while (eol < sm.size() && sm[eol] != '\n' && sm[eol] != '\r') eol++;
            if (eol == sm.size()) break;  // incomplete last line of the sample
            if (eol > pos) {
                rows++;
                int col = 0;
                size_t fs = pos;
                std::string key;
                for (size_t k = pos; k <= eol; k++) {
                    if (k == eol || sm[k] == (uint8_t)t->cfg.delimiter) {
                        for (int g = 0; g < P.ngc; g++) {
                            if (P.gcol[g] == col) {
                                key.append((const char*)&sm[fs], k - fs);
                                key.push_back('\0');
                                seen[g]++;
                                for (size_t j = fs; j < k; j++)
                                    if (!((sm[j] >= '0' && sm[j] <= '9') || sm[j] == '.')) numeric[g] = false;
                                if (k - fs > 7) numeric[g] = false;
                            }
                        }
                        col++;
                        fs = k + 1;
                    }
                }
                keys.push_back(key);
            }
            pos = eol + 1;
        }
        for (int g = 0; g < P.ngc; g++) narrow[g] = seen[g] > 0 && numeric[g];
        // more distinct keys in the sample than a CTA dictionary numbers: start in global mode (a guess that costs
        // or saves one aborted launch, nothing else)
        std::sort(keys.begin(), keys.end());
        const size_t distinct = (size_t)(std::unique(keys.begin(), keys.end()) - keys.begin());
        hp.start_global = distinct >= 48;
        // lean2k_kernel: one key column of plain text (what cannot start a number, no blank at either end, <= 16 bytes)
        // with few values. Again only a guess about where to start: the kernels check every row themselves.
        bool plain = P.ngc == 1 && P.gslot[0] >= 0 && rows > 0 && distinct <= (size_t)kL2KGroups;
        for (size_t k = 0; plain && k < distinct; k++) {
            const std::string& key = keys[k];
            const size_t n = key.size() ? key.size() - 1 : 0;  // (the trailing NUL of the key part)
            if (n == 0) continue;
            const unsigned char c0 = (unsigned char)key[0], c1 = (unsigned char)key[n - 1];
            if (n > 16 || (c0 >= '0' && c0 <= '9') || c0 == '+' || c0 == '-' || c0 == '.' || c0 == ' ' || c1 == ' ') plain = false;
        }
        hp.few_text_keys = plain;
    }
    int w = 1;  // word 0: state / first okey | tags
    for (int g = 0; g < P.ngc; g++) {
        P.pk.key_word[g] = (int16_t)w;
        P.pk.key_wide[g] = narrow[g] ? 0 : 1;
        w += narrow[g] ? 1 : 2;
    }
    P.pk.id_words = w;
    int off = 8 * w;
    P.pk.count_off = off;
    off += 8;
    for (int a = 0; a < 4; a++) {
        P.pk.agg_off[a] = -1;
        P.pk.agg_key[a] = -1;
    }
    for (int a = 0; a < P.l_nagg; a++) {
        // an aggregate over a GROUP BY column is a function of the key and the count: no state
        const AggSpec& sp = P.aggs[P.l_agg[a]];
        for (int g = 0; g < P.ngc; g++)
            if (sp.slot >= 0 && P.gslot[g] == sp.slot) P.pk.agg_key[a] = (int16_t)g;
        if (getenv("CQG_HC_NO_DERIVED")) P.pk.agg_key[a] = -1;  // (measurement: keep a state for every aggregate)
    }
    for (int a = 0; a < P.l_nagg; a++) {
        const int f = P.aggs[P.l_agg[a]].func;
        if (P.pk.agg_key[a] < 0 && (f == CQG_AGG_SUM || f == CQG_AGG_AVG)) {
            P.pk.agg_off[a] = (int16_t)off;
            off += 8;
        }
    }
    off = (off + 15) / 16 * 16;
    for (int a = 0; a < P.l_nagg; a++) {
        const int f = P.aggs[P.l_agg[a]].func;
        if (P.pk.agg_key[a] < 0 && (f == CQG_AGG_MIN || f == CQG_AGG_MAX)) {
            P.pk.agg_off[a] = (int16_t)off;
            off += 16;
        }
    }
    P.pk.entry_bytes = (off + 31) / 32 * 32;  // whole 32-byte chunks (sectors); at most 2 + 8 + 1 + 4 words + 4 pairs = 184 -> 192
    hp.packed_init.assign((size_t)P.pk.entry_bytes, 0);
    for (int a = 0; a < P.l_nagg; a++) {
        const int f = P.aggs[P.l_agg[a]].func;
        if (P.pk.agg_off[a] >= 0 && (f == CQG_AGG_MIN || f == CQG_AGG_MAX)) {
            uint64_t* st = (uint64_t*)(hp.packed_init.data() + P.pk.agg_off[a]);
            st[0] = f == CQG_AGG_MIN ? ~0ull : 0ull;
            st[1] = ~0ull;
        }
    }
}

// common part of a plan over table t (left) [+ right]
static int build_plan(HostPlan& hp, const cqg_table* t, const cqg_query_t* q, cudaStream_t st) {
    DevPlan& P = hp.P;
    set_file(P, t, false);
    const cqg_table* rt = q->join.right;
    // a column index is not bounded by the header: a row with more fields than the header still
    // yields them (row->values[col], evaluator_core.c:165); without a join every index is "left"
    P.n_left_cols = rt ? (int)t->names.size() : kMaxQueryCols;
    P.n_cols_total = kMaxQueryCols;
    if (P.n_left_cols > kMaxQueryCols) return fail(CQG_ERR_UNSUPPORTED, "more than %d columns", kMaxQueryCols);
    for (int c = 0; c < kMaxQueryCols; c++) P.colslot[c] = -1;
    P.mode = q->mode == CQG_MODE_SELECT ? SCAN_SELECT : SCAN_AGG;
    if (q->n_group_cols < 0 || q->n_group_cols > CQG_MAX_GROUP_COLS || q->n_aggs < 0 || q->n_aggs > CQG_MAX_AGGS ||
        q->n_out_cols < 0 || q->n_out_cols > CQG_MAX_OUT_COLS)
        return fail(CQG_ERR_ARG, "query dimensions out of range");
    int rc;
    if (rt) {
        if (rt->cfg.delimiter != t->cfg.delimiter || rt->cfg.quote != t->cfg.quote)
            return fail(CQG_ERR_ARG, "joined tables must share one CSV dialect");
        P.join = 1;
        if (q->join.type < CQG_JOIN_INNER || q->join.type > CQG_JOIN_FULL) return fail(CQG_ERR_ARG, "unknown join type %d", q->join.type);
        P.join_type = q->join.type;
        // with an unresolved key column no pair matches (joins.c:54): RIGHT / FULL would have to emit the whole right
        // table, which the (empty) join table cannot enumerate
        if (P.join_type >= CQG_JOIN_RIGHT && (q->join.left_col < 0 || q->join.right_col < 0))
            return fail(CQG_ERR_UNSUPPORTED, "RIGHT / FULL JOIN on an unknown key column");
        P.rdata = rt->d_data;
        P.rsize = rt->size;
        P.jl_col = q->join.left_col;
        P.jr_col = q->join.right_col;
        if (P.jl_col >= P.n_left_cols) P.jl_col = -1;
        if ((rc = want_column(P, P.jl_col))) return rc;
    }
    if ((rc = compile_predicate(hp, q->where, st))) return rc;
    if (P.mode == SCAN_AGG) {
        P.ngc = q->n_group_cols;
        for (int g = 0; g < P.ngc; g++) {
            P.gcol[g] = (int16_t)q->group_cols[g];
            if ((rc = want_column(P, q->group_cols[g]))) return rc;
        }
        P.naggs = q->n_aggs;
        for (int a = 0; a < P.naggs; a++) {
            P.aggs[a].func = q->aggs[a].func;
            P.aggs[a].col = q->aggs[a].col;
            if (q->aggs[a].func < CQG_AGG_COUNT_STAR || q->aggs[a].func > CQG_AGG_MAX)
                return fail(CQG_ERR_ARG, "unknown aggregate %d", q->aggs[a].func);
            if (P.aggs[a].col >= P.n_cols_total) P.aggs[a].col = -1;
            if (q->aggs[a].func >= CQG_AGG_SUM && (rc = want_column(P, P.aggs[a].col))) return rc;
        }
        layout_entry(hp);
        P.scalar_regs = (P.ngc == 0 && P.naggs <= 4) ? 1 : 0;
        // shared-memory table: as many entries as fit a 24 KB budget (one when there is no GROUP BY)
        int cap = 1;
        if (P.ngc > 0) {
            cap = 1024;
            while (cap > 1 && cap * P.entry_bytes > 24 * 1024) cap >>= 1;
            if (cap < 8) cap = 0;
        }
        P.smem_cap = cap;
        hp.table_smem_bytes = cap * P.entry_bytes;
        CU(hp.d_entry_init.alloc(hp.entry_init.size(), st));
        CU(cudaMemcpyAsync(hp.d_entry_init.p, hp.entry_init.data(), hp.entry_init.size(), cudaMemcpyHostToDevice, st));
        P.entry_init = hp.d_entry_init.as<uint8_t>();
    }
    if ((rc = number_slots(P, hp.pool))) return rc;
    layout_packed(hp, t);
    P.crlf = 0;
    if (P.simple && env_int("CQG_CRLF", 1)) {
        // a CR in the head of the file: the lean kernels then take '\r' as a terminator (CR LF files stay on them). A guess
        // about the rest of the file that only costs speed: a CR they were not told about sends its tile to the general kernel
        const std::vector<uint8_t>& sm = table_sample(t);
        P.crlf = memchr(sm.data(), '\r', sm.size()) ? 1 : 0;
        // a field that starts with '-' or '+': the written-out COUNT-WHERE loop (ONELEAF) hands such fields over row by row
        // and gives the tile up after two; the general scalar loop decodes them (cqg_lean2.cuh: CQG_L2_SIGNED)
        P.signed_hint = 0;
        for (size_t k = 0; k + 1 < sm.size(); k++)
            if ((sm[k] == (uint8_t)t->cfg.delimiter || sm[k] == '\n') && (sm[k + 1] == '-' || sm[k + 1] == '+')) P.signed_hint = 1;
    }
    P.need_right_fields = P.nwantR > 0;
    CU(hp.d_scalars.alloc(sizeof(ScalarBlock), st));
    CU(cudaMemsetAsync(hp.d_scalars.p, 0, sizeof(ScalarBlock), st));
    ScalarBlock* sb = hp.d_scalars.as<ScalarBlock>();
    P.errflags = &sb->errflags;
    P.jclass = sb->jclass;
    P.rows_scanned = &sb->rows_scanned;
    P.gcount = &sb->gcount;
    P.sel_count = &sb->sel_count;
    P.jrow_count = &sb->jrow_count;
    P.pcount = &sb->pcount;
    return CQG_OK;
}

static const char* flag_text(unsigned f) {
    if (f & KERR_NUMERIC_RANGE) return "a decimal field has more than 19 significant digits";
    if (f & KERR_KEY_RANGE) return "a DOUBLE group key is outside the exactly rendered %.6f range";
    if (f & KERR_MINMAX_TIE) return "MIN/MAX tie between an INTEGER and a DOUBLE of equal value (order dependent in the reference)";
    if (f & KERR_STACK) return "predicate stack overflow";
    if (f & KERR_JOIN_MIXED) return "join key columns mix comparison classes (cross-type value_compare)";
    if (f & KERR_BIGINT) return "INTEGER beyond 2^53 compared as double";
    if (f & KERR_KEY_TAB) return "composite GROUP BY key contains a tab";
    if (f & KERR_JOIN_FANOUT) return "a row joins with 65536 or more rows";
    if (f & KERR_STR_LONG) return "MIN/MAX over a string longer than 256 KiB";
    if (f & KERR_OFFSET_RANGE) return "file larger than 32 TiB";
    return "unsupported input";
}

// ------------------------------------------------------------------------------------------
// join build
// ------------------------------------------------------------------------------------------
struct JoinState {
    DevBuf slots, row_off, row_next, scalars;
    uint64_t cap = 0, rows = 0;
    unsigned flags = 0, key_classes = 0;  // what the build wrote into the probe plan's scalar block (the scan clears it)
};

static int count_rows_device(const cqg_table* t, bool whole, cudaStream_t st, int64_t* out) {
    HostPlan hp;
    DevPlan& P = hp.P;
    set_file(P, t, whole);
    P.mode = SCAN_COUNT_ROWS;
    for (int c = 0; c < kMaxQueryCols; c++) P.colslot[c] = -1;
    CU(hp.d_scalars.alloc(sizeof(ScalarBlock), st));
    CU(cudaMemsetAsync(hp.d_scalars.p, 0, sizeof(ScalarBlock), st));
    ScalarBlock* sb = hp.d_scalars.as<ScalarBlock>();
    P.errflags = &sb->errflags;
    P.rows_scanned = &sb->rows_scanned;
    P.jclass = sb->jclass;
    P.gcount = &sb->gcount;
    P.sel_count = &sb->sel_count;
    P.jrow_count = &sb->jrow_count;
    int rc = launch_scan(P, 0, st);
    if (rc) return rc;
    ScalarBlock h;
    CU(cudaMemcpyAsync(&h, sb, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *out = (int64_t)h.rows_scanned;
    return CQG_OK;
}

// rows != nullptr: the build side is exactly these `n_rows` rows of the right file (offsets in device memory,
// the rows this rank owns in a hash-partitioned join) instead of the whole right table.
static int build_join(HostPlan& probe, JoinState& js, const cqg_table* rt, int right_col, cudaStream_t st,
                      const uint64_t* rows = nullptr, int64_t n_rows = 0) {
    int64_t nrows = n_rows;
    int rc = rows ? CQG_OK : count_rows_device(rt, true, st, &nrows);
    if (rc) return rc;
    js.rows = (uint64_t)nrows;
    uint64_t cap = 1024;
    while (cap < js.rows * 2) cap <<= 1;
    js.cap = cap;
    CU(js.slots.alloc(cap * sizeof(JoinSlot), st));
    CU(cudaMemsetAsync(js.slots.p, 0, cap * sizeof(JoinSlot), st));
    CU(js.row_off.alloc((js.rows + 1) * 8, st));
    CU(js.row_next.alloc((js.rows + 1) * 4, st));
    HostPlan hp;
    DevPlan& B = hp.P;
    set_file(B, rt, true);
    B.mode = SCAN_JOIN_BUILD;
    for (int c = 0; c < kMaxQueryCols; c++) B.colslot[c] = -1;
    B.n_left_cols = (int)rt->names.size();
    B.n_cols_total = B.n_left_cols;
    B.jr_col = right_col;
    if (right_col >= 0 && right_col < B.n_cols_total) {
        B.wantL[0] = (int16_t)right_col;
        B.nwantL = 1;
        B.colslot[right_col] = 0;
    }
    // the build shares the probe plan's scalar block so that class masks and flags add up
    B.errflags = probe.P.errflags;
    B.jclass = probe.P.jclass;
    B.rows_scanned = probe.P.jrow_count + 0;  // not used for results; keep a valid address
    CU(js.scalars.alloc(sizeof(ScalarBlock), st));
    CU(cudaMemsetAsync(js.scalars.p, 0, sizeof(ScalarBlock), st));
    ScalarBlock* sb = js.scalars.as<ScalarBlock>();
    B.rows_scanned = &sb->rows_scanned;
    B.gcount = &sb->gcount;
    B.sel_count = &sb->sel_count;
    B.jrow_count = &sb->jrow_count;
    B.jslots = js.slots.as<JoinSlot>();
    B.jcap = cap;
    B.jrow_off = js.row_off.as<uint64_t>();
    B.jrow_next = js.row_next.as<uint32_t>();
    B.jrow_cap = js.rows;
    if (js.rows >= 0xfffffff0ull) return fail(CQG_ERR_UNSUPPORTED, "right table has too many rows");
    if (right_col >= 0 && !rows) {
        rc = launch_scan(B, 0, st);
        if (rc) return rc;
    } else if (right_col >= 0 && n_rows > 0) {
        int grid = (int)std::min<uint64_t>(((uint64_t)n_rows + 127) / 128, 148 * 8);
        deferred_rows_kernel<<<grid, 128, 0, st>>>(B, rows, (uint64_t)n_rows);
        g_launches++;
        CU(cudaGetLastError());
    }
    {
        // the probe scan starts from a cleared scalar block: keep the build's error flags and key classes
        ScalarBlock hb{};
        CU(cudaMemcpyAsync(&hb, probe.d_scalars.p, sizeof hb, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        js.flags = hb.errflags & kFatalMask;
        js.key_classes = hb.jclass[1];
    }
    DevPlan& P = probe.P;
    P.jslots = B.jslots;
    P.jcap = cap;
    P.jrow_off = B.jrow_off;
    P.jrow_next = B.jrow_next;
    P.jrow_cap = js.rows;
    return CQG_OK;
}

// ------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------
// bump allocator: results hold millions of small strings, one calloc each would dominate
struct Arena {
    std::vector<void*> blocks;
    char* cur = nullptr;
    size_t left = 0;
    size_t next_block = 16u << 10;  // blocks grow 16 KB -> 4 MB: a one-group result must not cost a 4 MB calloc
    void* alloc(size_t n) {
        n = (n + 15) & ~(size_t)15;
        if (n == 0) n = 16;
        if (n > (1u << 20)) {
            void* p = calloc(1, n);
            blocks.push_back(p);
            return p;
        }
        if (n > left) {
            size_t sz = next_block;
            while (sz < n) sz <<= 1;
            if (next_block < (4u << 20)) next_block <<= 2;
            cur = (char*)calloc(1, sz);
            blocks.push_back(cur);
            left = sz;
        }
        void* p = cur;
        cur += n;
        left -= n;
        return p;
    }
    void adopt(void* p) { blocks.push_back(p); }  // a malloc'ed block the arena frees with the rest
    // the one big block of a many-group result (finish_aggregate_device): handed back to big_block_put when the result
    // is freed, so that the next such result writes into pages that are already there
    void* big = nullptr;
    size_t big_cap = 0;
    ~Arena();
};

// One spare result block (hundreds of MB for ~10^6 groups): first touch of fresh pages is a third of the time the block
// takes to reach the host. CQG_RESULT_CACHE=0: always malloc and free.
static std::mutex g_big_mu;
static void* g_big_spare = nullptr;
static size_t g_big_spare_cap = 0;
static void* big_block_get(size_t need, size_t* cap) {
    {
        std::lock_guard<std::mutex> lock(g_big_mu);
        if (g_big_spare && g_big_spare_cap >= need) {
            void* p = g_big_spare;
            *cap = g_big_spare_cap;
            g_big_spare = nullptr;
            g_big_spare_cap = 0;
            return p;
        }
    }
    const size_t sz = (need + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    void* p = nullptr;
    if (posix_memalign(&p, 2u << 20, sz) != 0) return nullptr;
    madvise(p, sz, MADV_HUGEPAGE);  // (2 MB pages: ~10^5 first touches of small pages is what the copy would wait for)
    *cap = sz;
    return p;
}
static void big_block_put(void* p, size_t cap) {
    const char* e = getenv("CQG_RESULT_CACHE");
    if (!(e && e[0] == '0')) {
        std::lock_guard<std::mutex> lock(g_big_mu);
        if (!g_big_spare || g_big_spare_cap < cap) {
            std::swap(p, g_big_spare);
            std::swap(cap, g_big_spare_cap);
        }
    }
    free(p);  // (the smaller of the two, or nullptr)
}
Arena::~Arena() {
    for (void* p : blocks) free(p);
    if (big) big_block_put(big, big_cap);
}

// CQG_TIMING=1: phase timings of the host side on stderr
struct PhaseTimer {
    bool on;
    double t0;
    static double now() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    PhaseTimer() {
        const char* e = getenv("CQG_TIMING");
        on = e && e[0] == '1';
        t0 = now();
    }
    void lap(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        double t = now();
        fprintf(stderr, "[cqg timing] %-28s %9.3f ms\n", what, t - t0);
        t0 = t;
    }
};

static cqg_result_t* new_result(Arena** arena_out) {
    cqg_result_t* r = (cqg_result_t*)calloc(1, sizeof(cqg_result_t));
    Arena* a = new Arena();
    r->arena = a;
    *arena_out = a;
    return r;
}

CQG_API void cqg_result_free(cqg_result_t* r) {
    if (!r) return;
    delete (Arena*)r->arena;
    free(r);
}

// host loops over millions of result cells (the finish of a ~2 M group result): split over a few threads.
// fn(chunk, lo, hi); chunks are contiguous and in order, so per-chunk outputs concatenate deterministically.
static unsigned par_chunk_count(size_t n) {
    if (n < (1u << 16)) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    return std::min<unsigned>(std::max(1u, hw), 16u);
}
template <class F>
static void par_chunks(size_t n, unsigned T, F&& fn) {
    if (T <= 1 || n == 0) {
        fn(0u, (size_t)0, n);
        return;
    }
    const size_t per = (n + T - 1) / T;
    std::vector<std::thread> th;
    for (unsigned k = 0; k < T; k++) {
        const size_t lo = (size_t)k * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([&fn, k, lo, hi] { fn(k, lo, hi); });
    }
    for (auto& x : th) x.join();
}

// turn OutCells into cqg_value_t, pulling string bytes from the host view when there is one,
// else gathering them on the device
static int cells_to_values(const cqg_table* t, const cqg_table* rt, const std::vector<OutCell>& cells, cqg_value_t* out,
                           Arena* arena, cudaStream_t st) {
    size_t n = cells.size();
    std::vector<uint64_t> refs;
    std::vector<uint32_t> lens;
    std::vector<uint64_t> offs;
    std::vector<size_t> which;
    uint64_t total = 0;
    const unsigned T = par_chunk_count(n);
    struct Chunk {
        std::vector<uint64_t> refs;
        std::vector<uint32_t> lens;
        std::vector<uint64_t> offs;  // chunk-local
        std::vector<size_t> which;
        uint64_t total = 0;
        size_t str_bytes = 0;
        char* block = nullptr;
    };
    std::vector<Chunk> ch(T);
    // pass 1: string storage each chunk needs (one arena block per chunk: the arena itself is not thread safe)
    par_chunks(n, T, [&](unsigned k, size_t lo, size_t hi) {
        size_t bytes = 0;
        for (size_t i = lo; i < hi; i++)
            if (cells[i].type == CQG_TYPE_STRING) bytes += ((size_t)cells[i].len + 1 + 15) & ~(size_t)15;
        ch[k].str_bytes = bytes;
    });
    for (unsigned k = 0; k < T; k++)
        if (ch[k].str_bytes) ch[k].block = (char*)arena->alloc(ch[k].str_bytes);
    // pass 2: the values
    par_chunks(n, T, [&](unsigned k, size_t lo, size_t hi) {
        Chunk& c_ = ch[k];
        char* cur = c_.block;
        for (size_t i = lo; i < hi; i++) {
            const OutCell& c = cells[i];
            cqg_value_t v;
            memset(&v, 0, sizeof v);
            v.type = c.type;
            switch (c.type) {
                case CQG_TYPE_INTEGER: v.int_value = (long long)c.payload; break;
                case CQG_TYPE_DOUBLE: memcpy(&v.double_value, &c.payload, 8); break;
                case CQG_TYPE_DATE:
                    v.date_value.year = (int)(c.payload >> 16);
                    v.date_value.month = (int)((c.payload >> 8) & 0xff);
                    v.date_value.day = (int)(c.payload & 0xff);
                    break;
                case CQG_TYPE_STRING: {
                    bool right = (c.payload >> 63) != 0;
                    uint64_t off = c.payload & 0x7fffffffffffffffull;
                    const cqg_table* src = right ? rt : t;
                    char* s = cur;
                    cur += ((size_t)c.len + 1 + 15) & ~(size_t)15;
                    s[c.len] = 0;
                    v.string_value = s;
                    if (src && src->h_data) {
                        memcpy(s, src->h_data + off, c.len);
                    } else {
                        c_.refs.push_back(c.payload);
                        c_.lens.push_back(c.len);
                        c_.offs.push_back(c_.total);
                        c_.which.push_back(i);
                        c_.total += c.len;
                    }
                    break;
                }
                default: v.type = CQG_TYPE_NULL; break;
            }
            out[i] = v;
        }
    });
    for (unsigned k = 0; k < T; k++) {
        for (size_t j = 0; j < ch[k].refs.size(); j++) {
            refs.push_back(ch[k].refs[j]);
            lens.push_back(ch[k].lens[j]);
            offs.push_back(total + ch[k].offs[j]);
            which.push_back(ch[k].which[j]);
        }
        total += ch[k].total;
    }
    if (!refs.empty()) {
        DevBuf d_refs, d_lens, d_offs, d_dst;
        CU(d_refs.alloc(refs.size() * 8, st));
        CU(d_lens.alloc(lens.size() * 4, st));
        CU(d_offs.alloc(offs.size() * 8, st));
        CU(d_dst.alloc(total + 8, st));
        CU(cudaMemcpyAsync(d_refs.p, refs.data(), refs.size() * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_lens.p, lens.data(), lens.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_offs.p, offs.data(), offs.size() * 8, cudaMemcpyHostToDevice, st));
        int grid = (int)std::min<size_t>((refs.size() + 127) / 128, 148 * 16);
        pack_strings_kernel<<<grid, 128, 0, st>>>(t->d_data, rt ? rt->d_data : nullptr, d_refs.as<uint64_t>(), d_lens.as<uint32_t>(),
                                                 d_offs.as<uint64_t>(), refs.size(), d_dst.as<uint8_t>());
        g_launches++;
        CU(cudaGetLastError());
        std::vector<uint8_t> bytes(total + 8);
        CU(cudaMemcpyAsync(bytes.data(), d_dst.p, total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        par_chunks(which.size(), par_chunk_count(which.size()), [&](unsigned, size_t lo, size_t hi) {
            for (size_t k = lo; k < hi; k++) memcpy((char*)out[which[k]].string_value, bytes.data() + offs[k], lens[k]);
        });
    }
    return CQG_OK;
}

static int run_fetch(const cqg_table* t, const cqg_table* rt, const int32_t* cols, int ncols, const std::vector<uint64_t>& loff,
                     const uint64_t* d_roff, unsigned* d_err, std::vector<OutCell>& cells, cudaStream_t st) {
    size_t n = loff.size();
    cells.assign(n * (size_t)ncols, OutCell{});
    if (n == 0 || ncols == 0) return CQG_OK;
    DevBuf d_loff, d_cells;
    CU(d_loff.alloc(n * 8, st));
    CU(cudaMemcpyAsync(d_loff.p, loff.data(), n * 8, cudaMemcpyHostToDevice, st));
    CU(d_cells.alloc(cells.size() * sizeof(OutCell), st));
    FetchParams F{};
    F.data = t->d_data;
    F.size = t->size;
    F.rdata = rt ? rt->d_data : nullptr;
    F.rsize = rt ? rt->size : 0;
    F.delim = (uint8_t)t->cfg.delimiter;
    F.quote = (uint8_t)t->cfg.quote;
    F.n_left_cols = rt ? (int)t->names.size() : 0x7fff;
    F.ncols = ncols;
    for (int c = 0; c < ncols; c++) F.cols[c] = (int16_t)((cols[c] < 0 || cols[c] >= 0x7fff) ? -1 : cols[c]);
    F.loff = d_loff.as<uint64_t>();
    F.roff = d_roff;
    F.n = n;
    F.out = d_cells.as<OutCell>();
    F.errflags = d_err;
    size_t total = cells.size();
    int grid = (int)std::min<size_t>((total + 127) / 128, 148 * 16);
    fetch_kernel<<<grid, 128, 0, st>>>(F);
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cells.data(), d_cells.p, total * sizeof(OutCell), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return CQG_OK;
}

// ------------------------------------------------------------------------------------------
// aggregate finishing: compact entries -> host values
// ------------------------------------------------------------------------------------------
struct GroupTable {
    DevBuf tab;
    uint64_t cap = 0;
    bool dense = false;  // `tab` is a dense array of `cap` occupied entries (no empty slots, not probeable)
    DevBuf packed;       // lean GROUP BY, global mode: the packed table the scan updated
};

static int alloc_group_table(HostPlan& hp, GroupTable& gt, uint64_t cap, cudaStream_t st) {
    DevPlan& P = hp.P;
    gt.cap = cap;
    gt.dense = false;
    CU(gt.tab.alloc(cap * (uint64_t)P.entry_bytes, st));
    int grid = (int)std::min<uint64_t>((cap * (uint64_t)(P.entry_bytes / 8) + 255) / 256, 148 * 8);
    init_table_kernel<<<grid, 256, 0, st>>>(gt.tab.as<uint8_t>(), cap, P.entry_bytes, P.entry_init);
    g_launches++;
    CU(cudaGetLastError());
    P.gtab = gt.tab.as<uint8_t>();
    P.gcap = cap;
    CU(cudaMemsetAsync(P.gcount, 0, 8, st));
    return CQG_OK;
}

static uint64_t initial_group_cap(const DevPlan& P) {
    if (P.ngc == 0) return 16;
    uint64_t bytes = P.own_hi - P.own_lo;
    uint64_t cap = 1 << 12;
    while (cap < bytes / 32 && cap < (1ull << 22)) cap <<= 1;
    return cap;
}

// device block -> pageable host memory through two page-locked bounce buffers: the copy of chunk k+1 runs while
// host threads move chunk k to its place (first touch of fresh pages is what costs on the host side)
static int copy_block_to_host(void* dst, const void* d_src, size_t n, cudaStream_t st) {
    constexpr size_t kChunk = 32u << 20;
    static void* bounce[2] = {nullptr, nullptr};
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (int k = 0; k < 2; k++)
        if (!bounce[k] && cudaMallocHost(&bounce[k], kChunk) != cudaSuccess) {
            cudaGetLastError();
            bounce[k] = nullptr;
        }
    if (!bounce[0] || !bounce[1]) {
        CU(cudaMemcpyAsync(dst, d_src, n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return CQG_OK;
    }
    cudaEvent_t ev[2];
    CU(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    const size_t chunks = (n + kChunk - 1) / kChunk;
    auto issue = [&](size_t k) {
        const size_t lo = k * kChunk, len = std::min(kChunk, n - lo);
        cudaMemcpyAsync(bounce[k & 1], (const char*)d_src + lo, len, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(ev[k & 1], st);
    };
    if (chunks) issue(0);
    int rc = CQG_OK;
    for (size_t k = 0; k < chunks; k++) {
        if (cudaEventSynchronize(ev[k & 1]) != cudaSuccess) {
            rc = fail(CQG_ERR_CUDA, "result copy: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (k + 1 < chunks) issue(k + 1);  // (the other bounce buffer: its previous contents were moved in the last round)
        const size_t lo = k * kChunk, len = std::min(kChunk, n - lo);
        const char* src = (const char*)bounce[k & 1];
        par_chunks(len, par_chunk_count(len), [&](unsigned, size_t a, size_t b) { memcpy((char*)dst + lo + a, src + a, b - a); });
    }
    cudaStreamSynchronize(st);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return rc;
}

// finish of a result with many groups, on the device: order by first appearance (radix sort), aggregate values,
// MIN/MAX rows and first-row columns decoded by one kernel, cqg_value_t records and strings written in their
// final layout, one block to the host. Same results as the host path below (same IEEE operations for SUM/AVG).
static int finish_aggregate_device(HostPlan& hp, const cqg_table* t, const cqg_query_t* q, const uint8_t* d_entries, uint64_t G,
                                   int64_t rows_scanned, cqg_result_t** out, cudaStream_t st) {
    DevPlan& P = hp.P;
    PhaseTimer pt;
    const int eb = P.entry_bytes;
    const int A = q->n_aggs, O = q->n_out_cols;
    const uint64_t ncell = G * (uint64_t)(A + O);
    DevBuf d_keys, d_keys2, d_idx, d_idx2, d_tmp;
    CU(d_keys.alloc(G * 8, st));
    CU(d_keys2.alloc(G * 8, st));
    CU(d_idx.alloc(G * 4, st));
    CU(d_idx2.alloc(G * 4, st));
    int grid = (int)std::min<uint64_t>((G + 255) / 256, 148 * 8);
    extract_first_kernel<<<grid, 256, 0, st>>>(d_entries, G, eb, d_keys.as<uint64_t>(), d_idx.as<uint32_t>());
    g_launches++;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<uint64_t>(), d_keys2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                    d_idx2.as<uint32_t>(), (int)G, 16, 64, st);
    CU(d_tmp.alloc(tmp_bytes, st));
    cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<uint64_t>(), d_keys2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                    d_idx2.as<uint32_t>(), (int)G, 16, 64, st);
    g_launches += 4;
    CU(cudaGetLastError());
    // the host block and its device image: first_offset | count | sum | ncount | value | out | strings
    const size_t gn = (size_t)G, an = (size_t)std::max(A, 1), on = (size_t)std::max(O, 1);
    const size_t o_first = 0, o_count = o_first + 8 * gn, o_sum = o_count + 8 * gn, o_ncount = o_sum + 8 * gn * an,
                 o_value = o_ncount + 8 * gn * an, o_out = o_value + 24 * gn * an, o_str = o_out + 24 * gn * on;
    DevBuf d_cells, d_ssize, d_soff, d_scan_tmp, d_block;
    CU(d_cells.alloc(std::max<uint64_t>(ncell, 1) * sizeof(OutCell), st));
    CU(d_ssize.alloc((ncell + 1) * 8, st));
    CU(d_soff.alloc((ncell + 1) * 8, st));
    CU(d_block.alloc(o_str + 8, st));
    if (A == 0) CU(cudaMemsetAsync((char*)d_block.p + o_sum, 0, o_out - o_sum, st));  // (placeholder arrays of a query without aggregates)
    if (O == 0) CU(cudaMemsetAsync((char*)d_block.p + o_out, 0, o_str - o_out, st));
    FinishParams F{};
    F.entries = d_entries;
    F.idx = d_idx2.as<uint32_t>();
    F.G = G;
    F.entry_bytes = eb;
    F.n_aggs = A;
    F.n_out = O;
    for (int a = 0; a < A; a++) {
        F.agg_func[a] = q->aggs[a].func;
        F.agg_col[a] = P.aggs[a].col;
        F.agg_off[a] = P.aggs[a].off;
    }
    for (int c = 0; c < O; c++) F.out_cols[c] = (int16_t)((q->out_cols[c] < 0 || q->out_cols[c] >= 0x7fff) ? -1 : q->out_cols[c]);
    F.data = t->d_data;
    F.size = t->size;
    F.global_base = P.global_base;
    F.delim = (uint8_t)t->cfg.delimiter;
    F.quote = (uint8_t)t->cfg.quote;
    F.first_offset = (uint64_t*)((char*)d_block.p + o_first);
    F.count = (int64_t*)((char*)d_block.p + o_count);
    F.sum = (double*)((char*)d_block.p + o_sum);
    F.ncount = (int64_t*)((char*)d_block.p + o_ncount);
    F.cells = d_cells.as<OutCell>();
    F.str_size = d_ssize.as<uint64_t>();
    F.errflags = P.errflags;
    if (ncell) {
        grid = (int)std::min<uint64_t>((ncell + 127) / 128, 148 * 16);
        finish_cells_kernel<<<grid, 128, 0, st>>>(F);
        g_launches++;
        CU(cudaGetLastError());
    } else {
        return fail(CQG_ERR_ARG, "query without output columns");  // (the caller routes those to the host path)
    }
    CU(cudaMemsetAsync((char*)d_ssize.p + ncell * 8, 0, 8, st));
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_ssize.as<uint64_t>(), d_soff.as<uint64_t>(), (int64_t)(ncell + 1), st);
    CU(d_scan_tmp.alloc(scan_bytes, st));
    cub::DeviceScan::ExclusiveSum(d_scan_tmp.p, scan_bytes, d_ssize.as<uint64_t>(), d_soff.as<uint64_t>(), (int64_t)(ncell + 1), st);
    g_launches += 2;
    CU(cudaGetLastError());
    uint64_t str_total = 0;
    CU(cudaMemcpyAsync(&str_total, d_soff.as<uint64_t>() + ncell, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    pt.lap("finish: sort, cells, string sizes");
    const size_t block_bytes = o_str + (size_t)str_total;
    // 2 MB aligned and advised for huge pages, or the block of an earlier result of this size (big_block_get)
    size_t block_cap = 0;
    char* block = (char*)big_block_get(block_bytes + 16, &block_cap);
    if (!block) return fail(CQG_ERR_NOMEM, "result of %llu groups: out of host memory", (unsigned long long)G);
    DevBuf d_full;  // the arrays written so far and the strings in one piece
    CU(d_full.alloc(block_bytes + 16, st));
    CU(cudaMemcpyAsync(d_full.p, d_block.p, o_value, cudaMemcpyDeviceToDevice, st));
    {
        grid = (int)std::min<uint64_t>((ncell + 127) / 128, 148 * 16);
        // cells are [aggregate or column][group]: exactly the order of `value` followed by `out` (when A == 0 the
        // placeholder value array of one column lies between: cells then start at `out`)
        uint8_t* values = (uint8_t*)d_full.p + (A > 0 ? o_value : o_out);
        if (A > 0 && O == 0) CU(cudaMemsetAsync((char*)d_full.p + o_out, 0, o_str - o_out, st));
        if (A == 0) CU(cudaMemsetAsync((char*)d_full.p + o_value, 0, o_out - o_value, st));
        finish_values_kernel<<<grid, 128, 0, st>>>(d_cells.as<OutCell>(), d_soff.as<uint64_t>(), ncell, t->d_data, values,
                                                   (uint8_t*)d_full.p + o_str, (uint64_t)(uintptr_t)(block + o_str));
        g_launches++;
        CU(cudaGetLastError());
    }
    int rc = copy_block_to_host(block, d_full.p, block_bytes, st);
    if (rc != CQG_OK) {
        big_block_put(block, block_cap);
        return rc;
    }
    pt.lap("finish: values, block to host");
    Arena* arena;
    cqg_result_t* r = new_result(&arena);
    arena->big = block;
    arena->big_cap = block_cap;
    r->n_groups = (int64_t)G;
    r->n_aggs = A;
    r->n_out_cols = O;
    r->rows_scanned = rows_scanned;
    r->first_offset = (uint64_t*)(block + o_first);
    r->count = (int64_t*)(block + o_count);
    r->sum = (double*)(block + o_sum);
    r->ncount = (int64_t*)(block + o_ncount);
    r->value = (cqg_value_t*)(block + o_value);
    r->out = (cqg_value_t*)(block + o_out);
    *out = r;
    return CQG_OK;
}

static int finish_aggregate(HostPlan& hp, const cqg_table* t, const cqg_table* rt, const cqg_query_t* q, const uint8_t* d_entries,
                            uint64_t G, bool entries_on_device, int64_t rows_scanned, cqg_result_t** out, cudaStream_t st) {
    DevPlan& P = hp.P;
    // many groups, no join: the whole finish runs on the device (CQG_FINISH=host keeps the path below for A/B runs)
    if (G > 4096 && entries_on_device && !rt && !(q->n_group_cols == 1 && q->group_cols[0] < 0) && q->n_aggs + q->n_out_cols > 0) {
        const char* fe = getenv("CQG_FINISH");
        if (!(fe && fe[0] == 'h')) return finish_aggregate_device(hp, t, q, d_entries, G, rows_scanned, out, st);
    }
    PhaseTimer pt;
    const int eb = P.entry_bytes;
    std::vector<uint8_t> ent;
    const uint8_t* entp = nullptr;
    bool presorted = false;
    if (G && entries_on_device && G > 4096) {
        // many groups: order them by first appearance on the device (radix sort of the first-offset keys,
        // then a gather), and bring them over through a page-locked staging buffer
        DevBuf d_keys, d_keys2, d_idx, d_idx2, d_tmp, d_sorted;
        CU(d_keys.alloc(G * 8, st));
        CU(d_keys2.alloc(G * 8, st));
        CU(d_idx.alloc(G * 4, st));
        CU(d_idx2.alloc(G * 4, st));
        int grid = (int)std::min<uint64_t>((G + 255) / 256, 148 * 8);
        extract_first_kernel<<<grid, 256, 0, st>>>(d_entries, G, eb, d_keys.as<uint64_t>(), d_idx.as<uint32_t>());
        g_launches++;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<uint64_t>(), d_keys2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                        d_idx2.as<uint32_t>(), (int)G, 0, 64, st);
        CU(d_tmp.alloc(tmp_bytes, st));
        cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<uint64_t>(), d_keys2.as<uint64_t>(), d_idx.as<uint32_t>(),
                                        d_idx2.as<uint32_t>(), (int)G, 0, 64, st);
        g_launches += 4;
        CU(d_sorted.alloc(G * (uint64_t)eb, st));
        int g2 = (int)std::min<uint64_t>((G * (uint64_t)(eb / 16) + 255) / 256, 148 * 16);
        gather_entries_kernel<<<g2, 256, 0, st>>>(d_entries, d_idx2.as<uint32_t>(), G, eb, d_sorted.as<uint8_t>());
        g_launches++;
        CU(cudaGetLastError());
        static void* pinned = nullptr;
        static size_t pinned_bytes = 0;
        size_t need = (size_t)G * eb;
        if (need > pinned_bytes) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr;
            pinned_bytes = 0;
            if (cudaMallocHost(&pinned, need) == cudaSuccess) pinned_bytes = need;
        }
        if (pinned_bytes >= need) {
            CU(cudaMemcpyAsync(pinned, d_sorted.p, need, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            entp = (const uint8_t*)pinned;  // read in place (valid until the next finish on this process)
        } else {
            ent.resize(need + 8);
            CU(cudaMemcpyAsync(ent.data(), d_sorted.p, need, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            entp = ent.data();
        }
        presorted = true;
    } else {
        ent.resize((size_t)G * eb + 8);
        if (G) {
            if (entries_on_device) {
                CU(cudaMemcpyAsync(ent.data(), d_entries, (size_t)G * eb, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
            } else {
                memcpy(ent.data(), d_entries, (size_t)G * eb);
            }
        }
        entp = ent.data();
    }
    // create_groups with an unknown single key column makes no group at all (aggregates.c:114-116)
    bool zero_groups = q->n_group_cols == 1 && q->group_cols[0] < 0;
    if (zero_groups) G = 0;
    // no GROUP BY: one `_all_` group even over zero rows (src/evaluator.c:232-247)
    bool synth = false;
    if (q->n_group_cols == 0 && G == 0) {
        ent.assign(hp.entry_init.begin(), hp.entry_init.end());
        entp = ent.data();
        G = 1;
        synth = true;
    }
    pt.lap("entries to host");
    std::vector<uint32_t> order((size_t)G);
    std::iota(order.begin(), order.end(), 0u);
    auto first_of = [&](uint32_t i) { return *(const uint64_t*)(entp + (size_t)i * eb + kOffFirst); };
    if (!presorted) std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return first_of(a) < first_of(b); });
    pt.lap("sort by first appearance");

    Arena* arena;
    cqg_result_t* r = new_result(&arena);
    r->n_groups = (int64_t)G;
    r->n_aggs = q->n_aggs;
    r->n_out_cols = q->n_out_cols;
    r->rows_scanned = rows_scanned;
    size_t gn = std::max<size_t>(G, 1), an = (size_t)std::max(q->n_aggs, 1), on = (size_t)std::max(q->n_out_cols, 1);
    r->first_offset = (uint64_t*)arena->alloc(8 * gn);
    r->count = (int64_t*)arena->alloc(8 * gn);
    r->sum = (double*)arena->alloc(8 * gn * an);
    r->ncount = (int64_t*)arena->alloc(8 * gn * an);
    r->value = (cqg_value_t*)arena->alloc(sizeof(cqg_value_t) * gn * an);
    r->out = (cqg_value_t*)arena->alloc(sizeof(cqg_value_t) * gn * on);

    std::vector<uint64_t> loff(G);
    std::vector<std::pair<uint64_t, uint64_t>> numfetch[CQG_MAX_AGGS];  // (group, okey) of numeric MIN/MAX results
    std::vector<OutCell> strcells;  // MIN/MAX string results
    std::vector<size_t> strdst;
    unsigned host_flags = 0;
    const unsigned TF = par_chunk_count((size_t)G);
    struct FinChunk {
        std::vector<std::pair<uint64_t, uint64_t>> numfetch[CQG_MAX_AGGS];
        std::vector<OutCell> strcells;
        std::vector<size_t> strdst;
        unsigned host_flags = 0;
    };
    std::vector<FinChunk> fch(TF);
    par_chunks((size_t)G, TF, [&](unsigned chunk_k, size_t gi_lo, size_t gi_hi) {
    auto& numfetch = fch[chunk_k].numfetch;
    auto& strcells = fch[chunk_k].strcells;
    auto& strdst = fch[chunk_k].strdst;
    unsigned& host_flags = fch[chunk_k].host_flags;
    (void)host_flags;
    for (uint64_t gi = gi_lo; gi < gi_hi; gi++) {
        const uint8_t* e = entp + (size_t)order[gi] * eb;
        uint64_t first = *(const uint64_t*)(e + kOffFirst);
        int64_t count = (int64_t) * (const uint64_t*)(e + kOffCount);
        loff[gi] = first == ~0ull ? ~0ull : (first >> 16) - P.global_base;
        r->first_offset[gi] = first == ~0ull ? 0 : (first >> 16);
        r->count[gi] = count;
        for (int a = 0; a < q->n_aggs; a++) {
            size_t ix = (size_t)a * G + gi;
            const AggSpec& sp = P.aggs[a];
            cqg_value_t v;
            memset(&v, 0, sizeof v);
            int f = q->aggs[a].func;
            if (f == CQG_AGG_COUNT_STAR) {
                v.type = CQG_TYPE_INTEGER;
                v.int_value = (int)count;  // `result.int_value = row_count` with an int row_count (:270)
            } else if (sp.col < 0) {
                v.type = CQG_TYPE_NULL;
            } else if (f == CQG_AGG_COUNT) {
                v.type = CQG_TYPE_INTEGER;
                v.int_value = (int)count;
            } else if (f == CQG_AGG_SUM || f == CQG_AGG_AVG) {
                long long si = *(const long long*)(e + sp.off);
                double sd = *(const double*)(e + sp.off + 8);
                int64_t n = (int64_t) * (const uint64_t*)(e + sp.off + 16);
                long long s3 = *(const long long*)(e + sp.off + 24);
                double sum = (double)si + sd + (double)s3 / 1000.0;
                r->sum[ix] = sum;
                r->ncount[ix] = n;
                v.type = CQG_TYPE_DOUBLE;
                // `count` is an int in the reference (:288); AVG divides by it
                v.double_value = f == CQG_AGG_SUM ? sum : ((int)n > 0 ? sum / (double)(int)n : 0.0);
            } else {
                const uint64_t* s = (const uint64_t*)(e + sp.off);
                uint32_t cls = s[0] == ~0ull ? 0u : (uint32_t)(s[0] & 3u);
                if (cls == 1) {
                    // type and bits of the extreme come from the row that holds it
                    numfetch[a].push_back({(uint64_t)gi, s[3]});
                } else if (cls == 3) {
                    uint64_t d = s[1] - 1ull;
                    v.type = CQG_TYPE_DATE;
                    v.date_value.year = (int)(d >> 16);
                    v.date_value.month = (int)((d >> 8) & 0xff);
                    v.date_value.day = (int)(d & 0xff);
                } else if (cls == 2) {
                    OutCell c;
                    c.type = CQG_TYPE_STRING;
                    c.len = (uint32_t)(s[4] & 0x3ffffu);
                    c.payload = (s[4] & (1ull << 63)) | ((s[4] >> 18) & 0x1fffffffffffull);
                    strcells.push_back(c);
                    strdst.push_back(ix);
                    v.type = CQG_TYPE_STRING;
                }
            }
            r->value[ix] = v;
        }
    }
    });
    for (unsigned k = 0; k < TF; k++) {
        for (int a = 0; a < CQG_MAX_AGGS; a++) numfetch[a].insert(numfetch[a].end(), fch[k].numfetch[a].begin(), fch[k].numfetch[a].end());
        strcells.insert(strcells.end(), fch[k].strcells.begin(), fch[k].strcells.end());
        strdst.insert(strdst.end(), fch[k].strdst.begin(), fch[k].strdst.end());
        host_flags |= fch[k].host_flags;
    }
    if (host_flags) {
        cqg_result_free(r);
        return fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(host_flags));
    }
    pt.lap("aggregate values");
    int rc = CQG_OK;
    for (int a = 0; a < q->n_aggs && rc == CQG_OK; a++) {
        if (numfetch[a].empty()) continue;
        size_t n = numfetch[a].size();
        std::vector<uint64_t> lo(n), oks(n);
        for (size_t k = 0; k < n; k++) {
            oks[k] = numfetch[a][k].second;
            lo[k] = (oks[k] >> 16) - P.global_base;
        }
        DevBuf d_ok, d_ro;
        const uint64_t* roff_ptr = nullptr;
        int32_t col = P.aggs[a].col;
        if (rt && P.join && col >= P.n_left_cols) {
            CU(d_ok.alloc(n * 8, st));
            CU(d_ro.alloc(n * 8, st));
            CU(cudaMemcpyAsync(d_ok.p, oks.data(), n * 8, cudaMemcpyHostToDevice, st));
            int grid = (int)std::min<uint64_t>((n + 127) / 128, 148 * 8);
            resolve_first_right_kernel<<<grid, 128, 0, st>>>(P, d_ok.as<uint64_t>(), n, d_ro.as<uint64_t>());
            g_launches++;
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(st));
            roff_ptr = d_ro.as<uint64_t>();
        }
        std::vector<OutCell> cells;
        rc = run_fetch(t, rt, &col, 1, lo, roff_ptr, P.errflags, cells, st);
        if (rc != CQG_OK) break;
        std::vector<cqg_value_t> vals(n);
        rc = cells_to_values(t, rt, cells, vals.data(), arena, st);
        for (size_t k = 0; k < n && rc == CQG_OK; k++) r->value[(size_t)a * G + numfetch[a][k].first] = vals[k];
    }
    if (rc != CQG_OK) {
        cqg_result_free(r);
        return rc;
    }
    if (!strcells.empty()) {
        std::vector<cqg_value_t> sv(strcells.size());
        rc = cells_to_values(t, rt, strcells, sv.data(), arena, st);
        if (rc == CQG_OK)
            for (size_t k = 0; k < sv.size(); k++) r->value[strdst[k]] = sv[k];
    }
    pt.lap("min/max rows re-read");
    // bare columns: the group's first row (evaluator_aggregates.c:679-689)
    if (rc == CQG_OK && q->n_out_cols > 0 && G > 0 && !synth) {
        DevBuf d_first, d_roff;
        const uint64_t* roff_ptr = nullptr;
        if (rt && P.join) {
            std::vector<uint64_t> firsts(G);
            for (uint64_t gi = 0; gi < G; gi++) firsts[gi] = *(const uint64_t*)(entp + (size_t)order[gi] * eb + kOffFirst);
            CU(d_first.alloc(G * 8, st));
            CU(d_roff.alloc(G * 8, st));
            CU(cudaMemcpyAsync(d_first.p, firsts.data(), G * 8, cudaMemcpyHostToDevice, st));
            int grid = (int)std::min<uint64_t>((G + 127) / 128, 148 * 8);
            resolve_first_right_kernel<<<grid, 128, 0, st>>>(P, d_first.as<uint64_t>(), G, d_roff.as<uint64_t>());
            g_launches++;
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(st));
            roff_ptr = d_roff.as<uint64_t>();
        }
        std::vector<OutCell> cells;
        rc = run_fetch(t, rt, q->out_cols, q->n_out_cols, loff, roff_ptr, P.errflags, cells, st);
        if (rc == CQG_OK) {
            // cells are [group][col]; the result wants [col][group]
            std::vector<OutCell> tr(cells.size());
            par_chunks((size_t)G, par_chunk_count((size_t)G), [&](unsigned, size_t lo, size_t hi) {
                for (uint64_t gi = lo; gi < hi; gi++)
                    for (int c = 0; c < q->n_out_cols; c++) tr[(size_t)c * G + gi] = cells[(size_t)gi * q->n_out_cols + c];
            });
            rc = cells_to_values(t, rt, tr, r->out, arena, st);
        }
    }
    pt.lap("first-row columns");
    if (rc != CQG_OK) {
        cqg_result_free(r);
        return rc;
    }
    *out = r;
    return CQG_OK;
}

// ------------------------------------------------------------------------------------------
// execute
// ------------------------------------------------------------------------------------------
static int read_scalars(HostPlan& hp, ScalarBlock& hs, cudaStream_t st) {
    cudaError_t ce = cudaMemcpyAsync(&hs, hp.d_scalars.p, sizeof hs, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(CQG_ERR_CUDA, "scan kernel: %s", cudaGetErrorString(ce));
    return CQG_OK;
}

static void swap_buf(DevBuf& a, DevBuf& b) {
    std::swap(a.p, b.p);
    std::swap(a.n, b.n);
    std::swap(a.s, b.s);
}

// the packed table of the lean GROUP BY in global mode: `cap` empty lines
static int alloc_packed_table(HostPlan& hp, GroupTable& gt, uint64_t cap, cudaStream_t st) {
    DevPlan& P = hp.P;
    const int pb = P.pk.entry_bytes;
    if (!hp.d_packed_init.p) {
        CU(hp.d_packed_init.alloc(hp.packed_init.size(), st));
        CU(cudaMemcpyAsync(hp.d_packed_init.p, hp.packed_init.data(), hp.packed_init.size(), cudaMemcpyHostToDevice, st));
    }
    CU(gt.packed.alloc(cap * (uint64_t)pb, st));
    int grid = (int)std::min<uint64_t>((cap * (uint64_t)(pb / 8) + 255) / 256, 148 * 8);
    init_table_kernel<<<grid, 256, 0, st>>>(gt.packed.as<uint8_t>(), cap, pb, hp.d_packed_init.as<uint8_t>());
    g_launches++;
    CU(cudaGetLastError());
    P.ptab = gt.packed.as<uint8_t>();
    P.pcap = cap;
    return CQG_OK;
}

static int compact_groups(HostPlan& hp, GroupTable& gt, int owner, int world, DevBuf& out, uint64_t* n_out, cudaStream_t st,
                          long long known_occupied = -1, bool may_steal = false);

// DevPlan::simple plans: the lean kernel, then the general kernel on whatever it handed over.
// GROUP BY first numbers groups per CTA (shared-memory dictionary, <= 32/64 groups per CTA); when a CTA meets
// more, the lean kernel is rerun in GLOBAL mode: find-or-insert in a table of packed 64/128-byte lines
// (PackedLayout), the general kernel builds general entries for the handed-over tiles and rows in a table of
// its own, and the two are brought together as one dense array of general entries (gt.dense).
// *done = 0: the caller must run the general scan instead.
static int run_lean_scan(HostPlan& hp, GroupTable& gt, cudaStream_t st, ScalarBlock& hs, float* ms_out, cudaEvent_t e0,
                         cudaEvent_t e1, int* done) {
    DevPlan& P = hp.P;
    *done = 0;
    int rc;
    DevBuf d_tiles, d_rows;
    const uint64_t row_cap = (uint64_t)P.n_tiles * 8u + 1024u;
    CU(d_tiles.alloc((size_t)(P.n_tiles + 1) * 4, st));
    CU(d_rows.alloc(row_cap * 8, st));
    ScalarBlock* sb = hp.d_scalars.as<ScalarBlock>();
    P.def_tiles = d_tiles.as<int32_t>();
    P.def_rows = d_rows.as<uint64_t>();
    P.def_tile_count = &sb->def_tile_count;
    P.def_row_count = &sb->def_row_count;
    P.def_row_cap = row_cap;
    P.tile_list = nullptr;
    P.lean_global = (hp.start_global && P.pk.entry_bytes && env_int("CQG_START_GLOBAL", 1)) ? 1 : 0;
    P.lean_k = 0;
    if (!P.lean_global && hp.few_text_keys && P.ngc == 1 && P.l_nagg <= 3 && env_int("CQG_LEAN2K", 1)) {
        bool fits = true;
        for (int a = 0; a < P.l_nagg; a++) {
            const int f = P.aggs[P.l_agg[a]].func;
            fits = fits && (f == CQG_AGG_SUM || f == CQG_AGG_AVG);
        }
        for (int c = 0; c < P.l_nleaf; c++) fits = fits && P.l_leaf[c].kind == 0;
        P.lean_k = fits ? 1 : 0;
    }
    P.ptab = nullptr;
    P.pcap = 0;
    P.hc_debug = env_int("CQG_HC_DEBUG", 0);
    // big scans: the two tiles at the file's edges stay on the lean kernel (no second launch, no second round trip); small
    // files keep the general kernel for them (every tile of a file below 16 KB is an edge tile)
    P.edge_in_kernel = (P.n_tiles >= env_int("CQG_EDGE_MIN_TILES", 64)) ? 1 : 0;
    {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        P.dec_table = decimal_table(dev);
        if (!P.dec_table) return fail(CQG_ERR_CUDA, "decimal table allocation failed");
    }
    uint64_t cap = P.ngc == 0 ? 16 : P.lean_global ? initial_group_cap(P) : (1u << 14);
    PhaseTimer pt;
    for (int attempt = 0; attempt < 14; attempt++) {
        if (P.lean_global) {
            if ((rc = alloc_packed_table(hp, gt, cap, st))) return rc;
        } else {
            if ((rc = alloc_group_table(hp, gt, cap, st))) return rc;
        }
        CU(cudaMemsetAsync(sb, 0, sizeof(ScalarBlock), st));
        pt.lap("lean: buffers");
        cudaEventRecord(e0, st);
        if ((rc = launch_lean(P, st))) return rc;
        cudaEventRecord(e1, st);
        if ((rc = read_scalars(hp, hs, st))) return rc;
        pt.lap("lean: kernel + flags");
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out += ms;
        if (P.lean_k && (hs.errflags & (KERR_LEAN_GROUPS | KERR_LEAN_ABORT))) {
            P.lean_k = 0;  // more than 16 groups in a CTA, or rows lean2k_kernel does not take: lean2g_kernel next
            continue;
        }
        if ((hs.errflags & KERR_LEAN_GROUPS) && !P.lean_global && P.pk.entry_bytes) {
            P.lean_global = 1;  // too many groups for per-CTA numbering
            cap = initial_group_cap(P);
            continue;
        }
        if (hs.errflags & KERR_TABLE_FULL) {
            cap *= 4;
            if (cap > (1ull << 32)) return fail(CQG_ERR_NOMEM, "group table beyond 2^32 entries");
            continue;
        }
        break;
    }
    if ((hs.errflags & (KERR_LEAN_ABORT | KERR_TABLE_FULL)) || hs.def_row_count > row_cap) {
        P.simple = 0;  // the data is not what the lean kernel is for
        g_last_scan[0] = g_last_scan[1] = P.n_tiles;  // (everything goes to the general kernel)
        g_last_scan[2] = 0;
        P.lean_global = 0;
        P.ptab = nullptr;
        P.pcap = 0;
        gt.packed.release();
        return CQG_OK;
    }
    g_last_scan[0] = P.n_tiles;
    g_last_scan[1] = (int64_t)hs.def_tile_count;
    g_last_scan[2] = (int64_t)hs.def_row_count;
    const unsigned long long rows_after_lean = hs.rows_scanned;
    ScalarBlock h2{};
    if (hs.def_tile_count || hs.def_row_count) {
        uint64_t st_cap = 1u << 12;
        for (int attempt = 0; attempt < 12; attempt++) {
            if (P.lean_global) {
                // the general kernel's own table; sized by what was handed over, grown on overflow
                if ((rc = alloc_group_table(hp, gt, st_cap, st))) return rc;
                CU(cudaMemcpyAsync(&sb->rows_scanned, &rows_after_lean, 8, cudaMemcpyHostToDevice, st));
            }
            cudaEventRecord(e0, st);
            CU(cudaMemsetAsync(&sb->errflags, 0, 4, st));
            if (hs.def_tile_count) {
                DevPlan T = P;
                T.tile_list = P.def_tiles;
                T.first_tile = 0;
                T.n_tiles = (int32_t)hs.def_tile_count;
                if ((rc = launch_scan(T, hp.table_smem_bytes, st))) return rc;
            }
            if (hs.def_row_count) {
                DevPlan R = P;
                R.simple = 0;
                R.scalar_regs = 0;
                R.smem_cap = 0;
                int grid = (int)std::min<uint64_t>((hs.def_row_count + 127) / 128, 148 * 8);
                deferred_rows_kernel<<<grid, 128, 0, st>>>(R, P.def_rows, hs.def_row_count);
                g_launches++;
                CU(cudaGetLastError());
            }
            cudaEventRecord(e1, st);
            if ((rc = read_scalars(hp, h2, st))) return rc;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            *ms_out += ms;
            pt.lap("lean: handed-over tiles and rows");
            if (h2.errflags & KERR_TABLE_FULL) {
                if (P.lean_global) {
                    st_cap *= 4;
                    if (st_cap > (1ull << 32)) return fail(CQG_ERR_NOMEM, "group table beyond 2^32 entries");
                    continue;
                }
                // the handed-over part overflowed the table the lean pass sized: start over on the general kernel
                P.simple = 0;
                P.lean_global = 0;
                return CQG_OK;
            }
            hs.errflags |= h2.errflags & ~KERR_TABLE_FULL;
            hs.rows_scanned = h2.rows_scanned;
            hs.gcount = h2.gcount;
            break;
        }
    }
    if (P.lean_global) {
        // packed lines -> general entries, then the general kernel's entries folded in
        const uint64_t g_pt = hs.pcount, g_st = (hs.def_tile_count || hs.def_row_count) ? h2.gcount : 0;
        const uint64_t dense_cap = g_pt + g_st + 1;
        DevBuf dense, strecs;
        CU(dense.alloc(dense_cap * (uint64_t)P.entry_bytes, st));
        cudaEventRecord(e0, st);
        CU(cudaMemsetAsync(&sb->n_dense, 0, 8, st));
        {
            int grid = (int)std::min<uint64_t>((P.pcap + 255) / 256, 148 * 16);
            expand_packed_kernel<<<grid, 256, 0, st>>>(P, dense.as<uint8_t>(), dense_cap, &sb->n_dense);
            g_launches++;
            CU(cudaGetLastError());
        }
        if (g_st) {
            uint64_t n = 0;
            if ((rc = compact_groups(hp, gt, 0, 1, strecs, &n, st, (long long)g_st))) return rc;
            int grid = (int)std::min<uint64_t>((n + 127) / 128, 148 * 8);
            merge_general_into_dense_kernel<<<grid, 128, 0, st>>>(P, strecs.as<uint8_t>(), n, dense.as<uint8_t>(), dense_cap, &sb->n_dense);
            g_launches++;
            CU(cudaGetLastError());
        }
        cudaEventRecord(e1, st);
        ScalarBlock h3{};
        if ((rc = read_scalars(hp, h3, st))) return rc;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out += ms;
        if (h3.n_dense > dense_cap) return fail(CQG_ERR_CUDA, "expanded group entries overflowed their buffer");
        hs.errflags |= h3.errflags & ~KERR_TABLE_FULL;
        swap_buf(gt.tab, dense);
        gt.cap = h3.n_dense;
        gt.dense = true;
        gt.packed.release();
        P.ptab = nullptr;
        P.pcap = 0;
        P.gtab = gt.tab.as<uint8_t>();
        P.gcap = gt.cap;
        hs.gcount = h3.n_dense;
        CU(cudaMemcpyAsync(P.gcount, &hs.gcount, 8, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));  // (&hs.gcount is caller memory: the copy has left it when this returns)
        pt.lap("lean: packed lines -> entries");
    }
    *done = 1;
    return CQG_OK;
}

// RIGHT / FULL joins: after the probe scan, the right rows no left row matched (general operators, global table)
static int launch_unmatched_right(const DevPlan& P, cudaStream_t st) {
    if (!P.join || P.join_type < CQG_JOIN_RIGHT || !P.jcap) return CQG_OK;
    DevPlan R = P;
    R.simple = 0;
    R.scalar_regs = 0;
    R.smem_cap = 0;
    int grid = (int)std::min<uint64_t>((P.jcap + 127) / 128, 148 * 8);
    join_unmatched_right_kernel<<<grid, 128, 0, st>>>(R);
    g_launches++;
    CU(cudaGetLastError());
    return CQG_OK;
}

static int run_aggregate_scan(HostPlan& hp, GroupTable& gt, cudaStream_t st, ScalarBlock& hs, float* ms_out) {
    DevPlan& P = hp.P;
    uint64_t cap = initial_group_cap(P);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    int rc = CQG_OK;
    if (P.simple) {
        int done = 0;
        rc = run_lean_scan(hp, gt, st, hs, ms_out, e0, e1, &done);
        if (rc != CQG_OK || done) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return rc;
        }
    }
    for (int attempt = 0; attempt < 12; attempt++) {
        if ((rc = alloc_group_table(hp, gt, cap, st))) break;
        cudaMemsetAsync(hp.d_scalars.p, 0, sizeof(ScalarBlock), st);
        cudaEventRecord(e0, st);
        if ((rc = launch_scan(P, hp.table_smem_bytes, st))) break;
        if ((rc = launch_unmatched_right(P, st))) break;
        cudaEventRecord(e1, st);
        if ((rc = read_scalars(hp, hs, st))) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_out += ms;
        if (hs.errflags & KERR_TABLE_FULL) {
            cap *= 4;
            if (cap > (1ull << 32)) {
                rc = fail(CQG_ERR_NOMEM, "group table beyond 2^32 entries");
                break;
            }
            continue;
        }
        break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

static int check_flags(const ScalarBlock& hs, bool join) {
    unsigned f = hs.errflags & kFatalMask;
    if (join) {
        unsigned classes = (hs.jclass[0] | hs.jclass[1]) & ~1u;  // bit 0 = NULL keys
        if (classes & (classes - 1)) f |= KERR_JOIN_MIXED;
    }
    if (f) return fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(f));
    return CQG_OK;
}

// known_occupied >= 0: the table's entry count as the scan's scalar block already reported it (saves a round trip).
// A dense table (gt.dense) already is the list of its entries: it is handed over (may_steal) or copied.
static int compact_groups(HostPlan& hp, GroupTable& gt, int owner, int world, DevBuf& out, uint64_t* n_out, cudaStream_t st,
                          long long known_occupied, bool may_steal) {
    DevPlan& P = hp.P;
    if (gt.dense && world <= 1) {
        *n_out = gt.cap;
        if (may_steal) {
            swap_buf(out, gt.tab);
            gt.cap = 0;
        } else {
            CU(out.alloc((gt.cap + 1) * (uint64_t)P.entry_bytes, st));
            CU(cudaMemcpyAsync(out.p, gt.tab.p, gt.cap * (uint64_t)P.entry_bytes, cudaMemcpyDeviceToDevice, st));
        }
        return CQG_OK;
    }
    unsigned long long occupied = (unsigned long long)known_occupied;
    if (known_occupied < 0) {
        CU(cudaMemcpyAsync(&occupied, P.gcount, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    CU(out.alloc((occupied + 1) * (uint64_t)P.entry_bytes, st));
    DevBuf cnt;
    CU(cnt.alloc(8, st));
    CU(cudaMemsetAsync(cnt.p, 0, 8, st));
    int grid = (int)std::min<uint64_t>((gt.cap + 255) / 256, 148 * 8);
    if (gt.cap)
        compact_table_kernel<<<grid, 256, 0, st>>>(gt.tab.as<uint8_t>(), gt.cap, P.entry_bytes, out.as<uint8_t>(), occupied + 1,
                                                   cnt.as<unsigned long long>(), owner, world);
    g_launches++;
    CU(cudaGetLastError());
    if (world <= 1) {  // no owner filter: every occupied entry is written
        *n_out = occupied;
        return CQG_OK;
    }
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = std::min<uint64_t>(n, occupied + 1);
    return CQG_OK;
}

static int execute_select(HostPlan& hp, const cqg_table* t, const cqg_table* rt, const cqg_query_t* q, cqg_result_t** out,
                          cudaStream_t st, float* ms_total, const JoinState* js = nullptr) {
    DevPlan& P = hp.P;
    uint64_t cap = 1 << 16;
    ScalarBlock hs{};
    DevBuf okeys, roffs;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    int rc = CQG_OK;
    for (int attempt = 0; attempt < 4; attempt++) {
        CU(okeys.alloc(cap * 8, st));
        if (rt) CU(roffs.alloc(cap * 8, st));
        P.sel_okey = okeys.as<uint64_t>();
        P.sel_roff = rt ? roffs.as<uint64_t>() : nullptr;
        P.sel_cap = cap;
        cudaMemsetAsync(P.errflags, 0, 4, st);
        cudaMemsetAsync(P.rows_scanned, 0, 8, st);
        cudaMemsetAsync(P.sel_count, 0, 8, st);
        cudaEventRecord(e0, st);
        if ((rc = launch_scan(P, 0, st))) break;
        if ((rc = launch_unmatched_right(P, st))) break;
        cudaEventRecord(e1, st);
        cudaError_t ce = cudaMemcpyAsync(&hs, hp.d_scalars.p, sizeof hs, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
            rc = fail(CQG_ERR_CUDA, "scan kernel: %s", cudaGetErrorString(ce));
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_total += ms;
        if (hs.sel_count > cap) {
            cap = hs.sel_count + 16;
            continue;
        }
        break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (js) {  // what building the join table over the right file raised (a key beyond the decoders' range, ...)
        hs.errflags |= js->flags;
        hs.jclass[1] |= js->key_classes;
    }
    if ((rc = check_flags(hs, P.join != 0))) return rc;
    uint64_t n = hs.sel_count;
    std::vector<uint64_t> ok(n), ro(rt ? n : 0);
    if (n) {
        CU(cudaMemcpyAsync(ok.data(), okeys.p, n * 8, cudaMemcpyDeviceToHost, st));
        if (rt) CU(cudaMemcpyAsync(ro.data(), roffs.p, n * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ok[a] < ok[b]; });
    uint64_t nout = n;
    if (q->max_rows >= 0 && (uint64_t)q->max_rows < nout) nout = (uint64_t)q->max_rows;
    std::vector<uint64_t> loff(nout), roff_sorted(rt ? nout : 0);
    for (uint64_t i = 0; i < nout; i++) {
        loff[i] = (ok[order[i]] >> 16) - P.global_base;
        if (rt) roff_sorted[i] = ro[order[i]];
    }
    Arena* arena;
    cqg_result_t* r = new_result(&arena);
    r->n_selected = (int64_t)n;
    r->n_rows_out = (int64_t)nout;
    r->rows_scanned = (int64_t)hs.rows_scanned;
    r->n_aggs = 0;
    r->n_out_cols = q->n_out_cols;
    size_t rn = std::max<size_t>(nout, 1), on = (size_t)std::max(q->n_out_cols, 1);
    r->row_offset = (uint64_t*)arena->alloc(8 * rn);
    r->row_offset_right = rt ? (uint64_t*)arena->alloc(8 * rn) : nullptr;
    r->rows = (cqg_value_t*)arena->alloc(sizeof(cqg_value_t) * rn * on);
    for (uint64_t i = 0; i < nout; i++) {
        r->row_offset[i] = loff[i] + P.global_base;
        if (rt) r->row_offset_right[i] = roff_sorted[i];
    }
    DevBuf d_roff;
    const uint64_t* roff_ptr = nullptr;
    if (rt && nout) {
        CU(d_roff.alloc(nout * 8, st));
        CU(cudaMemcpyAsync(d_roff.p, roff_sorted.data(), nout * 8, cudaMemcpyHostToDevice, st));
        roff_ptr = d_roff.as<uint64_t>();
    }
    std::vector<OutCell> cells;
    rc = run_fetch(t, rt, q->out_cols, q->n_out_cols, loff, roff_ptr, P.errflags, cells, st);
    if (rc == CQG_OK) rc = cells_to_values(t, rt, cells, r->rows, arena, st);
    if (rc == CQG_OK) {
        unsigned f = 0;
        CU(cudaMemcpy(&f, P.errflags, 4, cudaMemcpyDeviceToHost));
        if (f & kFatalMask) rc = fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(f & kFatalMask));
    }
    if (rc != CQG_OK) {
        cqg_result_free(r);
        return rc;
    }
    *out = r;
    return CQG_OK;
}

static bool partial_query_ok(const cqg_query_t* q);

static int execute_multi(const cqg_table* t, const cqg_query_t* q, cqg_result_t** out);

CQG_API int cqg_execute(const cqg_table_t* t, const cqg_query_t* q, cqg_result_t** out) {
    if (!t || !q || !out) return fail(CQG_ERR_ARG, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t st = 0;
    if (t->ngpu > 1 || (q->join.right && q->join.right->ngpu > 1)) {
        // a table spread over several devices: aggregates without a join run on all of them; every other shape runs on
        // the first device of the table over the one address range (remote pages come over NVLink)
        const cqg_table* mt = t->ngpu > 1 ? t : q->join.right;
        {
            HostPlan probe;  // (shape check from the header alone, as below: a declined shape costs no upload)
            if ((rc = build_plan(probe, t, q, st))) return rc == CQG_ERR_UNSUPPORTED ? CQG_ERR_UNSUPPORTED_PLAN : rc;
        }
        if ((rc = ensure_staged(t)) || (rc = ensure_staged(q->join.right))) return rc;
        if (t->ngpu > 1 && !q->join.right && partial_query_ok(q)) return execute_multi(t, q, out);
        CU(cudaSetDevice(mt->devices[0]));
    }
    HostPlan hp;
    PhaseTimer pt;
    // the plan is built from the header alone: a shape the planner declines costs no upload (the caller keeps its own
    // route for it: CQG_ERR_UNSUPPORTED_PLAN, as opposed to data the kernels met and could not reproduce)
    if ((rc = build_plan(hp, t, q, st))) return rc == CQG_ERR_UNSUPPORTED ? CQG_ERR_UNSUPPORTED_PLAN : rc;
    pt.lap("execute: plan");
    const cqg_table* rt = q->join.right;
    if ((rc = ensure_staged(t)) || (rc = ensure_staged(rt))) return rc;
    hp.P.data = t->d_data;
    if (rt) hp.P.rdata = rt->d_data;
    pt.lap("execute: tables resident");
    JoinState js;
    float ms = 0;
    long long launches0 = g_launches.load();
    if (rt && (rc = build_join(hp, js, rt, q->join.right_col, st))) return rc;
    if (q->mode == CQG_MODE_SELECT) {
        rc = execute_select(hp, t, rt, q, out, st, &ms, &js);
    } else {
        GroupTable gt;
        ScalarBlock hs{};
        if ((rc = run_aggregate_scan(hp, gt, st, hs, &ms))) return rc;
        pt.lap("execute: scan");
        hs.errflags |= js.flags;
        hs.jclass[1] |= js.key_classes;
        if ((rc = check_flags(hs, hp.P.join != 0))) return rc;
        DevBuf entries;
        uint64_t G = 0;
        if ((rc = compact_groups(hp, gt, 0, 1, entries, &G, st, (long long)hs.gcount, true))) return rc;
        pt.lap("execute: compact");
        const long long launches_before_finish = g_launches.load();
        rc = finish_aggregate(hp, t, rt, q, entries.as<uint8_t>(), G, true, (int64_t)hs.rows_scanned, out, st);
        if (rc == CQG_OK && g_launches.load() != launches_before_finish) {  // the finish ran kernels that can raise flags
            unsigned f = 0;
            CU(cudaMemcpy(&f, hp.P.errflags, 4, cudaMemcpyDeviceToHost));
            if (f & kFatalMask) {
                cqg_result_free(*out);
                *out = nullptr;
                rc = fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(f & kFatalMask));
            }
        }
    }
    if (rc == CQG_OK) {
        (*out)->kernel_ms = ms;
        (*out)->kernel_launches = (int32_t)(g_launches.load() - launches0);
    }
    return rc;
}

// csv_load's Row::column_count for the rows at `row_offsets` (host array, e.g. cqg_result_t::row_offset of a projection)
CQG_API int cqg_table_field_counts(const cqg_table_t* t, const uint64_t* row_offsets, int64_t n, int32_t* counts) {
    if (!t || n < 0 || (n && (!row_offsets || !counts))) return fail(CQG_ERR_ARG, "bad argument");
    if (n == 0) return CQG_OK;
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = ensure_staged(t))) return rc;
    if (t->ngpu > 1) CU(cudaSetDevice(t->devices[0]));
    DevBuf d_off, d_cnt;
    CU(d_off.alloc((size_t)n * 8, 0));
    CU(d_cnt.alloc((size_t)n * 4, 0));
    CU(cudaMemcpyAsync(d_off.p, row_offsets, (size_t)n * 8, cudaMemcpyHostToDevice, 0));
    const int grid = (int)std::min<int64_t>((n + 127) / 128, 148 * 16);
    field_count_kernel<<<grid, 128>>>(t->d_data, t->size, d_off.as<uint64_t>(), (uint64_t)n, (uint8_t)t->cfg.delimiter, (uint8_t)t->cfg.quote,
                                      d_cnt.as<int32_t>());
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy(counts, d_cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return CQG_OK;
}

CQG_API int cqg_table_row_count(const cqg_table_t* t, int64_t* out) {
    if (!t || !out) return fail(CQG_ERR_ARG, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = ensure_staged(t))) return rc;
    return count_rows_device(t, false, 0, out);
}

// ------------------------------------------------------------------------------------------
// parse_value on the device
// ------------------------------------------------------------------------------------------
CQG_API int cqg_parse_value(const char* str, size_t len, cqg_value_t* out) {
    if (!out || (!str && len)) return fail(CQG_ERR_ARG, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (len > (1u << 30)) return fail(CQG_ERR_ARG, "field too long");
    DevBuf d_s, d_o, d_e;
    CU(d_s.alloc(len + 64, 0));
    CU(cudaMemsetAsync(d_s.p, 0, len + 64, 0));
    if (len) CU(cudaMemcpyAsync(d_s.p, str, len, cudaMemcpyHostToDevice, 0));
    CU(d_o.alloc(sizeof(OutCell), 0));
    CU(d_e.alloc(4, 0));
    CU(cudaMemsetAsync(d_e.p, 0, 4, 0));
    parse_value_kernel<<<1, 1>>>(d_s.as<uint8_t>(), (uint32_t)len, d_o.as<OutCell>(), d_e.as<unsigned>());
    g_launches++;
    CU(cudaGetLastError());
    OutCell c;
    unsigned f = 0;
    CU(cudaMemcpy(&c, d_o.p, sizeof c, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&f, d_e.p, 4, cudaMemcpyDeviceToHost));
    if (f & kFatalMask) return fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(f & kFatalMask));
    memset(out, 0, sizeof *out);
    out->type = c.type;
    switch (c.type) {
        case CQG_TYPE_INTEGER: out->int_value = (long long)c.payload; break;
        case CQG_TYPE_DOUBLE: memcpy(&out->double_value, &c.payload, 8); break;
        case CQG_TYPE_DATE:
            out->date_value.year = (int)(c.payload >> 16);
            out->date_value.month = (int)((c.payload >> 8) & 0xff);
            out->date_value.day = (int)(c.payload & 0xff);
            break;
        case CQG_TYPE_STRING: {
            char* s = (char*)calloc(1, (size_t)c.len + 1);
            memcpy(s, str + c.payload, c.len);
            out->string_value = s;
            break;
        }
        default: break;
    }
    return CQG_OK;
}

CQG_API void cqg_value_release(cqg_value_t* v) {
    if (v && v->type == CQG_TYPE_STRING) {
        free(v->string_value);
        v->string_value = nullptr;
    }
}

// ------------------------------------------------------------------------------------------
// multi-GPU partial aggregates
// ------------------------------------------------------------------------------------------
struct cqg_partial {
    HostPlan hp;
    GroupTable gt;
    bool fresh = false;  // made by cqg_partial_new_like and nothing merged into it yet: it holds no groups
    JoinState js;  // build side of an equi-join (whole right table, or the rows this rank owns)
    JoinState js_finish;  // hash-partitioned joins: the whole right table again, built by the rank that finishes
    bool owned_rows_only = false;  // js covers only the keys one rank owns
    cqg_query_t q{};
    int64_t rows_scanned = 0;
    double kernel_ms = 0;
};

// (RIGHT / FULL joins need every left row's matches before a right row is known to be without one: not per shard)
static bool partial_query_ok(const cqg_query_t* q) {
    return q->mode == CQG_MODE_AGGREGATE && !(q->join.right && q->join.type >= CQG_JOIN_RIGHT);
}

CQG_API int cqg_execute_partial(const cqg_table_t* t, const cqg_query_t* q, cqg_partial_t** out) {
    if (!t || !q || !out) return fail(CQG_ERR_ARG, "null argument");
    if (!partial_query_ok(q)) return fail(CQG_ERR_UNSUPPORTED, "partials cover aggregates (with or without a join)");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = ensure_staged(t)) || (rc = ensure_staged(q->join.right))) return rc;
    cqg_partial* p = new cqg_partial();
    p->q = *q;
    p->q.where.code = nullptr;  // the partial keeps no reference to caller memory
    p->q.where.consts = nullptr;
    p->q.where.n_code = p->q.where.n_consts = 0;
    if ((rc = build_plan(p->hp, t, q, 0))) {
        delete p;
        return rc;
    }
    ScalarBlock hs{};
    float ms = 0;
    // joins: every rank builds the WHOLE right table and probes it with its own shard of the left one
    if (q->join.right && (rc = build_join(p->hp, p->js, q->join.right, q->join.right_col, 0))) {
        delete p;
        return rc;
    }
    rc = run_aggregate_scan(p->hp, p->gt, 0, hs, &ms);
    hs.errflags |= p->js.flags;
    hs.jclass[1] |= p->js.key_classes;
    if (rc == CQG_OK) rc = check_flags(hs, q->join.right != nullptr);
    if (rc != CQG_OK) {
        delete p;
        return rc;
    }
    p->rows_scanned = (int64_t)hs.rows_scanned;
    p->kernel_ms = ms;
    *out = p;
    return CQG_OK;
}

// ------------------------------------------------------------------------------------------
// hash-partitioned equi-join across ranks: row offsets split by key owner, exchanged by the
// caller (NCCL all-to-all), then build + probe + aggregate over exactly the rows a rank owns
// ------------------------------------------------------------------------------------------
struct cqg_rowlist {
    DevBuf list;
    std::vector<int64_t> counts;
    unsigned key_classes = 0;  // OR of (1 << comparison class) over the keys of this shard (bit 0: NULL)
};

CQG_API int cqg_partition_rows(const cqg_table_t* t, int key_col, int world, cqg_rowlist_t** out) {
    if (!t || !out || world < 1 || world > 4096) return fail(CQG_ERR_ARG, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = ensure_staged(t))) return rc;
    cudaStream_t st = 0;
    cqg_rowlist* rl = new cqg_rowlist();
    rl->counts.assign((size_t)world, 0);
    const int ncols = (int)t->names.size();
    if (key_col < 0 || key_col >= ncols) {  // resolve_column -> NULL: the ON condition is false for every row (joins.c:54)
        *out = rl;
        return CQG_OK;
    }
    HostPlan hp;
    DevPlan& P = hp.P;
    set_file(P, t, false);
    P.mode = SCAN_PARTITION;
    for (int c = 0; c < kMaxQueryCols; c++) P.colslot[c] = -1;
    P.n_left_cols = ncols;
    P.n_cols_total = ncols;
    P.jr_col = key_col;
    P.wantL[0] = (int16_t)key_col;
    P.nwantL = 1;
    P.colslot[key_col] = 0;
    P.part_world = world;
    DevBuf d_counts, d_base;
    auto fail_free = [&](int code) {
        delete rl;
        return code;
    };
    if (hp.d_scalars.alloc(sizeof(ScalarBlock), st) != cudaSuccess || d_counts.alloc((size_t)world * 8, st) != cudaSuccess ||
        d_base.alloc((size_t)world * 8, st) != cudaSuccess)
        return fail_free(fail(CQG_ERR_NOMEM, "partition buffers"));
    ScalarBlock* sb = hp.d_scalars.as<ScalarBlock>();
    cudaMemsetAsync(sb, 0, sizeof(ScalarBlock), st);
    P.errflags = &sb->errflags;
    P.rows_scanned = &sb->rows_scanned;
    P.jclass = sb->jclass;
    P.gcount = &sb->gcount;
    P.sel_count = &sb->sel_count;
    P.jrow_count = &sb->jrow_count;
    P.part_counts = d_counts.as<unsigned long long>();
    P.part_base = d_base.as<uint64_t>();
    P.part_list = nullptr;
    // pass 1: rows per owner
    cudaMemsetAsync(d_counts.p, 0, (size_t)world * 8, st);
    if ((rc = launch_scan(P, 0, st))) return fail_free(rc);
    std::vector<uint64_t> cnt((size_t)world), base((size_t)world);
    if (cudaMemcpyAsync(cnt.data(), d_counts.p, (size_t)world * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return fail_free(fail(CQG_ERR_CUDA, "partition count: %s", cudaGetErrorString(cudaGetLastError())));
    uint64_t total = 0;
    for (int o = 0; o < world; o++) {
        base[(size_t)o] = total;
        total += cnt[(size_t)o];
        rl->counts[(size_t)o] = (int64_t)cnt[(size_t)o];
    }
    // pass 2: the offsets, every owner's segment dense
    if (rl->list.alloc((total + 1) * 8, st) != cudaSuccess) return fail_free(fail(CQG_ERR_NOMEM, "partition list"));
    cudaMemcpyAsync(d_base.p, base.data(), (size_t)world * 8, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_counts.p, 0, (size_t)world * 8, st);
    P.part_list = rl->list.as<uint64_t>();
    if ((rc = launch_scan(P, 0, st))) return fail_free(rc);
    ScalarBlock h{};
    if (cudaMemcpyAsync(&h, sb, sizeof h, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return fail_free(fail(CQG_ERR_CUDA, "partition scan: %s", cudaGetErrorString(cudaGetLastError())));
    if (h.errflags & kFatalMask) return fail_free(fail(CQG_ERR_UNSUPPORTED, "%s", flag_text(h.errflags & kFatalMask)));
    rl->key_classes = h.jclass[0];
    *out = rl;
    return CQG_OK;
}
CQG_API unsigned cqg_rowlist_key_classes(const cqg_rowlist_t* rl) { return rl ? rl->key_classes : 0u; }
CQG_API uint64_t cqg_rowlist_device_ptr(const cqg_rowlist_t* rl) { return rl ? (uint64_t)(uintptr_t)rl->list.p : 0; }
CQG_API int cqg_rowlist_counts(const cqg_rowlist_t* rl, int world, int64_t* counts) {
    if (!rl || !counts || world != (int)rl->counts.size()) return fail(CQG_ERR_ARG, "bad argument");
    for (int o = 0; o < world; o++) counts[o] = rl->counts[(size_t)o];
    return CQG_OK;
}
CQG_API int cqg_rowlist_copy(const cqg_rowlist_t* rl, uint64_t dst_device_ptr, int64_t capacity) {
    if (!rl) return fail(CQG_ERR_ARG, "null argument");
    int64_t total = 0;
    for (int64_t c : rl->counts) total += c;
    if (capacity < total) return fail(CQG_ERR_ARG, "row list holds %lld offsets", (long long)total);
    if (total) {
        CU(cudaMemcpyAsync((void*)(uintptr_t)dst_device_ptr, rl->list.p, (size_t)total * 8, cudaMemcpyDeviceToDevice, 0));
        CU(cudaStreamSynchronize(0));
    }
    return CQG_OK;
}
CQG_API void cqg_rowlist_free(cqg_rowlist_t* rl) { delete rl; }

// the aggregate of `q` (an equi-join) over `n_left` rows of the left file against `n_right` rows of the right one
CQG_API int cqg_execute_partial_rows(const cqg_table_t* t, const cqg_query_t* q, uint64_t left_rows_device_ptr, int64_t n_left,
                                     uint64_t right_rows_device_ptr, int64_t n_right, cqg_partial_t** out) {
    if (!t || !q || !out || n_left < 0 || n_right < 0) return fail(CQG_ERR_ARG, "bad argument");
    if (!partial_query_ok(q) || !q->join.right) return fail(CQG_ERR_UNSUPPORTED, "row-list partials cover aggregates over an equi-join");
    if (t->global_base != 0 || q->join.right->global_base != 0)
        return fail(CQG_ERR_UNSUPPORTED, "row-list partials read whole files: global offset must be 0");
    int rc = ensure_device();
    if (rc) return rc;
    if ((rc = ensure_staged(t)) || (rc = ensure_staged(q->join.right))) return rc;
    cudaStream_t st = 0;
    cqg_partial* p = new cqg_partial();
    p->q = *q;
    p->q.where.code = nullptr;
    p->q.where.consts = nullptr;
    p->q.where.n_code = p->q.where.n_consts = 0;
    auto bail = [&](int code) {
        delete p;
        return code;
    };
    if ((rc = build_plan(p->hp, t, q, st))) return bail(rc);
    const uint64_t* lrows = (const uint64_t*)(uintptr_t)left_rows_device_ptr;
    const uint64_t* rrows = (const uint64_t*)(uintptr_t)right_rows_device_ptr;
    if ((rc = build_join(p->hp, p->js, q->join.right, q->join.right_col, st, rrows ? rrows : (const uint64_t*)8, n_right))) return bail(rc);  // (any non-null pointer with n_right == 0: an empty build side)
    DevPlan& P = p->hp.P;
    ScalarBlock hs{};
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return bail(fail(CQG_ERR_CUDA, "events"));
    float ms_total = 0;
    uint64_t cap = P.ngc == 0 ? 16 : 1 << 12;
    while (P.ngc && cap < (uint64_t)n_left / 4 && cap < (1ull << 22)) cap <<= 1;
    for (int attempt = 0; attempt < 12; attempt++) {
        if ((rc = alloc_group_table(p->hp, p->gt, cap, st))) break;
        cudaMemsetAsync(p->hp.d_scalars.p, 0, sizeof(ScalarBlock), st);
        cudaEventRecord(e0, st);
        if (n_left > 0) {
            DevPlan R = P;
            R.simple = 0;
            R.scalar_regs = 0;
            R.smem_cap = 0;
            int grid = (int)std::min<uint64_t>(((uint64_t)n_left + 127) / 128, 148 * 8);
            deferred_rows_kernel<<<grid, 128, 0, st>>>(R, lrows, (uint64_t)n_left);
            g_launches++;
        }
        cudaEventRecord(e1, st);
        if ((rc = read_scalars(p->hp, hs, st))) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms_total += ms;
        if (hs.errflags & KERR_TABLE_FULL) {
            cap *= 4;
            if (cap > (1ull << 32)) {
                rc = fail(CQG_ERR_NOMEM, "group table beyond 2^32 entries");
                break;
            }
            continue;
        }
        break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    hs.errflags |= p->js.flags;
    hs.jclass[1] |= p->js.key_classes;
    if (rc == CQG_OK) rc = check_flags(hs, true);
    if (rc != CQG_OK) return bail(rc);
    p->rows_scanned = (int64_t)hs.rows_scanned;
    p->kernel_ms = ms_total;
    p->owned_rows_only = true;
    *out = p;
    return CQG_OK;
}

CQG_API double cqg_partial_kernel_ms(const cqg_partial_t* p) { return p ? p->kernel_ms : 0.0; }
CQG_API int64_t cqg_partial_rows_scanned(const cqg_partial_t* p) { return p ? p->rows_scanned : 0; }
CQG_API size_t cqg_partial_record_size(const cqg_partial_t* p) { return p ? (size_t)p->hp.P.entry_bytes : 0; }

CQG_API int64_t cqg_partial_count(const cqg_partial_t* p) {
    if (!p) return -1;
    unsigned long long n = 0;
    if (cudaMemcpy(&n, p->hp.P.gcount, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)n;
}

CQG_API int cqg_partial_owner_counts(const cqg_partial_t* p, int world, int64_t* counts) {
    if (!p || !counts || world < 1) return fail(CQG_ERR_ARG, "bad argument");
    DevBuf d;
    CU(d.alloc(8 * (size_t)world, 0));
    CU(cudaMemsetAsync(d.p, 0, 8 * (size_t)world, 0));
    int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((p->gt.cap + 255) / 256, 148 * 8));
    owner_count_kernel<<<grid, 256>>>(p->gt.tab.as<uint8_t>(), p->gt.cap, p->hp.P.entry_bytes, world, d.as<unsigned long long>());
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy(counts, d.p, 8 * (size_t)world, cudaMemcpyDeviceToHost));
    return CQG_OK;
}

CQG_API int cqg_partial_export(const cqg_partial_t* p, int owner, int world, uint64_t dst_device_ptr, int64_t capacity,
                               int64_t* n_out) {
    if (!p || !n_out || capacity < 0) return fail(CQG_ERR_ARG, "bad argument");
    DevBuf cnt;
    CU(cnt.alloc(8, 0));
    CU(cudaMemsetAsync(cnt.p, 0, 8, 0));
    int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((p->gt.cap + 255) / 256, 148 * 8));
    compact_table_kernel<<<grid, 256>>>(p->gt.tab.as<uint8_t>(), p->gt.cap, p->hp.P.entry_bytes, (uint8_t*)dst_device_ptr,
                                        (uint64_t)capacity, cnt.as<unsigned long long>(), owner, world);
    g_launches++;
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpy(&n, cnt.p, 8, cudaMemcpyDeviceToHost));
    *n_out = (int64_t)n;
    if ((int64_t)n > capacity) return fail(CQG_ERR_ARG, "export buffer too small: %lld records", (long long)n);
    return CQG_OK;
}

CQG_API int cqg_partial_new_like(const cqg_partial_t* like, cqg_partial_t** out) {
    if (!like || !out) return fail(CQG_ERR_ARG, "null argument");
    cqg_partial* p = new cqg_partial();
    p->q = like->q;
    p->owned_rows_only = like->owned_rows_only;
    p->hp.P = like->hp.P;
    p->hp.entry_init = like->hp.entry_init;
    p->hp.table_smem_bytes = like->hp.table_smem_bytes;
    DevPlan& P = p->hp.P;
    P.pred_kind = 0;  // merging never evaluates the predicate
    int rc = CQG_OK;
    do {
        cudaError_t e = p->hp.d_entry_init.alloc(p->hp.entry_init.size(), 0);
        if (e == cudaSuccess)
            e = cudaMemcpy(p->hp.d_entry_init.p, p->hp.entry_init.data(), p->hp.entry_init.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = p->hp.d_scalars.alloc(sizeof(ScalarBlock), 0);
        if (e == cudaSuccess) e = cudaMemsetAsync(p->hp.d_scalars.p, 0, sizeof(ScalarBlock), 0);
        if (e != cudaSuccess) {
            rc = fail(CQG_ERR_CUDA, "partial_new_like: %s", cudaGetErrorString(e));
            break;
        }
        P.entry_init = p->hp.d_entry_init.as<uint8_t>();
        ScalarBlock* sb = p->hp.d_scalars.as<ScalarBlock>();
        P.errflags = &sb->errflags;
        P.jclass = sb->jclass;
        P.rows_scanned = &sb->rows_scanned;
        P.gcount = &sb->gcount;
        P.sel_count = &sb->sel_count;
        P.jrow_count = &sb->jrow_count;
        P.pcount = &sb->pcount;
        P.ptab = nullptr;
        P.pcap = 0;
        uint64_t cap = 16;  // (a dense table's `cap` is its entry count, not a power of two)
        while (cap < (like->gt.dense ? 2 * like->gt.cap : like->gt.cap)) cap <<= 1;
        rc = alloc_group_table(p->hp, p->gt, cap, 0);
    } while (0);
    if (rc != CQG_OK) {
        delete p;
        return rc;
    }
    p->fresh = true;
    *out = p;
    return CQG_OK;
}

CQG_API int cqg_partial_merge(cqg_partial_t* p, uint64_t src_device_ptr, int64_t n) {
    if (!p || n < 0) return fail(CQG_ERR_ARG, "bad argument");
    if (n == 0) return CQG_OK;
    DevPlan& P = p->hp.P;
    // room first: every record may be a new group, and a batch that overflowed half way could not be retried
    // (the records already folded in would be added twice)
    unsigned long long have = 0;
    if (!p->fresh) CU(cudaMemcpy(&have, P.gcount, 8, cudaMemcpyDeviceToHost));  // (a fresh table: no round trip for a 0)
    p->fresh = false;
    while ((have + (unsigned long long)n) * 2ull > p->gt.cap) {
        if (p->gt.cap > (1ull << 32)) return fail(CQG_ERR_NOMEM, "merge table beyond 2^32 entries");
        DevBuf old_entries;
        uint64_t G = 0;
        int rc = compact_groups(p->hp, p->gt, 0, 1, old_entries, &G, 0);
        if (rc) return rc;
        GroupTable bigger;
        if ((rc = alloc_group_table(p->hp, bigger, p->gt.cap * 4, 0))) return rc;
        std::swap(p->gt.tab.p, bigger.tab.p);
        std::swap(p->gt.tab.n, bigger.tab.n);
        p->gt.cap = bigger.cap;
        if (G) {
            int g2 = (int)std::min<uint64_t>((G + 127) / 128, 148 * 8);
            merge_entries_kernel<<<g2, 128>>>(P, old_entries.as<uint8_t>(), G);
            g_launches++;
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
        }
    }
    CU(cudaMemsetAsync(P.errflags, 0, 4, 0));
    int grid = (int)std::min<int64_t>((n + 127) / 128, 148 * 8);
    merge_entries_kernel<<<grid, 128>>>(P, (const uint8_t*)src_device_ptr, (uint64_t)n);
    g_launches++;
    CU(cudaGetLastError());
    unsigned f = 0;
    CU(cudaMemcpy(&f, P.errflags, 4, cudaMemcpyDeviceToHost));
    if (f & KERR_TABLE_FULL) return fail(CQG_ERR_NOMEM, "merge table overflowed");
    return CQG_OK;
}

CQG_API int cqg_partial_finish(const cqg_partial_t* pc, const cqg_table_t* t, cqg_result_t** out) {
    if (!pc || !t || !out) return fail(CQG_ERR_ARG, "null argument");
    cqg_partial* p = const_cast<cqg_partial*>(pc);
    DevBuf entries;
    uint64_t G = 0;
    int rc = ensure_staged(t);
    if (rc) return rc;
    if ((rc = compact_groups(p->hp, p->gt, 0, 1, entries, &G, 0))) return rc;
    // first-row decoding and string MIN/MAX read the file `t` views
    p->hp.P.data = t->d_data;
    p->hp.P.size = t->size;
    p->hp.P.global_base = t->global_base;
    if (p->owned_rows_only && p->q.join.right) {
        // the groups' representative rows (bare columns, numeric MIN/MAX of right columns) are looked up by
        // (left row, rank of the match): that needs the matches of ANY key, i.e. a table over the whole right file
        bool need = p->q.n_out_cols > 0;
        for (int a = 0; a < p->q.n_aggs; a++)
            need = need || ((p->q.aggs[a].func == CQG_AGG_MIN || p->q.aggs[a].func == CQG_AGG_MAX) && p->q.aggs[a].col >= p->hp.P.n_left_cols);
        if (need) {
            if ((rc = build_join(p->hp, p->js_finish, p->q.join.right, p->q.join.right_col, 0))) return rc;
            p->owned_rows_only = false;
        }
    }
    rc = finish_aggregate(p->hp, t, p->q.join.right, &p->q, entries.as<uint8_t>(), G, true, p->rows_scanned, out, 0);
    if (rc == CQG_OK) {
        (*out)->kernel_ms = p->kernel_ms;
        (*out)->kernel_launches = 0;
    }
    return rc;
}

CQG_API void cqg_partial_free(cqg_partial_t* p) { delete p; }

// An aggregate over a multi-GPU table: every device scans the rows that start in its slice (cqg_execute_partial on the
// slice's view, one host thread per device), the partial group records of devices 1.. cross NVLink into device 0
// (cudaMemcpyPeer: one process, peer memory, no collective library needed), are merged there into one table and finished
// there - the groups' first rows and MIN/MAX strings are read through the table's one address range wherever they live.
static int execute_multi(const cqg_table* t, const cqg_query_t* q, cqg_result_t** out) {
    const int N = t->ngpu;
    int home = 0;
    CU(cudaGetDevice(&home));
    struct Slot {
        int rc = CQG_OK;
        std::string err;
        cqg_partial_t* part = nullptr;
        void* recs = nullptr;
        int64_t n = 0;
    };
    std::vector<Slot> slots(N);
    const long long launches0 = g_launches.load();
    std::vector<std::thread> th;
    for (int d = 0; d < N; d++) {
        th.emplace_back([&, d] {
            Slot& s = slots[d];
            if (cudaSetDevice(t->devices[d]) != cudaSuccess) {
                s.rc = CQG_ERR_CUDA;
                s.err = "cudaSetDevice failed";
                return;
            }
            s.rc = cqg_execute_partial(t->views[d], q, &s.part);
            if (s.rc == CQG_OK) {
                s.n = cqg_partial_count(s.part);
                if (s.n < 0) s.rc = CQG_ERR_CUDA;
            }
            if (s.rc == CQG_OK && s.n > 0) {
                const size_t rb = cqg_partial_record_size(s.part);
                if (cudaMalloc(&s.recs, (size_t)s.n * rb) != cudaSuccess) {
                    s.rc = fail(CQG_ERR_NOMEM, "records of %lld groups: out of device memory", (long long)s.n);
                } else {
                    int64_t got = 0;
                    s.rc = cqg_partial_export(s.part, 0, 1, (uint64_t)(uintptr_t)s.recs, s.n, &got);
                    s.n = got;
                }
            }
            if (s.rc != CQG_OK) s.err = cqg_last_error();
        });
    }
    for (std::thread& x : th) x.join();
    cudaSetDevice(t->devices[0]);
    int rc = CQG_OK;
    cqg_partial_t* merged = nullptr;
    for (int d = 0; d < N && rc == CQG_OK; d++)
        if (slots[d].rc != CQG_OK) rc = fail(slots[d].rc, "%s", slots[d].err.c_str());
    if (rc == CQG_OK) rc = cqg_partial_new_like(slots[0].part, &merged);
    double kernel_ms = 0;
    int64_t rows = 0;
    for (int d = 0; d < N && rc == CQG_OK; d++) {
        Slot& s = slots[d];
        kernel_ms = std::max(kernel_ms, cqg_partial_kernel_ms(s.part));
        rows += cqg_partial_rows_scanned(s.part);
        if (s.n == 0) continue;
        if (t->devices[d] == t->devices[0]) {
            rc = cqg_partial_merge(merged, (uint64_t)(uintptr_t)s.recs, s.n);
        } else {
            const size_t bytes = (size_t)s.n * cqg_partial_record_size(s.part);
            void* here = nullptr;
            if (cudaMalloc(&here, bytes) != cudaSuccess) {
                rc = fail(CQG_ERR_NOMEM, "records of %lld groups: out of device memory", (long long)s.n);
                break;
            }
            if (cudaMemcpyPeer(here, t->devices[0], s.recs, t->devices[d], bytes) != cudaSuccess)
                rc = fail(CQG_ERR_CUDA, "peer copy of partial records: %s", cudaGetErrorString(cudaGetLastError()));
            else
                rc = cqg_partial_merge(merged, (uint64_t)(uintptr_t)here, s.n);
            cudaDeviceSynchronize();
            cudaFree(here);
        }
    }
    if (rc == CQG_OK) {
        merged->rows_scanned = rows;
        merged->kernel_ms = kernel_ms;
        rc = cqg_partial_finish(merged, t, out);
        if (rc == CQG_OK) (*out)->kernel_launches = (int32_t)(g_launches.load() - launches0);
    }
    std::string keep = rc != CQG_OK ? cqg_last_error() : "";
    cqg_partial_free(merged);
    for (int d = 0; d < N; d++) {
        cudaSetDevice(t->devices[d]);
        if (slots[d].recs) cudaFree(slots[d].recs);
        cqg_partial_free(slots[d].part);
    }
    cudaSetDevice(home);
    if (rc != CQG_OK) return fail(rc, "%s", keep.c_str());
    return CQG_OK;
}


// ------------------------------------------------------------------------------------------
// synthetic data
// ------------------------------------------------------------------------------------------
CQG_API size_t cqg_generate_bigdata_bound(int64_t rows, int64_t key_card) {
    return 64 + (size_t)rows * (size_t)(31 + (key_card > 0 ? 21 : 0));
}

CQG_API int cqg_generate_bigdata(uint64_t device_ptr, size_t capacity, int64_t rows, uint64_t seed, int64_t key_card,
                                 size_t* size_out) {
    return cqg_generate_bigdata_range(device_ptr, capacity, 0, rows, seed, key_card, 1, size_out);
}

CQG_API int cqg_generate_bigdata_range(uint64_t device_ptr, size_t capacity, int64_t row_start, int64_t rows, uint64_t seed,
                                       int64_t key_card, int with_header, size_t* size_out) {
    if (!device_ptr || !size_out || rows < 0 || row_start < 0) return fail(CQG_ERR_ARG, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (capacity < cqg_generate_bigdata_bound(rows, key_card)) return fail(CQG_ERR_ARG, "capacity below cqg_generate_bigdata_bound");
    const char* hdr = key_card > 0 ? "name,surname,age,gender,height,uid\n" : "name,surname,age,gender,height\n";
    size_t hl = with_header ? strlen(hdr) : 0;
    if (hl) CU(cudaMemcpy((void*)device_ptr, hdr, hl, cudaMemcpyHostToDevice));
    int64_t nblocks = (rows + kGenRowsPerBlock - 1) / kGenRowsPerBlock;
    if (nblocks == 0) {
        *size_out = hl;
        return CQG_OK;
    }
    if (nblocks > 0x7fffffff) return fail(CQG_ERR_ARG, "too many rows");
    DevBuf d_sizes;
    CU(d_sizes.alloc((size_t)nblocks * 8, 0));
    gen_sizes_kernel<<<(unsigned)nblocks, 256>>>(row_start, rows, seed, key_card, d_sizes.as<unsigned long long>());
    g_launches++;
    CU(cudaGetLastError());
    std::vector<unsigned long long> sizes((size_t)nblocks);
    CU(cudaMemcpy(sizes.data(), d_sizes.p, (size_t)nblocks * 8, cudaMemcpyDeviceToHost));
    unsigned long long off = hl;
    for (int64_t b = 0; b < nblocks; b++) {
        unsigned long long s = sizes[(size_t)b];
        sizes[(size_t)b] = off;
        off += s;
    }
    if (off > capacity) return fail(CQG_ERR_ARG, "capacity too small");
    CU(cudaMemcpy(d_sizes.p, sizes.data(), (size_t)nblocks * 8, cudaMemcpyHostToDevice));
    gen_write_kernel<<<(unsigned)nblocks, 256>>>((uint8_t*)device_ptr, row_start, rows, seed, key_card, d_sizes.as<unsigned long long>());
    g_launches++;
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    *size_out = (size_t)off;
    return CQG_OK;
}
