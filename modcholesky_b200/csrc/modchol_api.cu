// modchol_api.cu -- the C-ABI of include/modchol_b200.h: argument checking, size-class
// dispatch, the overlapped host-staging pipeline and the multi-GPU batch split.
// There is no CPU fallback anywhere in this file: without a CUDA device every compute
// entry point fails with MODCHOL_ERR_NO_DEVICE.
#include "../../include/modchol_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "aux_kernels.cuh"
#include "launch_cache.cuh"
#include "wsm_se.cuh"
#include "common.cuh"
#include "cta_kernels.cuh"
#include "warp_kernels.cuh"

namespace {

using namespace modchol;

cudaError_t launch_solve(long long batch, int n, int nrhs, const double* l, const unsigned long long* p,
                         const double* b, double* x, int sms, cudaStream_t st);

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(MODCHOL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                         \
    } while (0)

// ---- per-device cache ------------------------------------------------------------------
constexpr int kSlots = 3;
struct Slot {
    cudaStream_t stream = nullptr;
    double *a = nullptr, *l = nullptr, *e = nullptr;
    unsigned long long* p = nullptr;
    int* k = nullptr;
    size_t cap_a = 0, cap_v = 0, cap_k = 0;  // capacities in elements: a/l, e/p, k
    // pinned bounce buffers for callers whose memory is pageable (numpy, ndarray): allocated on
    // first use, same capacities as the device buffers
    double *ha = nullptr, *hl = nullptr, *he = nullptr;
    unsigned long long* hp = nullptr;
    int* hk = nullptr;
    size_t hcap_a = 0, hcap_v = 0, hcap_k = 0;
};
struct DeviceCtx {
    std::mutex mu;  // serialises host-pointer calls on this device
    std::once_flag once;   // guards the three attribute fields below
    int init_rc = 0;       // cudaError_t of the attribute queries
    int sms = 0;
    int smem_optin = 0;
    Slot slots[kSlots];
};
constexpr int kMaxDevices = 64;
DeviceCtx g_ctx[kMaxDevices];

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// Device attributes, queried once per device; safe to call from any thread (DEVICE-mode calls
// never take the device mutex).
int ctx_init(int dev)
{
    DeviceCtx& c = g_ctx[dev];
    std::call_once(c.once, [&c, dev]() {
        cudaError_t e = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess)
            e = cudaDeviceGetAttribute(&c.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        // the batched solve keeps 4 n doubles per block in dynamic shared memory: 64 KB at
        // n = MODCHOL_MAX_N, above the 48 KB default
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(4 * MODCHOL_MAX_N * sizeof(double)));
        // stream-ordered scratch (factor_solve_device, the spill buffers of the large-n kernels)
        // stays in the device's pool between calls instead of going back to the driver at every
        // synchronisation
        if (e == cudaSuccess) {
            cudaMemPool_t pool = nullptr;
            unsigned long long keep = (unsigned long long)1 << 30;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            cudaGetLastError();
        }
        c.init_rc = (int)e;
    });
    if (c.init_rc != 0)
        return fail(MODCHOL_ERR_CUDA, "device %d attribute query failed: %s", dev,
                    cudaGetErrorString((cudaError_t)c.init_rc));
    return MODCHOL_OK;
}

void slot_free(Slot& s)
{
    if (s.a) cudaFree(s.a);
    if (s.l) cudaFree(s.l);
    if (s.e) cudaFree(s.e);
    if (s.p) cudaFree(s.p);
    if (s.k) cudaFree(s.k);
    s.a = s.l = s.e = nullptr;
    s.p = nullptr;
    s.k = nullptr;
    s.cap_a = s.cap_v = s.cap_k = 0;
}

void slot_free_host(Slot& s)
{
    if (s.ha) cudaFreeHost(s.ha);
    if (s.hl) cudaFreeHost(s.hl);
    if (s.he) cudaFreeHost(s.he);
    if (s.hp) cudaFreeHost(s.hp);
    if (s.hk) cudaFreeHost(s.hk);
    s.ha = s.hl = s.he = nullptr;
    s.hp = nullptr;
    s.hk = nullptr;
    s.hcap_a = s.hcap_v = s.hcap_k = 0;
}

int slot_reserve_host(Slot& s, size_t mats, int n)
{
    const size_t need_a = mats * (size_t)n * n, need_v = mats * (size_t)n, need_k = mats;
    if (s.ha && need_a <= s.hcap_a && need_v <= s.hcap_v && need_k <= s.hcap_k)
        return MODCHOL_OK;
    const size_t ca = std::max(need_a, s.hcap_a), cv = std::max(need_v, s.hcap_v),
                 ck = std::max(need_k, s.hcap_k);
    slot_free_host(s);
    CUDA_TRY(cudaHostAlloc(&s.ha, ca * sizeof(double), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s.hl, ca * sizeof(double), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s.he, cv * sizeof(double), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s.hp, cv * sizeof(unsigned long long), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc(&s.hk, ck * sizeof(int), cudaHostAllocDefault));
    s.hcap_a = ca;
    s.hcap_v = cv;
    s.hcap_k = ck;
    return MODCHOL_OK;
}

// true when the DMA engines can read/write this host range directly (cudaHostAlloc,
// cudaHostRegister or managed memory); pageable memory goes through the pinned bounce buffers
bool host_ptr_is_pinned(const void* ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// memcpy on a few host threads (one core moves ~10 GB/s; the PCIe link wants ~50)
void parallel_memcpy(void* dst, const void* src, size_t bytes)
{
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned want = (unsigned)std::min<size_t>(std::min(8u, std::max(1u, hw / 2)), bytes >> 22);
    if (want <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / want) + 4095) & ~(size_t)4095;
    for (unsigned t = 0; t < want; ++t) {
        const size_t lo = std::min(bytes, (size_t)t * per), hi = std::min(bytes, lo + per);
        if (hi > lo)
            th.emplace_back([=]() { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    for (auto& t : th)
        t.join();
}

int slot_reserve(Slot& s, size_t mats, int n)
{
    if (!s.stream)
        CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    const size_t need_a = mats * (size_t)n * n, need_v = mats * (size_t)n, need_k = mats;
    if (s.a && need_a <= s.cap_a && need_v <= s.cap_v && need_k <= s.cap_k)
        return MODCHOL_OK;
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    const size_t ca = std::max(need_a, s.cap_a), cv = std::max(need_v, s.cap_v),
                 ck = std::max(need_k, s.cap_k);
    slot_free(s);
    CUDA_TRY(cudaMalloc(&s.a, ca * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s.l, ca * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s.e, cv * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s.p, cv * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&s.k, ck * sizeof(int)));
    s.cap_a = ca;
    s.cap_v = cv;
    s.cap_k = ck;
    return MODCHOL_OK;
}

// ---- size-class dispatch ---------------------------------------------------------------
enum KernelClass { kClsWarp = 0, kClsCtaSmem = 1, kClsCtaGmem = 2, kClsWarpSmem = 3, kClsCtaPacked = 4, kClsCtaBlocked = 5, kClsWarpPanel = 6, kClsThread = 7 };

// Which small-n family serves n <= 32 (0 auto, 1 registers, 2 shared memory, 3 thread per matrix
// wherever it exists); initialised from MODCHOL_SMALL_KERNEL=reg|wsm|tpm, changed by
// modchol_set_small_kernel_family().
std::atomic<int> g_small_family{-1};
int small_kernel_pref()
{
    int v = g_small_family.load();
    if (v < 0) {
        const char* e = getenv("MODCHOL_SMALL_KERNEL");
        v = !e ? 0 : (e[0] == 'r' ? 1 : (e[0] == 'w' ? 2 : (e[0] == 't' ? 3 : 0)));
        g_small_family.store(v);
    }
    return v;
}

int cta_ld(int n) { return (n & 1) ? n : n + 1; }

// below this batch the warp kernels serve the call (factor_device): one thread per matrix needs a
// batch to fill the machine.  Measured for 8 < n <= 16 (STRICT, gpurun_out/exp_batch.txt): 2 048
// matrices 77 against 56 us (n = 10), 8 192 matrices 78 against 110 us.
inline long long thread_min_batch(int n) { return n <= kTpmMaxNHost ? 512 : 4096; }

// largest n served by the thread-per-matrix kernels by default (MODCHOL_TPM_MAX / MODCHOL_TPS_MAX,
// tuning knobs read per call; 0 switches them off)
int tpm_max_n(int n, int mode, int algo)
{
    if (n <= kTpmMaxNHost) {
        const char* e = getenv("MODCHOL_TPM_MAX");
        return e ? atoi(e) : 8;   // measured: ahead of the warp kernels at every n <= 8, both modes
    }
    // 8 < n <= 16 (triangle in shared memory): ahead of the position-ordered STRICT kernels at
    // every n <= 16; against the implicit-pivoting FAST kernels ahead up to n = 12 (SE99: 602
    // against 567 M/s on symmetric-uniform input, 608 against 423 on the reference's bench family)
    // and up to n = 10 for GMW81 (profiles/r02_thread_per_matrix.txt)
    const char* e = getenv("MODCHOL_TPS_MAX");
    return e ? atoi(e) : (mode == MODCHOL_MODE_FAST ? (algo == MODCHOL_GMW81 ? 10 : 12) : 16);
}

// smallest n served by the blocked large-n kernels (MODCHOL_BSM_MIN, tuning knob)
int bsm_min_n()
{
    static const int v = [] {
        const char* e = getenv("MODCHOL_BSM_MIN");
        return e ? atoi(e) : 129;   // measured: ahead of the packed kernels from n = 136 on (GMW81)
    }();
    return v;
}

// The warp-per-matrix panel kernels (lwm_kernels.cuh, 64 < n <= 256) are opt-in: measured behind
// the block-level kernels at every size.  MODCHOL_LWM_MIN = smallest n they
// serve, read per call so that tests and benchmarks can switch them on.
int lwm_min_n()
{
    const char* e = getenv("MODCHOL_LWM_MIN");
    return e ? atoi(e) : 100000;
}

KernelClass classify(int algo, int mode, int n, int smem_optin)
{
    const bool reg_ok = warp_kernel_available(algo, mode, n);
    const bool wsm_ok = wsm_kernel_available(algo, mode, n);
    const int pref = small_kernel_pref();
    // Measured: the shared-memory family (left-looking, 32 warps per SM,
    // 32/16/8/4 lanes per matrix) is at least as fast as the register family at every n in both
    // modes for all but a few (n <= 8, STRICT, mixed-phase) cases, and it is the only warp-level
    // kernel for 32 < n <= 64 -- it is the default; the register family stays selectable.
    const bool auto_wsm = wsm_ok;
    if (tpm_kernel_available(algo, n) && (pref == 3 || (pref == 0 && n <= tpm_max_n(n, mode, algo))))
        return kClsThread;
    if (wsm_ok && (pref == 2 || (pref == 0 && auto_wsm) || !reg_ok))
        return kClsWarpSmem;
    if (reg_ok)
        return kClsWarp;
    if (lwm_kernel_available(algo, mode, n) && n >= lwm_min_n())
        return kClsWarpPanel;
    if (bsm_kernel_available(algo, mode, n) && n >= bsm_min_n())
        return kClsCtaBlocked;
    if (csm_kernel_available(algo, mode, n))
        return kClsCtaPacked;
    if (smem_optin <= 0)
        smem_optin = 227 * 1024;
    if (cta_shared_bytes(n, cta_ld(n), true) <= (size_t)smem_optin)
        return kClsCtaSmem;
    return kClsCtaGmem;
}

template <int T>
int launch_cta(int algo, int n, long long batch, const double* a, double* l, double* e,
               unsigned long long* p, int* k, bool use_smem, int sms, cudaStream_t st)
{
    const int ld = use_smem ? cta_ld(n) : n;
    const size_t smem = cta_shared_bytes(n, ld, use_smem);
    int occ = 1;
    if (algo == MODCHOL_GMW81) {
        auto kern = gmw81_cta_kernel<T>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
        const long long grid = std::max(1LL, std::min(batch, (long long)sms * std::max(occ, 1)));
        kern<<<(unsigned)grid, T, smem, st>>>(n, ld, use_smem ? 1 : 0, batch, a, l, e, p, k);
    } else {
        auto kern = se_cta_kernel<T>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
        const long long grid = std::max(1LL, std::min(batch, (long long)sms * std::max(occ, 1)));
        kern<<<(unsigned)grid, T, smem, st>>>(algo, n, ld, use_smem ? 1 : 0, batch, a, l, e, p, k);
    }
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return MODCHOL_OK;
}

// all pointers are device pointers on the current device
int factor_device(int dev, int algo, int mode, long long batch, int n, const double* a, double* l,
                  double* e, unsigned long long* p, int* k, cudaStream_t st)
{
    if (batch == 0)
        return MODCHOL_OK;
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const DeviceCtx& c = g_ctx[dev];
    KernelClass cls = classify(algo, mode, n, c.smem_optin);
    // one thread per matrix needs a batch to fill the machine: a handful of matrices (the trait
    // call on one Array2) is served faster by a warp per matrix (12 x 12: 21 against 27 us)
    if (cls == kClsThread && batch < thread_min_batch(n) && small_kernel_pref() != 3)
        cls = kClsWarpSmem;
    if (cls == kClsWarp) {
        cudaError_t ce = launch_warp_kernel(algo, mode, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "warp kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    if (cls == kClsThread) {
        cudaError_t ce = launch_tpm_kernel(algo, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "tpm kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    if (cls == kClsWarpSmem) {
        cudaError_t ce = launch_wsm_kernel(algo, mode, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "wsm kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    if (cls == kClsWarpPanel) {
        cudaError_t ce = launch_lwm_kernel(algo, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "lwm kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    if (cls == kClsCtaBlocked) {
        cudaError_t ce = launch_bsm_kernel(algo, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "bsm kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    if (cls == kClsCtaPacked) {
        cudaError_t ce = launch_csm_kernel(algo, mode, n, batch, a, l, e, p, k, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "csm kernel launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    const bool use_smem = cls == kClsCtaSmem;
    if (n <= 128)
        return launch_cta<128>(algo, n, batch, a, l, e, p, k, use_smem, c.sms, st);
    if (n <= 512)
        return launch_cta<256>(algo, n, batch, a, l, e, p, k, use_smem, c.sms, st);
    return launch_cta<512>(algo, n, batch, a, l, e, p, k, use_smem, c.sms, st);
}

// host pointers: stage the batch through the device in chunks on kSlots streams so that
// H2D of chunk c+1, the kernel of chunk c and D2H of chunk c-1 overlap.  Pinned caller memory is
// read and written by the DMA engines directly; pageable caller memory (numpy, ndarray) goes
// through cached pinned bounce buffers, copied by a few host threads while the other slots'
// transfers run -- a cudaMemcpyAsync on pageable memory would block the host until the chunk's
// kernel is done and serialise the pipeline.
int factor_host_on_device(int dev, int algo, int mode, long long batch, int n, const double* a,
                          double* l, double* e, unsigned long long* p, int* k)
{
    CUDA_TRY(cudaSetDevice(dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    if (batch == 0)
        return MODCHOL_OK;
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lock(c.mu);

    const size_t nn = (size_t)n * n;
    const size_t target_bytes = (size_t)96 << 20;  // per-slot input staging size
    long long chunk = (long long)std::max<size_t>(1, target_bytes / (nn * sizeof(double)));
    chunk = std::min(chunk, (batch + kSlots - 1) / kSlots);
    chunk = std::max<long long>(chunk, 1);
    const int nslots = (int)std::min<long long>(kSlots, (batch + chunk - 1) / chunk);
    const bool bounce_in = !host_ptr_is_pinned(a);
    const bool bounce_out = !(host_ptr_is_pinned(l) && host_ptr_is_pinned(e) && host_ptr_is_pinned(p) &&
                              (!k || host_ptr_is_pinned(k)));
    for (int s = 0; s < nslots; ++s) {
        rc = slot_reserve(c.slots[s], (size_t)chunk, n);
        if (rc == MODCHOL_OK && (bounce_in || bounce_out))
            rc = slot_reserve_host(c.slots[s], (size_t)chunk, n);
        if (rc)
            return rc;
    }
    struct Pending { long long off = 0, cnt = 0; };
    Pending pend[kSlots];
    // results of the chunk a slot last worked on: wait for its stream, then (pageable caller) copy
    // them out of the bounce buffers -- on a helper thread, so that the copy-out of chunk c runs
    // beside the copy-in of chunk c + kSlots on the calling thread (the fresh output pages of a
    // numpy / ndarray caller fault in on first touch, which makes the copy-out the slower half)
    std::future<void> outcopy[kSlots];
    auto wait_outcopy = [&](int si) {
        if (outcopy[si].valid())
            outcopy[si].get();
    };
    auto retire = [&](int si) -> cudaError_t {
        Slot& s = c.slots[si];
        wait_outcopy(si);
        const cudaError_t ce = cudaStreamSynchronize(s.stream);
        const Pending pd = pend[si];
        pend[si] = Pending();
        if (ce != cudaSuccess || pd.cnt == 0 || !bounce_out)
            return ce;
        outcopy[si] = std::async(std::launch::async, [=, &s]() {
            parallel_memcpy(l + (size_t)pd.off * nn, s.hl, (size_t)pd.cnt * nn * sizeof(double));
            memcpy(e + (size_t)pd.off * n, s.he, (size_t)pd.cnt * n * sizeof(double));
            memcpy(p + (size_t)pd.off * n, s.hp, (size_t)pd.cnt * n * sizeof(unsigned long long));
            if (k)
                memcpy(k + pd.off, s.hk, (size_t)pd.cnt * sizeof(int));
        });
        static const bool sync_out = getenv("MODCHOL_SYNC_OUTCOPY") != nullptr;   // A/B knob
        if (sync_out)
            wait_outcopy(si);
        return cudaSuccess;
    };
    cudaError_t ce = cudaSuccess;
    rc = MODCHOL_OK;
    long long off = 0;
    for (int ci = 0; off < batch && ce == cudaSuccess && rc == MODCHOL_OK; ++ci) {
        const long long cnt = std::min(chunk, batch - off);
        const int si = ci % nslots;
        Slot& s = c.slots[si];
        if (pend[si].cnt && (bounce_in || bounce_out))
            ce = retire(si);   // the slot's bounce buffers are about to be reused
        if (ce != cudaSuccess)
            break;
        const double* src = a + (size_t)off * nn;
        if (bounce_in) {
            parallel_memcpy(s.ha, src, (size_t)cnt * nn * sizeof(double));
            src = s.ha;
        }
        ce = cudaMemcpyAsync(s.a, src, (size_t)cnt * nn * sizeof(double), cudaMemcpyHostToDevice, s.stream);
        if (ce != cudaSuccess)
            break;
        rc = factor_device(dev, algo, mode, cnt, n, s.a, s.l, s.e, s.p, k ? s.k : nullptr, s.stream);
        if (rc)
            break;
        wait_outcopy(si);   // the slot's output bounce buffers are about to be overwritten
        double* dl = bounce_out ? s.hl : l + (size_t)off * nn;
        double* de = bounce_out ? s.he : e + (size_t)off * n;
        unsigned long long* dp = bounce_out ? s.hp : p + (size_t)off * n;
        int* dk = bounce_out ? s.hk : (k ? k + off : nullptr);
        ce = cudaMemcpyAsync(dl, s.l, (size_t)cnt * nn * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(de, s.e, (size_t)cnt * n * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(dp, s.p, (size_t)cnt * n * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, s.stream);
        if (ce == cudaSuccess && k)
            ce = cudaMemcpyAsync(dk, s.k, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, s.stream);
        pend[si].off = off;
        pend[si].cnt = cnt;
        off += cnt;
    }
    // drain -- on the error paths too: no copy may still be writing the caller's buffers when
    // this function returns
    for (int si = 0; si < nslots; ++si) {
        if (ce == cudaSuccess && rc == MODCHOL_OK)
            ce = retire(si);
        else
            cudaStreamSynchronize(c.slots[si].stream);
    }
    for (int si = 0; si < kSlots; ++si)
        wait_outcopy(si);
    if (rc)
        return rc;
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "host staging pipeline: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int check_common(int algo, int mode, long long batch, int n, const void* a, const void* l,
                 const void* e, const void* p)
{
    if (algo < MODCHOL_GMW81 || algo > MODCHOL_SE99)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown algorithm %d", algo);
    if (mode != MODCHOL_MODE_STRICT && mode != MODCHOL_MODE_FAST)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mode %d", mode);
    if (batch < 0)
        return fail(MODCHOL_ERR_BAD_ARG, "negative batch %lld", batch);
    if (n < 1)  // the reference panics for n == 0 (`0..(n-1)` underflow, se99.rs:190)
        return fail(MODCHOL_ERR_BAD_ARG, "n must be >= 1 (got %d)", n);
    if (n > MODCHOL_MAX_N)
        return fail(MODCHOL_ERR_UNSUPPORTED, "n = %d exceeds MODCHOL_MAX_N = %d", n, MODCHOL_MAX_N);
    if (batch > 0 && (!a || !l || !e || !p))
        return fail(MODCHOL_ERR_BAD_ARG, "null buffer");
    return MODCHOL_OK;
}

// Factorise and solve with device pointers.  n <= 64 (shared-memory warp kernels): ONE kernel,
// the factor never leaves the SM unless l is asked for.  Otherwise:
// factor kernel into l / p (stream-ordered scratch when the caller passes NULL), then the batched
// solve kernel.
int factor_solve_device(int dev, int algo, int mode, long long batch, int n, int nrhs,
                        const double* a, const double* b, double* x, double* l, double* e,
                        unsigned long long* p, int* k, cudaStream_t st)
{
    if (batch == 0)
        return MODCHOL_OK;
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const DeviceCtx& c = g_ctx[dev];
    // every shared-memory warp class (n <= 64, both modes) carries its own fused solve
    // (the thread-per-matrix class carries its own solve for n <= 8; for 8 < n <= 16 its factor
    // kernel followed by the solve kernel is ahead of the fused warp kernels: SE99 n=16 STRICT 213
    // against 196 M systems/s)
    KernelClass cls = classify(algo, mode, n, c.smem_optin);
    if (cls == kClsThread && batch < thread_min_batch(n) && small_kernel_pref() != 3)
        cls = kClsWarpSmem;
    if (cls == kClsThread && n <= kTpmMaxNHost) {
        cudaError_t ce = launch_tpm_solve_kernel(algo, n, batch, a, l, e, p, k, b, x, nrhs, c.sms, st);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "fused factor+solve launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    const bool fused = cls == kClsWarpSmem;
    if (fused) {
        cudaError_t ce = launch_wsm_kernel(algo, mode, n, batch, a, l, e, p, k, c.sms, st, b, x, nrhs);
        g_launches.fetch_add(1);
        if (ce != cudaSuccess)
            return fail(MODCHOL_ERR_CUDA, "fused factor+solve launch failed: %s", cudaGetErrorString(ce));
        return MODCHOL_OK;
    }
    // Scratch for the outputs the caller did not ask for is stream-ordered and bounded: the batch
    // runs in chunks of at most kScratchBytes of L, so a call with l == NULL costs 512 MB of pool
    // memory whatever the batch (the pool keeps it between calls: ctx_init raises its release
    // threshold), and each chunk's factor and solve kernels follow each other on the stream.
    const size_t nn = (size_t)n * n;
    constexpr size_t kScratchBytes = (size_t)512 << 20;
    long long chunk = batch;
    if (!l)
        chunk = std::min<long long>(batch, std::max<long long>(1, (long long)(kScratchBytes / (nn * sizeof(double)))));
    double *lt = nullptr, *et = nullptr;
    unsigned long long* pt = nullptr;
    cudaError_t ce = cudaSuccess;
    if (!l) ce = cudaMallocAsync(&lt, (size_t)chunk * nn * sizeof(double), st);
    if (ce == cudaSuccess && !e) ce = cudaMallocAsync(&et, (size_t)chunk * n * sizeof(double), st);
    if (ce == cudaSuccess && !p) ce = cudaMallocAsync(&pt, (size_t)chunk * n * sizeof(unsigned long long), st);
    for (long long c0 = 0; ce == cudaSuccess && rc == MODCHOL_OK && c0 < batch; c0 += chunk) {
        const long long cb = std::min(chunk, batch - c0);
        double* lw = l ? l + (size_t)c0 * nn : lt;
        double* ew = e ? e + (size_t)c0 * n : et;
        unsigned long long* pw = p ? p + (size_t)c0 * n : pt;
        rc = factor_device(dev, algo, mode, cb, n, a + (size_t)c0 * nn, lw, ew, pw, k ? k + c0 : nullptr, st);
        if (rc == MODCHOL_OK) {
            const size_t off = (size_t)c0 * nrhs * n;
            ce = launch_solve(cb, n, nrhs, lw, pw, b + off, x + off, c.sms, st);
            g_launches.fetch_add(1);
        }
    }
    if (lt) cudaFreeAsync(lt, st);
    if (et) cudaFreeAsync(et, st);
    if (pt) cudaFreeAsync(pt, st);
    if (rc)
        return rc;
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "factor+solve: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

// Batched solve for n <= 64: the lower triangle of L is staged once in shared memory (packed, 32/G
// matrices per warp as in wsm_se.cuh) and every right-hand side runs the substitution of the fused
// kernel (wsm_solve_in_place: forward over the columns of L, backward over its rows, one broadcast
// + one LDS + one FMA per row and step) -- no shuffle reduction and no division inside a step.
template <int G, int R>
__global__ void __launch_bounds__(128)
solve_warp_kernel(int n, int wstride, long long batch, int nrhs, const double* __restrict__ L,
                  const unsigned long long* __restrict__ P, const double* B, double* X)   // X may alias B
{
    extern __shared__ double sm[];
    constexpr int kPack = 32 / G, kWarps = 4;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane & (G - 1), grp = lane / G;
    const unsigned mask = wsm_group_mask<G>(grp);
    const unsigned wb = (unsigned)__cvta_generic_to_shared(sm + (size_t)(wib * kPack + grp) * wstride);
    int row[R], perm[R];
    unsigned tib[R];
    bool rv[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        row[s] = sub + G * s;
        rv[s] = row[s] < n;
        tib[s] = wb + 8u * (unsigned)tri(rv[s] ? row[s] : 0);
    }
    const size_t nn = (size_t)n * n;
    const long long nslots = (long long)gridDim.x * kWarps * kPack;
    for (long long b = ((long long)blockIdx.x * kWarps + wib) * kPack + grp; b < batch; b += nslots) {
        __syncwarp(mask);
        wsm_load_lower<G, R, false>(L + (size_t)b * nn, n, wb, sub);
        __syncwarp(mask);
#pragma unroll
        for (int s = 0; s < R; ++s)
            perm[s] = rv[s] ? (int)P[(size_t)b * n + row[s]] : 0;
        wsm_solve_in_place<G, R>(mask, n, wb, tib, row, rv, perm, B + (size_t)b * nrhs * n,
                                 X + (size_t)b * nrhs * n, nrhs);
    }
}

template <int G, int R>
cudaError_t launch_solve_warp(int n, long long batch, int nrhs, const double* l, const unsigned long long* p,
                              const double* b, double* x, int sms, cudaStream_t st)
{
    auto kern = solve_warp_kernel<G, R>;
    constexpr int kPack = 32 / G;
    const int wstride = wsm_warp_doubles(n);
    const size_t smem = (size_t)4 * kPack * wstride * sizeof(double);
    const size_t smem_max = (size_t)4 * kPack * wsm_warp_doubles(G * R) * sizeof(double);
    int occ = 1;
    cudaError_t ce = cached_occupancy(kern, 128, smem, smem_max, &occ);
    if (ce != cudaSuccess) return ce;
    long long blocks = (batch + 4 * kPack - 1) / (4 * kPack);
    const long long resident = (long long)sms * occ;
    if (blocks > resident) blocks = resident;
    kern<<<(unsigned)blocks, 128, smem, st>>>(n, wstride, batch, nrhs, l, p, b, x);
    return cudaGetLastError();
}

// n <= 64: the shared-memory warp kernel; larger n: one warp per system on rows of L in global memory
cudaError_t launch_solve(long long batch, int n, int nrhs, const double* l, const unsigned long long* p,
                         const double* b, double* x, int sms, cudaStream_t st)
{
    if (n <= 4) return launch_solve_warp<4, 1>(n, batch, nrhs, l, p, b, x, sms, st);
    if (n <= 8) return launch_solve_warp<8, 1>(n, batch, nrhs, l, p, b, x, sms, st);
    if (n <= 16) return launch_solve_warp<16, 1>(n, batch, nrhs, l, p, b, x, sms, st);
    if (n <= 32) return launch_solve_warp<32, 1>(n, batch, nrhs, l, p, b, x, sms, st);
    if (n <= 64) return launch_solve_warp<32, 2>(n, batch, nrhs, l, p, b, x, sms, st);
    const size_t smem = (size_t)4 * n * sizeof(double);
    const long long total = batch * (long long)nrhs;
    const unsigned grid = (unsigned)std::min<long long>((total + 3) / 4, (long long)sms * 16);
    solve_kernel<<<grid, 128, smem, st>>>(batch, n, nrhs, l, p, b, x);
    return cudaGetLastError();
}

// n <= 64: one warp per matrix, L L^T on the FP64 tensor pipe; larger n: one CTA per matrix
void launch_verify(long long batch, int n, const double* a, const double* l, const double* e,
                   const unsigned long long* p, double* resid, int* flags, int sms, cudaStream_t st)
{
    if (n <= 64) {
        const unsigned grid = (unsigned)std::min<long long>((batch + 3) / 4, (long long)sms * 16);
        if (n <= 16)
            verify_warp_kernel<2><<<grid, 128, 0, st>>>(batch, n, a, l, e, p, resid, flags);
        else if (n <= 32)
            verify_warp_kernel<4><<<grid, 128, 0, st>>>(batch, n, a, l, e, p, resid, flags);
        else if (n <= 48)
            verify_warp_kernel<6><<<grid, 128, 0, st>>>(batch, n, a, l, e, p, resid, flags);
        else
            verify_warp_kernel<8><<<grid, 128, 0, st>>>(batch, n, a, l, e, p, resid, flags);
        return;
    }
    const unsigned grid = (unsigned)std::min<long long>(batch, (long long)sms * 8);
    verify_kernel<<<grid, 256, 0, st>>>(batch, n, a, l, e, p, resid, flags);
}

}  // namespace

// =========================================================================================
extern "C" {

int modchol_version(void) { return MODCHOL_VERSION; }

const char* modchol_last_error(void) { return g_err.c_str(); }

int modchol_device_count(void) { return device_count(); }

int64_t modchol_launch_count(void) { return g_launches.load(); }

int modchol_set_small_kernel_family(int family)
{
    const int prev = small_kernel_pref();
    g_small_family.store(family >= 1 && family <= 3 ? family : 0);
    return prev;
}

const char* modchol_kernel_class(int algo, int mode, int32_t n)
{
    if (n < 1 || n > MODCHOL_MAX_N || algo < 0 || algo > 2)
        return "unsupported";
    switch (classify(algo, mode, n, 0)) {
    case kClsWarp: return "warp32";
    case kClsThread: return "thread";
    case kClsWarpSmem: return "warp_smem";
    case kClsCtaPacked: return "cta_packed";
    case kClsCtaBlocked: return "cta_blocked";
    case kClsWarpPanel: return "warp_panel";
    case kClsCtaSmem: return "cta_smem";
    default: return "cta_gmem";
    }
}

void modchol_shutdown(void)
{
    const int nd = std::min(device_count(), kMaxDevices);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < nd; ++d) {
        DeviceCtx& c = g_ctx[d];
        std::lock_guard<std::mutex> lock(c.mu);
        bool any = false;
        for (auto& s : c.slots)
            any = any || s.stream || s.a;
        if (!any)
            continue;
        cudaSetDevice(d);
        for (auto& s : c.slots) {
            if (s.stream) {
                cudaStreamSynchronize(s.stream);
                slot_free(s);
                slot_free_host(s);
                cudaStreamDestroy(s.stream);
                s.stream = nullptr;
            }
        }
    }
    cudaSetDevice(cur);
}

int modchol_factor_batched_f64(int algo, int mode, int64_t batch, int32_t n, const double* a,
                               double* l, double* e, uint64_t* p, int32_t* phase2_start,
                               int mem_kind, void* stream)
{
    g_err.clear();
    int rc = check_common(algo, mode, batch, n, a, l, e, p);
    if (rc)
        return rc;
    if (mem_kind != MODCHOL_MEM_HOST && mem_kind != MODCHOL_MEM_DEVICE)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= kMaxDevices)
        return fail(MODCHOL_ERR_UNSUPPORTED, "device index %d too large", dev);
    if (mem_kind == MODCHOL_MEM_DEVICE)
        return factor_device(dev, algo, mode, batch, n, a, l, e,
                             reinterpret_cast<unsigned long long*>(p), phase2_start,
                             static_cast<cudaStream_t>(stream));
    return factor_host_on_device(dev, algo, mode, batch, n, a, l, e,
                                 reinterpret_cast<unsigned long long*>(p), phase2_start);
}

int modchol_factor_batched_multi_gpu_f64(int algo, int mode, int64_t batch, int32_t n,
                                         const double* a, double* l, double* e, uint64_t* p,
                                         int32_t* phase2_start, int ngpus)
{
    g_err.clear();
    int rc = check_common(algo, mode, batch, n, a, l, e, p);
    if (rc)
        return rc;
    const int nd = std::min(device_count(), kMaxDevices);
    if (nd < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (ngpus <= 0)
        ngpus = nd;
    if (ngpus > nd)
        return fail(MODCHOL_ERR_BAD_ARG, "ngpus = %d but only %d devices are visible", ngpus, nd);
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    // contiguous slices: device g gets [g*per, min(batch, (g+1)*per))
    const long long per = (batch + ngpus - 1) / ngpus;
    std::vector<int> rcs(ngpus, MODCHOL_OK);
    std::vector<std::string> errs(ngpus);
    std::vector<std::thread> th;
    const size_t nn = (size_t)n * n;
    for (int g = 0; g < ngpus; ++g) {
        const long long lo = std::min<long long>(batch, (long long)g * per);
        const long long hi = std::min<long long>(batch, lo + per);
        if (hi <= lo)
            continue;
        th.emplace_back([=, &rcs, &errs]() {
            rcs[g] = factor_host_on_device(
                g, algo, mode, hi - lo, n, a + (size_t)lo * nn, l + (size_t)lo * nn,
                e + (size_t)lo * n, reinterpret_cast<unsigned long long*>(p) + (size_t)lo * n,
                phase2_start ? phase2_start + lo : nullptr);
            if (rcs[g])
                errs[g] = g_err;  // thread-local of the worker
        });
    }
    for (auto& t : th)
        t.join();
    cudaSetDevice(cur);
    for (int g = 0; g < ngpus; ++g)
        if (rcs[g])
            return fail(rcs[g], "device %d: %s", g, errs[g].c_str());
    return MODCHOL_OK;
}

int modchol_gershgorin_circles_batched_f64(int64_t batch, int32_t n, const double* a,
                                           double* circles, int mem_kind, void* stream)
{
    g_err.clear();
    if (batch < 0 || n < 1 || (batch > 0 && (!a || !circles)))
        return fail(MODCHOL_ERR_BAD_ARG, "bad argument");
    if (n > MODCHOL_MAX_N)
        return fail(MODCHOL_ERR_UNSUPPORTED, "n too large");
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (batch == 0)
        return MODCHOL_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const size_t nn = (size_t)n * n;
    const long long total = batch * (long long)n;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)g_ctx[dev].sms * 16);
    if (mem_kind == MODCHOL_MEM_DEVICE) {
        gershgorin_circles_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(batch, n, a, circles);
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        return MODCHOL_OK;
    }
    if (mem_kind != MODCHOL_MEM_HOST)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    double *da = nullptr, *dc = nullptr;
    CUDA_TRY(cudaMalloc(&da, batch * nn * sizeof(double)));
    cudaError_t ce = cudaMalloc(&dc, (size_t)total * 2 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpy(da, a, batch * nn * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        gershgorin_circles_kernel<<<grid, 256>>>(batch, n, da, dc);
        g_launches.fetch_add(1);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(circles, dc, (size_t)total * 2 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(da);
    cudaFree(dc);
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "gershgorin: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int modchol_verify_batched_f64(int64_t batch, int32_t n, const double* a, const double* l,
                               const double* e, const uint64_t* p, double* resid, int32_t* flags,
                               int mem_kind, void* stream)
{
    g_err.clear();
    if (batch < 0 || n < 1 || (batch > 0 && (!a || !l || !e || !p)))
        return fail(MODCHOL_ERR_BAD_ARG, "bad argument");
    if (n > MODCHOL_MAX_N)
        return fail(MODCHOL_ERR_UNSUPPORTED, "n too large");
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (batch == 0)
        return MODCHOL_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    auto pu = reinterpret_cast<const unsigned long long*>(p);
    if (mem_kind == MODCHOL_MEM_DEVICE) {
        launch_verify(batch, n, a, l, e, pu, resid, flags, g_ctx[dev].sms, static_cast<cudaStream_t>(stream));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        return MODCHOL_OK;
    }
    if (mem_kind != MODCHOL_MEM_HOST)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    const size_t nn = (size_t)n * n;
    double *da = nullptr, *dl = nullptr, *de = nullptr, *dr = nullptr;
    unsigned long long* dp = nullptr;
    int* df = nullptr;
    cudaError_t ce = cudaMalloc(&da, batch * nn * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&dl, batch * nn * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&de, (size_t)batch * n * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&dp, (size_t)batch * n * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&dr, (size_t)batch * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&df, (size_t)batch * 4);
    if (ce == cudaSuccess) ce = cudaMemcpy(da, a, batch * nn * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(dl, l, batch * nn * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(de, e, (size_t)batch * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(dp, p, (size_t)batch * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        launch_verify(batch, n, da, dl, de, dp, dr, df, g_ctx[dev].sms, nullptr);
        g_launches.fetch_add(1);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && resid) ce = cudaMemcpy(resid, dr, (size_t)batch * 8, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && flags) ce = cudaMemcpy(flags, df, (size_t)batch * 4, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(dl); cudaFree(de); cudaFree(dp); cudaFree(dr); cudaFree(df);
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "verify: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int modchol_solve_batched_f64(int64_t batch, int32_t n, int32_t nrhs, const double* l,
                              const uint64_t* p, const double* b, double* x, int mem_kind,
                              void* stream)
{
    g_err.clear();
    if (batch < 0 || n < 1 || nrhs < 1 || (batch > 0 && (!l || !p || !b || !x)))
        return fail(MODCHOL_ERR_BAD_ARG, "bad argument");
    if (n > MODCHOL_MAX_N)
        return fail(MODCHOL_ERR_UNSUPPORTED, "n too large");
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (batch == 0)
        return MODCHOL_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const size_t smem = (size_t)4 * n * sizeof(double);
    const long long total = batch * (long long)nrhs;
    const unsigned grid = (unsigned)std::min<long long>((total + 3) / 4, (long long)g_ctx[dev].sms * 16);
    auto pu = reinterpret_cast<const unsigned long long*>(p);
    if (mem_kind == MODCHOL_MEM_DEVICE) {
        if (b == x) {
            // x[p_i] = w_i is a scatter: an aliased rhs would be overwritten while still needed by
            // another lane only within one system, and each system is staged in shared memory
            // before any write -- aliasing is safe.
        }
        const cudaError_t ce = launch_solve(batch, n, nrhs, l, pu, b, x, g_ctx[dev].sms, static_cast<cudaStream_t>(stream));
        g_launches.fetch_add(1);
        CUDA_TRY(ce);
        return MODCHOL_OK;
    }
    if (mem_kind != MODCHOL_MEM_HOST)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    const size_t nn = (size_t)n * n;
    double *dl = nullptr, *db = nullptr;
    unsigned long long* dp = nullptr;
    cudaError_t ce = cudaMalloc(&dl, batch * nn * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&dp, (size_t)batch * n * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&db, (size_t)total * n * 8);
    if (ce == cudaSuccess) ce = cudaMemcpy(dl, l, batch * nn * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(dp, p, (size_t)batch * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(db, b, (size_t)total * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        ce = launch_solve(batch, n, nrhs, dl, dp, db, db, g_ctx[dev].sms, nullptr);
        g_launches.fetch_add(1);
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(x, db, (size_t)total * n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dl); cudaFree(dp); cudaFree(db);
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "solve: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int modchol_factor_solve_batched_f64(int algo, int mode, int64_t batch, int32_t n, int32_t nrhs,
                                     const double* a, const double* b, double* x, double* l,
                                     double* e, uint64_t* p, int32_t* phase2_start, int mem_kind,
                                     void* stream)
{
    g_err.clear();
    int rc = check_common(algo, mode, batch, n, a, a, a, a);   // l, e, p are optional here
    if (rc)
        return rc;
    if (nrhs < 1 || (batch > 0 && (!b || !x)))
        return fail(MODCHOL_ERR_BAD_ARG, "factor_solve: nrhs >= 1 and non-null b, x required");
    if (mem_kind != MODCHOL_MEM_HOST && mem_kind != MODCHOL_MEM_DEVICE)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (batch == 0)
        return MODCHOL_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= kMaxDevices)
        return fail(MODCHOL_ERR_UNSUPPORTED, "device index %d too large", dev);
    auto pu = reinterpret_cast<unsigned long long*>(p);
    if (mem_kind == MODCHOL_MEM_DEVICE)
        return factor_solve_device(dev, algo, mode, batch, n, nrhs, a, b, x, l, e, pu, phase2_start,
                                   static_cast<cudaStream_t>(stream));
    // host pointers: chunks of at most ~256 MB of input through plain device buffers
    const size_t nn = (size_t)n * n;
    const long long per = std::max<long long>(1, (long long)((256ull << 20) / (nn * sizeof(double))));
    const long long cmax = std::min<long long>(per, batch);
    const size_t rb = (size_t)nrhs * n * sizeof(double);
    double *da = nullptr, *db = nullptr, *dl = nullptr, *de = nullptr;
    unsigned long long* dp = nullptr;
    int* dk = nullptr;
    cudaError_t ce = cudaMalloc(&da, cmax * nn * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc(&db, cmax * rb);
    if (ce == cudaSuccess && l) ce = cudaMalloc(&dl, cmax * nn * sizeof(double));
    if (ce == cudaSuccess && e) ce = cudaMalloc(&de, cmax * n * sizeof(double));
    if (ce == cudaSuccess && p) ce = cudaMalloc(&dp, cmax * n * sizeof(unsigned long long));
    if (ce == cudaSuccess && phase2_start) ce = cudaMalloc(&dk, cmax * sizeof(int));
    rc = MODCHOL_OK;
    for (long long o = 0; ce == cudaSuccess && rc == MODCHOL_OK && o < batch; o += cmax) {
        const long long cnt = std::min<long long>(cmax, batch - o);
        ce = cudaMemcpy(da, a + o * nn, cnt * nn * sizeof(double), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess)
            ce = cudaMemcpy(db, b + (size_t)o * nrhs * n, cnt * rb, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) break;
        rc = factor_solve_device(dev, algo, mode, cnt, n, nrhs, da, db, db, dl, de, dp, dk, nullptr);
        if (rc) break;
        ce = cudaMemcpy(x + (size_t)o * nrhs * n, db, cnt * rb, cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && l) ce = cudaMemcpy(l + o * nn, dl, cnt * nn * sizeof(double), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && e) ce = cudaMemcpy(e + o * n, de, cnt * n * sizeof(double), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && p) ce = cudaMemcpy(p + o * n, dp, cnt * n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess && phase2_start)
            ce = cudaMemcpy(phase2_start + o, dk, cnt * sizeof(int), cudaMemcpyDeviceToHost);
    }
    cudaFree(da); cudaFree(db); cudaFree(dl); cudaFree(de); cudaFree(dp); cudaFree(dk);
    if (rc)
        return rc;
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "factor_solve: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int modchol_assemble_qdqt_batched_f64(int64_t batch, int32_t n, int32_t nrefl, const double* v,
                                      const double* d, double* a, int mem_kind, void* stream)
{
    g_err.clear();
    if (batch < 0 || n < 1 || nrefl < 0 || (batch > 0 && (!d || !a || (nrefl > 0 && !v))))
        return fail(MODCHOL_ERR_BAD_ARG, "bad argument");
    if (n > MODCHOL_MAX_N)
        return fail(MODCHOL_ERR_UNSUPPORTED, "n too large");
    if (mem_kind != MODCHOL_MEM_HOST && mem_kind != MODCHOL_MEM_DEVICE)
        return fail(MODCHOL_ERR_BAD_ARG, "unknown mem_kind %d", mem_kind);
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    if (batch == 0)
        return MODCHOL_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const size_t smem = (size_t)(3 * n + 66) * sizeof(double);   // <= 48.5 KB at n = 2048
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
        cudaFuncSetAttribute(assemble_qdqt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)((3 * MODCHOL_MAX_N + 66) * sizeof(double)));
    });
    const unsigned grid = (unsigned)std::min<long long>(batch, (long long)g_ctx[dev].sms * 8);
    if (mem_kind == MODCHOL_MEM_DEVICE) {
        assemble_qdqt_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(batch, n, nrefl, v, d, a);
        g_launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        return MODCHOL_OK;
    }
    const size_t nn = (size_t)n * n;
    double *dv = nullptr, *ddg = nullptr, *da = nullptr;
    cudaError_t ce = cudaMalloc(&da, batch * nn * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&ddg, (size_t)batch * n * 8);
    if (ce == cudaSuccess && nrefl > 0) ce = cudaMalloc(&dv, (size_t)batch * nrefl * n * 8);
    if (ce == cudaSuccess) ce = cudaMemcpy(ddg, d, (size_t)batch * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && nrefl > 0) ce = cudaMemcpy(dv, v, (size_t)batch * nrefl * n * 8, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        assemble_qdqt_kernel<<<grid, 256, smem>>>(batch, n, nrefl, dv, ddg, da);
        g_launches.fetch_add(1);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpy(a, da, batch * nn * 8, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(ddg); cudaFree(dv);
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "assemble_qdqt: %s", cudaGetErrorString(ce));
    return MODCHOL_OK;
}

int modchol_fp64_peak_tflops(double* dfma_tflops, double* dmma_tflops)
{
    g_err.clear();
    if (device_count() < 1)
        return fail(MODCHOL_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ctx_init(dev);
    if (rc)
        return rc;
    const int sms = g_ctx[dev].sms, iters = 1 << 14, threads = 512, blocks = sms * 4;
    double* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {   // rep 0 warms up
            cudaEventRecord(e0);
            if (which == 0)
                fp64_dfma_probe_kernel<<<blocks, threads>>>(iters, 0.999999, sink);
            else
                fp64_dmma_probe_kernel<<<blocks, threads>>>(iters, 0.999999, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            g_launches.fetch_add(1);
            // DFMA: 8 per thread and iteration, 2 flops each; DMMA: 4 per warp, 8x8x4x2 flops each
            const double flops = which == 0 ? (double)blocks * threads * iters * 8.0 * 2.0
                                            : (double)blocks * (threads / 32) * iters * 4.0 * 512.0;
            if (rep > 0 && ms > 0.f)
                best[which] = std::max(best[which], flops / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t ce = cudaGetLastError();
    cudaFree(sink);
    if (ce != cudaSuccess)
        return fail(MODCHOL_ERR_CUDA, "fp64 probe: %s", cudaGetErrorString(ce));
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return MODCHOL_OK;
}

}  // extern "C"
