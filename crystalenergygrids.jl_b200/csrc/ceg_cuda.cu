// libceg_cuda — C ABI implementation (include/ceg_cuda.h).  sm_100a only; no CPU fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>

#include "common.cuh"
#include "ewald.cuh"
#include "grid_build.cuh"
#include "guest.cuh"
#include "interp.cuh"
#include "latency.cuh"

namespace ceg {

std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int fail(const std::string& msg) {
    last_error() = msg;
    return 1;
}
int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    char buf[768];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear sticky-less error state
    return fail(buf);
}

// ---- device allocation cache (common.cuh) --------------------------------------------------------
namespace {
struct PoolEntry { void* p; size_t bytes; int device; };
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool;
constexpr size_t POOL_MIN = 4096;          // smaller buffers go straight to cudaMalloc/cudaFree
constexpr size_t POOL_CAP = 4ull << 30;      // at most 4 GiB parked per process
size_t g_pool_bytes = 0;
}  // namespace
void* pool_acquire(size_t bytes, size_t* got) {
    if (bytes < POOL_MIN) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); i++)
        if (g_pool[i].device == dev && g_pool[i].bytes >= bytes && g_pool[i].bytes <= 2 * bytes &&
            (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
            best = i;
    if (best < 0) return nullptr;
    void* p = g_pool[best].p;
    *got = g_pool[best].bytes;
    g_pool_bytes -= g_pool[best].bytes;
    g_pool.erase(g_pool.begin() + best);
    return p;
}
void pool_release(void* p, size_t bytes) {
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (bytes >= POOL_MIN && g_pool_bytes + bytes <= POOL_CAP) {
            g_pool.push_back({p, bytes, dev});
            g_pool_bytes += bytes;
            return;
        }
    }
    cudaFree(p);
}
void pool_trim() {
    std::vector<PoolEntry> old;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        old.swap(g_pool);
        g_pool_bytes = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& e : old) { cudaSetDevice(e.device); cudaFree(e.p); }
    cudaSetDevice(cur);
}

namespace {
std::mutex g_user_mu;
std::vector<PoolEntry> g_user_bufs;     // buffers handed out by ceg_device_malloc
}  // namespace
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int use_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(std::string("no usable CUDA device (libceg_cuda has no CPU fallback): ") + cudaGetErrorString(e));
    if (dev < 0 || dev >= n) return fail("device ordinal out of range");
    CEG_CK(cudaSetDevice(dev));
    return 0;
}

// Streams are recycled across handles (creating and destroying a pair per ceg_probe_create / ceg_probe_free cycle is
// part of the fixed cost of a create_grid_* call).
namespace {
std::mutex g_stream_mu;
std::vector<std::pair<int, cudaStream_t>> g_streams;
}  // namespace
cudaError_t acquire_stream(cudaStream_t* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_stream_mu);
        for (size_t i = 0; i < g_streams.size(); i++)
            if (g_streams[i].first == dev) {
                *out = g_streams[i].second;
                g_streams.erase(g_streams.begin() + i);
                return cudaSuccess;
            }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
void release_stream(cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_stream_mu);
    if (g_streams.size() < 64) g_streams.push_back({dev, s});
    else cudaStreamDestroy(s);
}

// multiprocessor count per device (launch sizes are multiples of it)
int sm_count_of(int dev) {
    static std::mutex mu;
    static std::vector<int> cache;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)cache.size() <= dev) cache.resize(dev + 1, 0);
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        cache[dev] = n;
    }
    return cache[dev];
}
static int sm_count_here() {
    int dev = 0;
    cudaGetDevice(&dev);
    return sm_count_of(dev);
}

}  // namespace ceg

using namespace ceg;

// ================================================================================================
// handles
// ================================================================================================
struct ceg_probe_s {
    int device = 0;
    std::mutex mu;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaStream_t s_compute2 = nullptr;      // odd chunks of ceg_grid_build: the tail of one chunk's launch overlaps the head of the next
    cudaEvent_t ev_join = nullptr;
    CellDev cell;
    double cutoff2 = 0, alpha = 0;
    int n_probes = 0, want_coulomb = 0, n_types = 0;
    int64_t n_atoms = 0;
    std::vector<double> h_xyz, h_q;
    std::vector<int> h_type;
    double max_hs_r2 = 0;
    double min_perp_width = 0;
    DevBuf d_xyz, d_type, d_q, d_table, d_stats;
    // cached expanded image list
    // cached expanded image lists, one per (slab, grid spec) recently built — callers alternate between a few slabs (the
    // halves of a pipelined build, C1 and C4 of one framework): least recently used of EXT_CACHE entries is replaced
    struct ExtList {
        DevBuf d_ext;                      // [images | meta | per-plane ranges]
        size_t off_meta = 0, off_range = 0;
        int n_ext = 0, plane0 = 0, planes = 0;
        double key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int key_planes[2] = {-1, -1};
        bool valid = false;
        unsigned long long used = 0;
    };
    static constexpr int EXT_CACHE = 4;
    ExtList ext_cache[EXT_CACHE];
    unsigned long long ext_clock = 0;
    int sm_count = 148;
    // events of the write-back ring of ceg_grid_build, created once per handle
    cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr}, ev_free[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_time[3] = {nullptr, nullptr, nullptr};   // first launch, last launch done, last copy done
    double last_timing[4] = {0, 0, 0, 0};                   // image list (host) ms, kernel span ms, copy tail ms, whole call ms
    double list_ms = 0.0;                                   // host time spent (re)building image lists in the current call
    // last slab whose image list was too long to be scanned by every tile: remembered so that repeated builds of
    // the same slab go straight to its x windows (list_windows)
    int split_i0 = -1, split_i1 = -1, split_dims[3] = {0, 0, 0};
    // scratch for host-facing builds
    DevBuf d_out, d_raw, d_pos, d_der;
    int64_t last_points = 0;
    ~ceg_probe_s() {
        for (int r = 0; r < 3; r++) {
            if (ev_done[r]) cudaEventDestroy(ev_done[r]);
            if (ev_free[r]) cudaEventDestroy(ev_free[r]);
            if (ev_time[r]) cudaEventDestroy(ev_time[r]);
        }
        if (ev_join) cudaEventDestroy(ev_join);
        if (s_compute) release_stream(s_compute);
        if (s_compute2) release_stream(s_compute2);
        if (s_copy) release_stream(s_copy);
    }
};

struct ceg_gridpair_s;
struct ceg_grid_s {
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // ping-pong pair for the pipelined host path
    GridDev g;
    DevBuf d_nodes, d_xyz[2], d_out[2];
    ceg_gridpair_s* pair = nullptr;                     // fused {this VdW grid, a Coulomb grid} records (ceg_grid_pair_create)
    ~ceg_grid_s() {
        if (stream) release_stream(stream);
        if (stream2) release_stream(stream2);
    }
};

struct ceg_gridpair_s {
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    PairGridDev g;
    ceg_grid_s* vdw = nullptr;
    ceg_grid_s* coul = nullptr;
    DevBuf d_nodes, d_xyz[2], d_out[2];
    ~ceg_gridpair_s() {
        if (stream) release_stream(stream);
        if (stream2) release_stream(stream2);
    }
};

struct ceg_ewald_s {
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    double invmat[9];
    int kmax[3];
    int64_t n_k = 0;
    DevBuf d_kv, d_xyz, d_out, d_q;
    // thread-per-placement kernel: rows of the k-list, kfactor and kfactor * S_f in list order
    bool rows_ok = false;
    double sum_kf = 0.0;
    DevBuf d_rows, d_kf, d_ksf, d_ksf_rest;
    ~ceg_ewald_s() {
        if (stream) cudaStreamDestroy(stream);
    }
};

struct ceg_block_s {
    int device = 0;
    BlockDev b;
    DevBuf d_mask;
};

// ================================================================================================
// misc
// ================================================================================================
extern "C" const char* ceg_last_error(void) { return last_error().c_str(); }
extern "C" int ceg_abi_version(void) { return CEG_ABI_VERSION; }

extern "C" int ceg_device_count(int32_t* n) {
    if (!n) return fail("null argument");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *n = c;
    return 0;
}

extern "C" int ceg_device_info(int32_t device, char* name256, int32_t* sm_count, int64_t* mem_bytes, int32_t* clock_khz) {
    if (use_device(device)) return 1;
    cudaDeviceProp p;
    CEG_CK(cudaGetDeviceProperties(&p, device));
    if (name256) { strncpy(name256, p.name, 255); name256[255] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (int64_t)p.totalGlobalMem;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *clock_khz = khz;
    }
    return 0;
}

extern "C" int ceg_trim_device_cache(void) {
    pool_trim();
    return 0;
}

extern "C" int ceg_kernel_launch_count(int64_t* n) {
    if (!n) return fail("null argument");
    *n = (int64_t)g_launches.load();
    return 0;
}

namespace ceg {
// 8 independent FMA chains per thread; the result is stored so the chains cannot be removed.
// REGOPS = false: x = fma(x, a, b) with uniform a, b (one register operand per DFMA);
// REGOPS = true : x = fma(x, y_i, z_i) with per-thread registers (three register operands per DFMA),
// which is what a real kernel mostly issues.  On B200 the second form peaks ~30 % lower (register
// read bandwidth), so both are reported.
template <bool REGOPS>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = threadIdx.x + i;
        y[i] = REGOPS ? 0.999999 + 1e-9 * threadIdx.x * (i + 1) : a;
        z[i] = REGOPS ? 1e-9 * (threadIdx.x + i) : b;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], y[i], z[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
}

template <bool REGOPS>
static int measure_fp64(int device, double* tflops) {
    cudaDeviceProp p;
    CEG_CK(cudaGetDeviceProperties(&p, device));
    const int blocks = p.multiProcessorCount * 2, threads = 256, iters = 4096;
    DevBuf buf;
    CEG_CK(buf.reserve(sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CEG_CK(cudaEventCreate(&e0));
    CEG_CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CEG_CK(cudaEventRecord(e0, 0));
        fp64_peak_kernel<REGOPS><<<blocks, threads>>>(buf.as<double>(), iters, 0.999999, 1e-9);
        count_launch();
        CEG_CK(cudaGetLastError());
        CEG_CK(cudaEventRecord(e1, 0));
        CEG_CK(cudaEventSynchronize(e1));
        float ms = 0;
        CEG_CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}
}  // namespace ceg

extern "C" int ceg_measure_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return fail("null argument");
    if (use_device(device)) return 1;
    return measure_fp64<false>(device, tflops);
}

extern "C" int ceg_measure_fp64_peak_regops(int32_t device, double* tflops) {
    if (!tflops) return fail("null argument");
    if (use_device(device)) return 1;
    return measure_fp64<true>(device, tflops);
}

namespace ceg {
__global__ void selftest_math_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[6 * i] = exp_fast(-v);
    out[6 * i + 1] = erfcx_fast(v);
    out[6 * i + 2] = rsqrt_fast(v);
    out[6 * i + 3] = rcp_fast(v);
    out[6 * i + 4] = exp_acc(-v);
    out[6 * i + 5] = exp_fast(v);
}
}  // namespace ceg

extern "C" int ceg_selftest_math(int32_t device, int64_t n, const double* x, double* out) {
    if (n < 0 || (n > 0 && (!x || !out))) return fail("ceg_selftest_math: bad argument");
    if (n == 0) return 0;
    if (use_device(device)) return 1;
    DevBuf dx, dout;
    CEG_CK(dx.reserve(sizeof(double) * n));
    CEG_CK(dout.reserve(sizeof(double) * 6 * n));
    CEG_CK(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    selftest_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, dx.as<double>(), dout.as<double>());
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaMemcpy(out, dout.p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int ceg_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail("ceg_host_register: bad argument");
    CEG_CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
extern "C" int ceg_host_unregister(void* ptr) {
    CEG_CK(cudaHostUnregister(ptr));
    return 0;
}

// ================================================================================================
// probes / grid construction
// ================================================================================================
static double perp_width_min(const double* m) {
    auto cross = [](const double* u, const double* v, double* o) {
        o[0] = u[1] * v[2] - u[2] * v[1]; o[1] = u[2] * v[0] - u[0] * v[2]; o[2] = u[0] * v[1] - u[1] * v[0];
    };
    const double *a = m, *b = m + 3, *c = m + 6;
    double bxc[3], cxa[3], axb[3];
    cross(b, c, bxc); cross(c, a, cxa); cross(a, b, axb);
    double vol = std::fabs(a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2]);
    auto nrm = [](const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    return std::min(vol / nrm(bxc), std::min(vol / nrm(cxa), vol / nrm(axb)));
}

extern "C" int ceg_probe_create(int32_t device, const ceg_cell_t* supercell, int64_t n_atoms, const double* xyz,
                                const int32_t* type, const double* charge, int32_t n_types, int32_t n_probes,
                                const ceg_pair_t* table, double cutoff2, int32_t want_coulomb, double alpha,
                                ceg_probe_h* out) {
    if (!supercell || !out || n_atoms < 0 || (n_atoms > 0 && (!xyz || !type))) return fail("ceg_probe_create: null argument");
    if (n_probes < 0 || n_probes > MAX_GRIDS - 1) return fail("ceg_probe_create: at most 4 fused VdW probes per handle");
    if (n_probes + (want_coulomb ? 1 : 0) == 0) return fail("ceg_probe_create: nothing to compute");
    if (want_coulomb && !charge && n_atoms > 0) return fail("ceg_probe_create: charges required for the Coulomb grid");
    if (n_probes > 0 && (!table || n_types <= 0)) return fail("ceg_probe_create: pair table required");
    if (!(cutoff2 > 0) || !std::isfinite(cutoff2)) return fail("ceg_probe_create: cutoff must be finite and positive");
    for (int64_t i = 0; i < n_atoms; i++)
        if (n_probes > 0 && (type[i] < 0 || type[i] >= n_types)) return fail("ceg_probe_create: atom type out of range");
    for (int p = 0; p < n_probes; p++)
        for (int64_t i = 0; i < n_atoms; i++)
            if (table[(size_t)p * n_types + type[i]].flags < 0) return fail("Undefined interaction");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_probe_s> h(new ceg_probe_s);
    h->device = device;
    h->cell = to_dev(*supercell);
    h->cutoff2 = cutoff2;
    h->alpha = alpha;
    h->n_probes = n_probes;
    h->want_coulomb = want_coulomb ? 1 : 0;
    h->n_atoms = n_atoms;
    h->h_xyz.assign(xyz, xyz + 3 * n_atoms);
    if (want_coulomb) h->h_q.assign(charge, charge + n_atoms); else h->h_q.assign(n_atoms, 0.0);
    h->min_perp_width = perp_width_min(supercell->mat);
    // dense re-indexing of the species actually present in the framework: the kernels stage the
    // (probe x present species) pair table in shared memory
    std::vector<int> dense(std::max(1, n_types), -1), present;
    h->h_type.assign(n_atoms, 0);
    for (int64_t i = 0; i < n_atoms; i++) {
        const int t = n_probes > 0 ? type[i] : 0;
        if (dense[t] < 0) { dense[t] = (int)present.size(); present.push_back(t); }
        h->h_type[i] = dense[t];
    }
    const int n_present = std::max<int>(1, (int)present.size());
    // the kernels stage the (probe x present species) table in shared memory; 64 KB hold 744 entries (round 1
    // stopped at 24 species)
    if (n_probes > 0 && build_table_bytes(n_probes, n_present) > TABLE_SMEM_MAX)
        return fail("ceg_probe_create: the (probe x framework species) pair table exceeds 64 KB of shared memory: fuse fewer "
                    "probes per handle");
    h->n_types = n_present;
    h->sm_count = sm_count_of(device);
    std::vector<PairDev> tab((size_t)std::max(1, n_probes * n_present));
    memset(tab.data(), 0, sizeof(PairDev) * tab.size());
    for (int p = 0; p < n_probes; p++)
        for (int j = 0; j < (int)present.size(); j++) {
            const ceg_pair_t& s = table[(size_t)p * n_types + present[j]];
            PairDev& d = tab[(size_t)p * n_present + j];
            d.flags = s.flags; d._pad = 0;
            d.hs_r2 = s.hs_r * s.hs_r;
            d.lj_eps = s.lj_eps; d.lj_sig2 = s.lj_sigma * s.lj_sigma; d.lj_shift = s.lj_shift;
            { const double s6 = d.lj_sig2 * d.lj_sig2 * d.lj_sig2; d.lj_c6 = 4.0 * s.lj_eps * s6; d.lj_c12 = d.lj_c6 * s6; }
            d.bA = s.buck_A; d.bB = s.buck_B; d.bC = s.buck_C; d.b_shift = s.buck_shift;
            d.bB2 = s.buck_B * s.buck_B;
            d.b_lnA = 0.0;
            if ((s.flags & CEG_PAIR_BUCK) && s.buck_A > 0.0 && std::isfinite(s.buck_A)) {
                const double l = std::log(s.buck_A);
                if (std::fabs(l) < 500.0) { d.b_lnA = l; d.flags |= CEG_PAIRHOT_LNA; }
            }
            if (s.flags & CEG_PAIR_HS) h->max_hs_r2 = std::max(h->max_hs_r2, d.hs_r2);
        }
    CEG_CK(acquire_stream(&h->s_compute));
    CEG_CK(acquire_stream(&h->s_compute2));
    CEG_CK(acquire_stream(&h->s_copy));
    CEG_CK(h->d_xyz.reserve(sizeof(double) * 3 * std::max<int64_t>(1, n_atoms)));
    CEG_CK(h->d_type.reserve(sizeof(int) * std::max<int64_t>(1, n_atoms)));
    CEG_CK(h->d_q.reserve(sizeof(double) * std::max<int64_t>(1, n_atoms)));
    CEG_CK(h->d_table.reserve(sizeof(PairDev) * tab.size()));
    CEG_CK(h->d_stats.reserve(2 * sizeof(unsigned long long)));
    // one stream, one synchronisation (the sources are pageable: each copy is staged before the call returns)
    if (n_atoms > 0) {
        CEG_CK(cudaMemcpyAsync(h->d_xyz.p, h->h_xyz.data(), sizeof(double) * 3 * n_atoms, cudaMemcpyHostToDevice, h->s_compute));
        CEG_CK(cudaMemcpyAsync(h->d_type.p, h->h_type.data(), sizeof(int) * n_atoms, cudaMemcpyHostToDevice, h->s_compute));
        CEG_CK(cudaMemcpyAsync(h->d_q.p, h->h_q.data(), sizeof(double) * n_atoms, cudaMemcpyHostToDevice, h->s_compute));
    }
    CEG_CK(cudaMemcpyAsync(h->d_table.p, tab.data(), sizeof(PairDev) * tab.size(), cudaMemcpyHostToDevice, h->s_compute));
    CEG_CK(cudaMemsetAsync(h->d_stats.p, 0, 2 * sizeof(unsigned long long), h->s_compute));
    CEG_CK(cudaStreamSynchronize(h->s_compute));
    *out = h.release();
    return 0;
}

extern "C" int ceg_probe_free(ceg_probe_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    delete h;
    return 0;
}

extern "C" int ceg_probe_num_grids(ceg_probe_h h, int32_t* n) {
    if (!h || !n) return fail("null argument");
    *n = h->n_probes + h->want_coulomb;
    return 0;
}

extern "C" int ceg_compute_derivatives(ceg_probe_h h, int64_t n_pos, const double* xyz, double* out) {
    if (!h || (n_pos > 0 && (!xyz || !out)) || n_pos < 0) return fail("ceg_compute_derivatives: bad argument");
    if (n_pos == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    const int ng = h->n_probes + h->want_coulomb;
    CEG_CK(h->d_pos.reserve(sizeof(double) * 3 * n_pos));
    CEG_CK(h->d_der.reserve(sizeof(double) * 8 * n_pos * ng));
    CEG_CK(cudaMemcpyAsync(h->d_pos.p, xyz, sizeof(double) * 3 * n_pos, cudaMemcpyHostToDevice, h->s_compute));
    LiteralParams P;
    P.cell = h->cell; P.cutoff2 = h->cutoff2; P.alpha = h->alpha;
    P.n_atoms = h->n_atoms; P.n_pos = n_pos;
    P.xyz = h->d_xyz.as<double>(); P.type = h->d_type.as<int>(); P.q = h->d_q.as<double>();
    P.table = h->d_table.as<PairDev>();
    P.n_types = h->n_types; P.n_probes = h->n_probes; P.want_coulomb = h->want_coulomb;
    P.pos = h->d_pos.as<double>(); P.out = h->d_der.as<double>();
    const int blocks = (int)std::min<int64_t>((n_pos + 127) / 128, (int64_t)h->sm_count * 16);
    derivatives_literal_kernel<<<blocks, 128, 0, h->s_compute>>>(P);
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaMemcpyAsync(out, h->d_der.p, sizeof(double) * 8 * n_pos * ng, cudaMemcpyDeviceToHost, h->s_compute));
    CEG_CK(cudaStreamSynchronize(h->s_compute));
    return 0;
}

// ---- expanded image list ----------------------------------------------------------------------------
// All periodic images a + n1*A + n2*B + n3*C (A,B,C = supercell vectors) lying inside the box
// [lo - reach, hi + reach], sorted along x (ties: atom order), so that a tile only scans the images its x planes
// can reach (prepare_build tabulates that index range per plane).
static void build_ext_list(const ceg_probe_s* h, const double lo[3], const double hi[3], double reach,
                           std::vector<ExtAtom>& ext, std::vector<ExtMeta>& meta) {
    const double* m = h->cell.mat;
    const double* im = h->cell.invmat;
    double blo[3], bhi[3];
    for (int d = 0; d < 3; d++) { blo[d] = lo[d] - reach; bhi[d] = hi[d] + reach; }
    // fractional bounding range of the box
    double fmin[3] = {1e300, 1e300, 1e300}, fmax[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < 8; c++) {
        double p[3] = {(c & 1) ? bhi[0] : blo[0], (c & 2) ? bhi[1] : blo[1], (c & 4) ? bhi[2] : blo[2]};
        for (int r = 0; r < 3; r++) {
            double f = im[r] * p[0] + im[3 + r] * p[1] + im[6 + r] * p[2];
            fmin[r] = std::min(fmin[r], f);
            fmax[r] = std::max(fmax[r], f);
        }
    }
    ext.clear();
    meta.clear();
    ext.reserve((size_t)h->n_atoms * 8);
    meta.reserve((size_t)h->n_atoms * 8);
    for (int64_t a = 0; a < h->n_atoms; a++) {
        const double* p = &h->h_xyz[3 * a];
        double f[3];
        for (int r = 0; r < 3; r++) f[r] = im[r] * p[0] + im[3 + r] * p[1] + im[6 + r] * p[2];
        int n0[3], n1[3];
        for (int r = 0; r < 3; r++) {
            n0[r] = (int)std::floor(fmin[r] - f[r] - 1e-9);
            n1[r] = (int)std::ceil(fmax[r] - f[r] + 1e-9);
        }
        for (int i1 = n0[0]; i1 <= n1[0]; i1++)
            for (int i2 = n0[1]; i2 <= n1[1]; i2++)
                for (int i3 = n0[2]; i3 <= n1[2]; i3++) {
                    // (the same expression for every image, whatever loop it sits in: positions are bit-reproducible)
                    const double q0 = p[0] + ((double)i1 * m[0] + (double)i2 * m[3] + (double)i3 * m[6]);
                    if (q0 < blo[0] || q0 > bhi[0]) continue;
                    const double q1 = p[1] + ((double)i1 * m[1] + (double)i2 * m[4] + (double)i3 * m[7]);
                    if (q1 < blo[1] || q1 > bhi[1]) continue;
                    const double q2 = p[2] + ((double)i1 * m[2] + (double)i2 * m[5] + (double)i3 * m[8]);
                    if (q2 < blo[2] || q2 > bhi[2]) continue;
                    ext.push_back(ExtAtom{q0, q1, q2, h->h_q[a]});
                    meta.push_back(ExtMeta{h->h_type[a], (int)a});
                }
    }
    // sorted by x, ties in generation order (= a stable sort): the summation order inside a tile is the list order
    std::vector<std::pair<double, int>> order(ext.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = {ext[i].x, (int)i};
    std::sort(order.begin(), order.end());
    std::vector<ExtAtom> ext2(ext.size());
    std::vector<ExtMeta> meta2(meta.size());
    for (size_t i = 0; i < order.size(); i++) { ext2[i] = ext[order[i].second]; meta2[i] = meta[order[i].second]; }
    ext.swap(ext2);
    meta.swap(meta2);
}

// Shared memory per CTA.  Coulomb passes carry TX + TY + TZ Gaussian factors per image: the list capacity is chosen so
// that two CTAs still fit an SM (2 x ~104 KB of the 227 KB), which holds the whole list of a tile in one round for
// frameworks up to ~530 images within reach (CHA: ~420); VdW-only passes keep 1024 entries (40 KB).
constexpr size_t BUILD_SMEM_TARGET = 104 << 10;
static int build_list_cap(int nt, int n_axes, bool coul, int n_probes, int n_types) {
    if (!build_uses_gtable(n_probes, coul)) return LIST_CAP_MAX;
    const size_t per = sizeof(ExtAtom) + sizeof(ListMeta) + sizeof(int) + sizeof(double) * n_axes;
    const size_t tab = build_table_bytes(n_probes, n_types);
    int cap = tab + per < BUILD_SMEM_TARGET ? (int)((BUILD_SMEM_TARGET - tab - per) / per) : 0;
    cap = std::min(cap, LIST_CAP_MAX) & ~3;
    return std::max(cap, nt);          // a scan chunk (one image per thread) must always fit
}

template <int NT, int NP, bool COUL, bool POLY>
static cudaError_t launch_tiled(const BuildParams& P, int n_tiles, cudaStream_t s) {
    static std::mutex mu;
    static size_t granted[64] = {0};     // per device: dynamic shared memory already opted in for this instantiation
    const size_t smem = build_smem_bytes(P.list_cap, P.tile[0] + P.tile[1] + P.tile[2], build_uses_gtable(NP, COUL), NP, P.n_types);
    if (smem > (48u << 10)) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && granted[dev] < smem) {
            cudaError_t e = cudaFuncSetAttribute(grid_build_tiled_kernel<NT, NP, COUL, POLY>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)(200u << 10)));
            if (e != cudaSuccess) return e;
            granted[dev] = std::max(smem, (size_t)(200u << 10));
        }
    }
    grid_build_tiled_kernel<NT, NP, COUL, POLY><<<n_tiles, NT, smem, s>>>(P);
    count_launch();
    return cudaGetLastError();
}

// Tile shapes.  The main shape is the most compact one (best pruning radius); a slab whose plane count
// is not a multiple of its x extent finishes with flatter shapes of the same thread count (x extent 4, 2,
// 1), so that no CTA carries idle planes and x-slabs of any size stay balanced across GPUs.
struct TileShape { int tx, ty, tz; };
#ifndef CEG_BUILD_NT
#define CEG_BUILD_NT 256       // threads per CTA of the <= 2 probe variants
#endif
#ifndef CEG_BUILD_NT_MANY
#define CEG_BUILD_NT_MANY 256  // threads per CTA of the 3 and 4 probe variants
#endif
static int build_nt(int n_probes) { return n_probes <= 2 ? CEG_BUILD_NT : CEG_BUILD_NT_MANY; }
static int main_tx(int n_probes) { return build_nt(n_probes) / 64; }
// tx x 8 x 8 nodes, the remaining factor of the thread count going to z and y alternately:
// 512 threads: {8,8,8} {4,8,16} {2,16,16} {1,16,32};  256: {4,8,8} {2,8,16} {1,16,16};  128: {2,8,8} {1,8,16}
static TileShape tile_shape_for(int n_probes, int tx) {
    TileShape t = {tx, 8, 8};
    for (int f = build_nt(n_probes) / (tx * 64); f > 1; f >>= 1) {
        if (t.tz <= t.ty) t.tz *= 2; else t.ty *= 2;
    }
    return t;
}
static double tile_radius(const TileShape& t, const double* delta) {
    const double hx = 0.5 * (t.tx - 1) * delta[0], hy = 0.5 * (t.ty - 1) * delta[1], hz = 0.5 * (t.tz - 1) * delta[2];
    return std::sqrt(hx * hx + hy * hy + hz * hz);
}

template <int NP, bool COUL>
static cudaError_t dispatch_np(const BuildParams& P, int n_tiles, cudaStream_t s) {
    constexpr int NT = NP <= 2 ? CEG_BUILD_NT : CEG_BUILD_NT_MANY;
    if (COUL && P.poly_erfcx) return launch_tiled<NT, NP, COUL, true>(P, n_tiles, s);
    return launch_tiled<NT, NP, COUL, false>(P, n_tiles, s);
}

static cudaError_t dispatch_tiled(int np, bool coul, const BuildParams& P, int n_tiles, cudaStream_t s) {
    switch (np * 2 + (coul ? 1 : 0)) {
        case 1: return dispatch_np<0, true>(P, n_tiles, s);
        case 2: return dispatch_np<1, false>(P, n_tiles, s);
        case 3: return dispatch_np<1, true>(P, n_tiles, s);
        case 4: return dispatch_np<2, false>(P, n_tiles, s);
        case 5: return dispatch_np<2, true>(P, n_tiles, s);
        case 6: return dispatch_np<3, false>(P, n_tiles, s);
        case 7: return dispatch_np<3, true>(P, n_tiles, s);
        case 8: return dispatch_np<4, false>(P, n_tiles, s);
        case 9: return dispatch_np<4, true>(P, n_tiles, s);
        default: return cudaErrorInvalidValue;
    }
}

// Launch the planes [a, b) of the slab described by P: full rows of the main shape, then the tail.
static cudaError_t launch_planes(const ceg_probe_s* h, BuildParams P, const ceg_gridspec_t* gs, double cutoff, int a, int b,
                                 cudaStream_t s) {
    int tx = main_tx(h->n_probes);
    while (a < b) {
        while (tx > b - a) tx >>= 1;
        const int rows = (b - a) / tx;
        const TileShape t = tile_shape_for(h->n_probes, tx);
        const double reach = cutoff + tile_radius(t, gs->delta) + 1e-6;
        P.reach2 = reach * reach;
        P.tile[0] = t.tx; P.tile[1] = t.ty; P.tile[2] = t.tz;
        P.list_cap = build_list_cap(build_nt(h->n_probes), t.tx + t.ty + t.tz, h->want_coulomb != 0, h->n_probes, h->n_types);
        P.a0 = a; P.a1 = a + rows * tx;
        P.tiles_x = rows;
        P.tiles_y = (gs->dims[1] + 1 + t.ty - 1) / t.ty;
        P.tiles_z = (gs->dims[2] + 1 + t.tz - 1) / t.tz;
        cudaError_t e = dispatch_tiled(h->n_probes, h->want_coulomb != 0, P, P.tiles_x * P.tiles_y * P.tiles_z, s);
        if (e != cudaSuccess) return e;
        a = P.a1;
    }
    return cudaSuccess;
}

// Fill BuildParams and make sure the image list for slab [i0,i1) is resident.
static int prepare_build(ceg_probe_s* h, const ceg_gridspec_t* gs, int i0, int i1, double lam_v, double thr_v,
                         double lam_c, double thr_c, BuildParams& P) {
    if (!gs) return fail("null gridspec");
    for (int d = 0; d < 3; d++)
        if (gs->dims[d] < 1 || !(gs->size[d] > 0)) return fail("bad grid spec");
    if (i0 < 0 || i1 > gs->dims[0] + 1 || i0 >= i1) return fail("bad x-slab range");
    const double cutoff = std::sqrt(h->cutoff2);
    if (!(2.0 * cutoff <= h->min_perp_width * (1.0 + 1e-12)))
        return fail("cutoff exceeds half the perpendicular width of the probe supercell: the tiled kernel requires "
                    "the supercell of probes.jl:24 (use ceg_compute_derivatives for the literal algorithm)");
    P.cell = h->cell;
    P.gs = to_dev(*gs);
    P.cutoff2 = h->cutoff2;
    P.alpha = h->alpha;
    P.alpha2 = h->alpha * h->alpha;
    P.kk[0] = 2.0 * h->alpha / std::sqrt(M_PI);
    P.kk[1] = 2.0 * P.alpha2 * P.kk[0];
    P.kk[2] = 4.0 * P.alpha2 * P.alpha2 * P.kk[0];
    P.alpha_p = h->alpha * ERFCX_P;
    for (int d = 0; d < 3; d++) P.gauss_step[d] = std::exp(-2.0 * P.alpha2 * gs->delta[d] * gs->delta[d]);
    P.cut_hi = h->cutoff2 + GUARD_BAND;
    // fast path window: away from every mask threshold (hard spheres, the 1 Å rule, the cutoff, the
    // entry of the reference's non-orthogonal image search) and from r -> 0
    double flo = std::max(h->max_hs_r2, NEAR_R2);
    if (h->want_coulomb) flo = std::max(flo, 1.0);
    P.fast_lo = flo + GUARD_BAND;
    double fhi = h->cutoff2 - GUARD_BAND;
    if (!h->cell.ortho && h->cell.safemin2 < h->cutoff2) fhi = std::min(fhi, h->cell.safemin2 - GUARD_BAND);
    P.fast_hi = fhi;
    {   // integer versions (high words): fast iff hi_fast_lo <= hi(d2) < hi_fast_lo + span, i.e. strictly
        // above the high word of fast_lo and strictly below that of fast_hi
        auto hiword = [](double v) { long long b; memcpy(&b, &v, 8); return (int)(b >> 32); };
        P.hi_cut = hiword(P.cut_hi);
        P.hi_fast_lo = hiword(P.fast_lo) + 1;
        const int top = hiword(P.fast_hi);          // exclusive
        P.hi_fast_span = top > P.hi_fast_lo ? (unsigned)(top - P.hi_fast_lo) : 0u;
    }
    P.poly_erfcx = (h->alpha * cutoff <= ERFCX_XMAX && h->alpha * std::sqrt(P.fast_lo) >= ERFCX_XMIN) ? 1 : 0;
    P.list_cap = LIST_CAP_MAX;                       // per launch: launch_planes
    double rmax = 0.0;
    for (int tx = main_tx(h->n_probes); tx >= 1; tx >>= 1)
        rmax = std::max(rmax, tile_radius(tile_shape_for(h->n_probes, tx), gs->delta));
    const double reach = cutoff + rmax + 1e-6;      // image-list margin: the widest tile shape
    P.reach2 = reach * reach;                        // per launch: launch_planes
    P.lam_v = lam_v; P.thr_v = thr_v; P.lam_c = lam_c; P.thr_c = thr_c;
    P.i0 = i0; P.i1 = i1; P.a0 = i0; P.a1 = i1;
    P.tiles_x = P.tiles_y = P.tiles_z = 0;
    P.tile[0] = P.tile[1] = P.tile[2] = 0;
    P.atoms_xyz = h->d_xyz.as<double>();
    P.table = h->d_table.as<PairDev>();
    P.n_types = h->n_types;
    for (int g = 0; g < MAX_GRIDS; g++) P.out[g] = nullptr;
    P.raw = nullptr;
    P.raw_slab = 0;
    P.raw_plane0 = 0;
    P.n_dest = 0;
    P.rec_block = 0;
    for (int t = 0; t < MAX_BUILD_DEST; t++)
        for (int g = 0; g < MAX_GRIDS; g++) P.rec[t][g] = nullptr;
    for (int g = 0; g < MAX_GRIDS; g++) P.rec_scale[g] = 1.0;
    P.stats = h->d_stats.as<unsigned long long>();
    // image list for the slab's bounding box
    double lo[3], hi[3];
    lo[0] = i0 * gs->size[0] / gs->dims[0] + gs->shift[0];
    hi[0] = (i1 - 1) * gs->size[0] / gs->dims[0] + gs->shift[0];
    for (int d = 1; d < 3; d++) { lo[d] = gs->shift[d]; hi[d] = gs->shift[d] + gs->size[d]; }
    double key[8] = {lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], cutoff, reach};
    ceg_probe_s::ExtList* L = nullptr;
    for (auto& c : h->ext_cache)
        if (c.valid && memcmp(key, c.key, sizeof key) == 0 && c.key_planes[0] == i0 && c.key_planes[1] == i1) L = &c;
    if (!L) {
        const auto t_list0 = std::chrono::steady_clock::now();
        L = &h->ext_cache[0];
        for (auto& c : h->ext_cache)
            if (!c.valid || (L->valid && c.used < L->used)) L = &c;
        std::vector<ExtAtom> ext;
        std::vector<ExtMeta> meta;
        build_ext_list(h, lo, hi, reach + 1e-6, ext, meta);
        // per plane: the index range of the images within the cutoff along x (list sorted by x)
        const int planes = i1 - i0;
        std::vector<int> range(2 * (size_t)planes);
        const double xcut = std::sqrt(P.cut_hi) + 1e-9;
        auto xless = [](const ExtAtom& a, double v) { return a.x < v; };
        auto xgreater = [](double v, const ExtAtom& a) { return v < a.x; };
        for (int i = 0; i < planes; i++) {
            const double x = (double)(i0 + i) * gs->size[0] / gs->dims[0] + gs->shift[0];
            range[i] = (int)(std::lower_bound(ext.begin(), ext.end(), x - xcut, xless) - ext.begin());
            range[planes + i] = (int)(std::upper_bound(ext.begin(), ext.end(), x + xcut, xgreater) - ext.begin());
        }
        if (L->valid) CEG_CK(cudaDeviceSynchronize());   // an earlier asynchronous build may still be reading the list being replaced
        L->valid = false;
        L->n_ext = (int)ext.size();
        // one device buffer, one upload: [images | meta | per-plane ranges]
        const size_t b_ext = sizeof(ExtAtom) * std::max<size_t>(1, ext.size());
        const size_t b_meta = (sizeof(ExtMeta) * std::max<size_t>(1, meta.size()) + 15) & ~(size_t)15;
        const size_t b_range = sizeof(int) * range.size();
        L->off_meta = b_ext;
        L->off_range = b_ext + b_meta;
        std::vector<char> blob(b_ext + b_meta + b_range);
        if (!ext.empty()) {
            memcpy(blob.data(), ext.data(), sizeof(ExtAtom) * ext.size());
            memcpy(blob.data() + L->off_meta, meta.data(), sizeof(ExtMeta) * meta.size());
        }
        memcpy(blob.data() + L->off_range, range.data(), b_range);
        CEG_CK(L->d_ext.reserve(blob.size()));
        CEG_CK(cudaMemcpy(L->d_ext.p, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        memcpy(L->key, key, sizeof key);
        L->key_planes[0] = i0; L->key_planes[1] = i1;
        L->plane0 = i0; L->planes = planes;
        L->valid = true;
        h->list_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_list0).count();
    }
    L->used = ++h->ext_clock;
    P.ext = L->d_ext.as<ExtAtom>();
    P.ext_meta = reinterpret_cast<const ExtMeta*>(static_cast<const char*>(L->d_ext.p) + L->off_meta);
    P.n_ext = L->n_ext;
    P.ext_lo = reinterpret_cast<const int*>(static_cast<const char*>(L->d_ext.p) + L->off_range);
    P.ext_hi = P.ext_lo + L->planes;
    P.ext_plane0 = L->plane0;
    return 0;
}

// Every tile scans the image list of the range it belongs to, and that list grows with the range's bounding box:
// C1-C5 slabs give <= 8k images (the scan hides behind the co-resident CTA's pair phase), a 512-plane slab of an 85 Å
// cell 60k (15 % slower, tools/exp_stress.py).  Thick slabs of large cells are therefore processed as x windows of
// about 64 planes, each with its own list.
constexpr int EXT_SPLIT_MIN = 12000;
constexpr int LIST_WINDOW_PLANES = 64;
static int list_windows(ceg_probe_s* h, const ceg_gridspec_t* gs, int i0, int i1, double lam_v, double thr_v, double lam_c,
                        double thr_c, std::vector<std::pair<int, int>>& win) {
    win.clear();
    const int planes = i1 - i0;
    bool split = (h->split_i0 == i0 && h->split_i1 == i1 && gs && memcmp(h->split_dims, gs->dims, sizeof h->split_dims) == 0);
    if (!split && planes > 2 * LIST_WINDOW_PLANES) {
        BuildParams P;
        if (prepare_build(h, gs, i0, i1, lam_v, thr_v, lam_c, thr_c, P)) return 1;
        split = P.n_ext > EXT_SPLIT_MIN;
        if (split) {
            h->split_i0 = i0; h->split_i1 = i1;
            memcpy(h->split_dims, gs->dims, sizeof h->split_dims);
        }
    }
    if (!split) {
        win.push_back({i0, i1});
        return 0;
    }
    const int mtx = main_tx(h->n_probes);
    const int n_win = (planes + LIST_WINDOW_PLANES - 1) / LIST_WINDOW_PLANES;
    const int w = ((planes + n_win - 1) / n_win + mtx - 1) / mtx * mtx;
    for (int a = i0; a < i1; a += w) win.push_back({a, std::min(i1, a + w)});
    return 0;
}

// argument checks shared by ceg_grid_build / ceg_grid_build_dev: nothing is dereferenced or sized before they pass
static int check_build_args(const char* who, ceg_probe_h h, const ceg_gridspec_t* gs, int i0, int i1) {
    if (!h) return fail(std::string(who) + ": null handle");
    if (!gs) return fail(std::string(who) + ": null gridspec");
    for (int d = 0; d < 3; d++)
        if (gs->dims[d] < 1 || !(gs->size[d] > 0) || !std::isfinite(gs->size[d]) || !std::isfinite(gs->shift[d]) ||
            !(gs->delta[d] > 0))
            return fail(std::string(who) + ": bad grid spec");
    if (i0 < 0 || i1 > gs->dims[0] + 1 || i0 >= i1) return fail(std::string(who) + ": bad x-slab range");
    return 0;
}

// Record destinations of a build (BuildParams::rec): validated, peer access enabled where a destination lives on
// another device of this process (IPC mappings of other processes' memory arrive with access enabled, ceg_ipc_open).
struct RecordTargets {
    int n_dest = 0, block = 0;
    Rec8* rec[MAX_BUILD_DEST][MAX_GRIDS];
    double scale[MAX_GRIDS];
};
static int fill_record_targets(const char* who, ceg_probe_h h, int32_t n_dest, void* const* d_records, int32_t layout,
                               const double* scale, RecordTargets& T) {
    const int ng = h->n_probes + h->want_coulomb;
    if (n_dest < 1 || n_dest > MAX_BUILD_DEST) return fail(std::string(who) + ": between 1 and 8 destinations expected");
    if (!d_records || !scale) return fail(std::string(who) + ": null argument");
    if (layout != LAYOUT_NODE && layout != LAYOUT_BLOCK) return fail(std::string(who) + ": records are written in the NODE or BLOCK layout");
    T.n_dest = n_dest;
    T.block = layout == LAYOUT_BLOCK;
    for (int g = 0; g < ng; g++) {
        if (!(scale[g] > 0.0)) return fail(std::string(who) + ": scale must be positive");
        T.scale[g] = scale[g];
    }
    for (int t = 0; t < n_dest; t++)
        for (int g = 0; g < ng; g++) {
            void* p = d_records[(size_t)t * ng + g];
            if (!p || ((uintptr_t)p & 31)) return fail(std::string(who) + ": null or misaligned record buffer");
            T.rec[t][g] = reinterpret_cast<Rec8*>(p);
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeDevice) {
                cudaGetLastError();
                return fail(std::string(who) + ": record buffer is not device memory");
            }
            if (a.device != h->device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, h->device, a.device);
                if (!can) return fail(std::string(who) + ": no peer access from the probe's device to a destination device");
                cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
        }
    return 0;
}

static int grid_build_dev_impl(const char* who, ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1, double lambda_vdw,
                               double thr_vdw, double lambda_coul, double thr_coul, float* const* d_out,
                               const RecordTargets* T, void* stream) {
    std::vector<std::pair<int, int>> win;
    if (list_windows(h, gs, i0, i1, lambda_vdw, thr_vdw, lambda_coul, thr_coul, win)) return 1;
    const int ng = h->n_probes + h->want_coulomb;
    cudaStream_t s = (cudaStream_t)stream;
    CEG_CK(cudaMemsetAsync(h->d_stats.p, 0, 2 * sizeof(unsigned long long), s));
    for (const auto& w : win) {
        BuildParams P;
        if (prepare_build(h, gs, w.first, w.second, lambda_vdw, thr_vdw, lambda_coul, thr_coul, P)) return 1;
        P.i0 = i0; P.i1 = i1;                       // the output buffers hold the caller's whole slab
        for (int g = 0; g < ng; g++) P.out[g] = d_out ? d_out[g] : nullptr;
        if (T) {
            P.n_dest = T->n_dest;
            P.rec_block = T->block;
            for (int t = 0; t < T->n_dest; t++)
                for (int g = 0; g < ng; g++) P.rec[t][g] = T->rec[t][g];
            for (int g = 0; g < ng; g++) P.rec_scale[g] = T->scale[g];
        }
        cudaError_t e = launch_planes(h, P, gs, std::sqrt(h->cutoff2), w.first, w.second, s);
        if (e != cudaSuccess) return fail_cuda(e, who, __FILE__, __LINE__);
    }
    h->last_points = (int64_t)(i1 - i0) * (gs->dims[1] + 1) * (gs->dims[2] + 1);
    return 0;
}

extern "C" int ceg_grid_build_dev(ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1, double lambda_vdw,
                                  double thr_vdw, double lambda_coul, double thr_coul, float* const* d_out, void* stream) {
    if (check_build_args("ceg_grid_build_dev", h, gs, i0, i1)) return 1;
    if (!d_out) return fail("ceg_grid_build_dev: null argument");
    for (int g = 0; g < h->n_probes + h->want_coulomb; g++)
        if (!d_out[g]) return fail("ceg_grid_build_dev: null output buffer");
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    return grid_build_dev_impl("ceg_grid_build_dev", h, gs, i0, i1, lambda_vdw, thr_vdw, lambda_coul, thr_coul, d_out, nullptr, stream);
}

// The build with its hand-off fused in: every node is written as a finished interpolation record (parse_grid's
// conversion applied, NODE or BLOCK layout) into the record storage of n_dest grids — ceg_grid_create_empty handles on
// this device and on its NVLink peers — so a sharded build leaves a complete replica on every GPU when its kernels
// have finished; no all-gather, no re-layout pass.  d_out (optional) receives the `.grid`-order slab as well.
extern "C" int ceg_grid_build_records(ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1, double lambda_vdw,
                                      double thr_vdw, double lambda_coul, double thr_coul, int32_t n_dest,
                                      void* const* d_records, int32_t layout, const double* scale, float* const* d_out,
                                      void* stream) {
    if (check_build_args("ceg_grid_build_records", h, gs, i0, i1)) return 1;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    RecordTargets T;
    if (fill_record_targets("ceg_grid_build_records", h, n_dest, d_records, layout, scale, T)) return 1;
    if (d_out)
        for (int g = 0; g < h->n_probes + h->want_coulomb; g++)
            if (!d_out[g]) return fail("ceg_grid_build_records: null output buffer");
    return grid_build_dev_impl("ceg_grid_build_records", h, gs, i0, i1, lambda_vdw, thr_vdw, lambda_coul, thr_coul, d_out, &T, stream);
}

// Device buffers across processes (CUDA IPC): the owner exports, a peer process opens the mapping (peer access
// enabled) and passes it as a record destination.  Library buffers are whole cudaMalloc allocations or sit at a fixed
// offset inside one; the offset travels with the handle.
extern "C" int ceg_ipc_export(int32_t device, const void* d_ptr, uint8_t* handle64, int64_t* offset) {
    if (!d_ptr || !handle64 || !offset) return fail("ceg_ipc_export: null argument");
    if (use_device(device)) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ceg_ipc_export: 64-byte handles");
    // base of the allocation d_ptr lies in (driver entry point fetched through the runtime: the library does not link libcuda)
    typedef int (*range_fn_t)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CEG_CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail("ceg_ipc_export: cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn_t>(fn)(&base, &size, (unsigned long long)(uintptr_t)d_ptr) != 0)
        return fail("ceg_ipc_export: not a device allocation");
    cudaIpcMemHandle_t hd;
    CEG_CK(cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base)));
    memcpy(handle64, &hd, 64);
    *offset = (int64_t)((unsigned long long)(uintptr_t)d_ptr - base);
    return 0;
}
namespace {
std::mutex g_ipc_mu;
std::vector<std::pair<void*, void*>> g_ipc_open;     // (pointer handed out, mapping base)
}  // namespace
extern "C" int ceg_ipc_open(int32_t device, const uint8_t* handle64, int64_t offset, void** d_ptr) {
    if (!handle64 || !d_ptr || offset < 0) return fail("ceg_ipc_open: bad argument");
    if (use_device(device)) return 1;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    void* base = nullptr;
    CEG_CK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = (char*)base + offset;
    std::lock_guard<std::mutex> lk(g_ipc_mu);
    g_ipc_open.push_back({*d_ptr, base});
    return 0;
}
extern "C" int ceg_ipc_close(int32_t device, void* d_ptr) {
    if (!d_ptr) return 0;
    void* base = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ipc_mu);
        for (size_t i = 0; i < g_ipc_open.size(); i++)
            if (g_ipc_open[i].first == d_ptr) { base = g_ipc_open[i].second; g_ipc_open.erase(g_ipc_open.begin() + i); break; }
    }
    if (!base) return fail("ceg_ipc_close: not a ceg_ipc_open pointer");
    if (use_device(device)) return 1;
    CEG_CK(cudaDeviceSynchronize());
    CEG_CK(cudaIpcCloseMemHandle(base));
    return 0;
}

extern "C" int ceg_grid_build(ceg_probe_h h, const ceg_gridspec_t* gs, int32_t i0, int32_t i1, double lambda_vdw,
                              double thr_vdw, double lambda_coul, double thr_coul, float* const* out_f32, double* raw_f64) {
    if (check_build_args("ceg_grid_build", h, gs, i0, i1)) return 1;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    const auto t_call0 = std::chrono::steady_clock::now();
    std::vector<std::pair<int, int>> win;
    h->list_ms = 0.0;
    if (list_windows(h, gs, i0, i1, lambda_vdw, thr_vdw, lambda_coul, thr_coul, win)) return 1;
    const int ng = h->n_probes + h->want_coulomb;
    const int64_t ny = gs->dims[1] + 1, nz = gs->dims[2] + 1, nx = gs->dims[0] + 1;
    const int64_t plane = ny * nz, slab = (int64_t)(i1 - i0) * plane, npts = nx * plane;
    bool any_out = false;
    for (int g = 0; g < ng; g++) any_out |= (out_f32 && out_f32[g]);
    // Device scratch is a ring of RING x-chunks (not the whole slab): chunk c is computed on s_compute into
    // slot c % RING while earlier chunks drain to the host on s_copy, so the footprint stays bounded (3 x 192 MB at
    // most) for any grid size and no large allocation is made or freed per call.  Large slabs start with big chunks
    // (few launches, small launch tails) and finish with small ones (only the last chunk's copy is not overlapped).
    constexpr int RING = 3;
    const int mtx = main_tx(h->n_probes);
    const int small = std::max(mtx, ((int)((48ll << 20) / (plane * 8 * 4 * ng)) / mtx) * mtx);
    const int big = (i1 - i0 >= 16 * small) ? 4 * small : small;
    // A call that builds the WHOLE grid has the host link to itself: its copies take about half a chunk's kernel time, and
    // chunks alternate between two compute streams so that the last, partly filled wave of one launch shares the SMs
    // with the first CTAs of the next (13 launches x half a wave was 1 ms of 62 on C4).  Co-resident chunks can delay
    // the older chunk's completion and with it its copy, though: an x-slab of a sharded build — where the ranks compete
    // for the host's ingest bandwidth and the copies, not the kernels, set the pace — keeps one stream and starts with
    // a single tile row instead, so that the copy engine gets work as early as possible
    // (on a box whose two-GPU ingest is 51 GB/s the call takes 35.5 ms = the copy time under every chunk policy tried).
    const bool whole_grid = (i0 == 0 && i1 == (int)nx);
    bool first_chunk = !whole_grid;
    auto next_chunk = [&](int remaining) {
        if (first_chunk) { first_chunk = false; if (remaining > 2 * mtx) return mtx; }
        if (remaining >= big + small) return big;
        if (remaining > 2 * small) return std::max(small, (remaining / 2) / mtx * mtx);
        if (remaining > mtx) return std::min(small, std::max(mtx, (remaining - mtx) / mtx * mtx));   // the last chunk is one tile row:
        return remaining;                                                            // its copy is the only one not overlapped
    };
    const int64_t chunk_slab = (int64_t)big * plane;
    if (any_out) CEG_CK(h->d_out.reserve(sizeof(float) * 8 * chunk_slab * ng * RING));
    if (raw_f64) CEG_CK(h->d_raw.reserve(sizeof(double) * 8 * slab * ng));
    CEG_CK(cudaMemsetAsync(h->d_stats.p, 0, 2 * sizeof(unsigned long long), h->s_compute));
    static_assert(RING == 3, "ceg_probe_s caches three ring events");
    for (int r = 0; r < RING; r++) {      // created once per handle
        if (!h->ev_done[r]) CEG_CK(cudaEventCreateWithFlags(&h->ev_done[r], cudaEventDisableTiming));
        if (!h->ev_free[r]) CEG_CK(cudaEventCreateWithFlags(&h->ev_free[r], cudaEventDisableTiming));
        if (!h->ev_time[r]) CEG_CK(cudaEventCreate(&h->ev_time[r]));
    }
    cudaEvent_t* ev_done = h->ev_done;
    cudaEvent_t* ev_free = h->ev_free;
    if (!h->ev_join) CEG_CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CEG_CK(cudaEventRecord(h->ev_time[0], h->s_compute));
    CEG_CK(cudaStreamWaitEvent(h->s_compute2, h->ev_time[0], 0));      // (after the statistics reset)
    int rc = 0, c = 0;
    for (size_t wi = 0; wi < win.size() && !rc; wi++) {
    BuildParams P;   // image list of this window (prepare_build drains the device before replacing the list)
    if (prepare_build(h, gs, win[wi].first, win[wi].second, lambda_vdw, thr_vdw, lambda_coul, thr_coul, P)) { rc = 1; break; }
    const int w1 = win[wi].second;
    for (int a = win[wi].first, b = a; a < w1 && !rc; a = b, c++) {
        b = a + next_chunk(w1 - a);
        const int slot = c % RING;
        float* base = any_out ? h->d_out.as<float>() + (int64_t)slot * 8 * chunk_slab * ng : nullptr;
        const int64_t cslab = (int64_t)(b - a) * plane;       // the kernel indexes outputs relative to [P.i0, P.i1)
        BuildParams Q = P;
        Q.i0 = a; Q.i1 = b;
        for (int g = 0; g < ng; g++) Q.out[g] = (out_f32 && out_f32[g]) ? base + (int64_t)g * 8 * cslab : nullptr;
        Q.raw = nullptr;
        if (raw_f64) {   // parity-test output: whole slab on the device, written through an offset view
            Q.raw = h->d_raw.as<double>();
            Q.raw_slab = slab;
            Q.raw_plane0 = a - i0;
        }
        cudaStream_t cs = ((c & 1) && whole_grid) ? h->s_compute2 : h->s_compute;
        if (c >= RING) cudaStreamWaitEvent(cs, ev_free[slot], 0);   // slot drained?
        cudaError_t e = launch_planes(h, Q, gs, std::sqrt(h->cutoff2), a, b, cs);
        if (e != cudaSuccess) { rc = fail_cuda(e, "grid_build_tiled_kernel", __FILE__, __LINE__); break; }
        if (any_out) {
            cudaEventRecord(ev_done[slot], cs);
            cudaStreamWaitEvent(h->s_copy, ev_done[slot], 0);
            for (int g = 0; g < ng && !rc; g++) {
                if (!Q.out[g]) continue;
                // one copy per channel run (a pitched 2-D copy is refused when only the runs, not the rows between them,
                // are registered — what a multi-process driver does with its slab, sharding.slab_file_runs)
                for (int ch = 0; ch < 8; ch++) {
                    e = cudaMemcpyAsync(out_f32[g] + (int64_t)ch * npts + (int64_t)a * plane, Q.out[g] + (int64_t)ch * cslab,
                                        sizeof(float) * cslab, cudaMemcpyDeviceToHost, h->s_copy);
                    if (e != cudaSuccess) { rc = fail_cuda(e, "cudaMemcpyAsync(D2H slab)", __FILE__, __LINE__); break; }
                }
            }
            cudaEventRecord(ev_free[slot], h->s_copy);
        }
    }
    }
    cudaEventRecord(h->ev_join, h->s_compute2);
    cudaStreamWaitEvent(h->s_compute, h->ev_join, 0);
    cudaEventRecord(h->ev_time[1], h->s_compute);
    cudaStreamWaitEvent(h->s_copy, h->ev_time[1], 0);
    cudaEventRecord(h->ev_time[2], h->s_copy);
    cudaError_t e1 = cudaStreamSynchronize(h->s_compute);
    cudaError_t e2 = cudaStreamSynchronize(h->s_copy);
    if (rc) return rc;
    if (e1 != cudaSuccess) return fail_cuda(e1, "grid build (compute stream)", __FILE__, __LINE__);
    if (e2 != cudaSuccess) return fail_cuda(e2, "grid build (copy stream)", __FILE__, __LINE__);
    if (raw_f64) CEG_CK(cudaMemcpy(raw_f64, h->d_raw.p, sizeof(double) * 8 * slab * ng, cudaMemcpyDeviceToHost));
    h->last_points = slab;
    float k_ms = 0.f, t_ms = 0.f;
    cudaEventElapsedTime(&k_ms, h->ev_time[0], h->ev_time[1]);
    cudaEventElapsedTime(&t_ms, h->ev_time[1], h->ev_time[2]);
    h->last_timing[0] = h->list_ms;
    h->last_timing[1] = k_ms;
    h->last_timing[2] = t_ms;
    h->last_timing[3] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call0).count();
    return 0;
}

extern "C" int ceg_probe_last_stats(ceg_probe_h h, int64_t* n_points, int64_t* pairs_tested, int64_t* pairs_in_cutoff) {
    if (!h) return fail("null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    unsigned long long st[2];
    CEG_CK(cudaMemcpy(st, h->d_stats.p, sizeof st, cudaMemcpyDeviceToHost));
    if (n_points) *n_points = h->last_points;
    if (pairs_tested) *pairs_tested = (int64_t)st[0];
    if (pairs_in_cutoff) *pairs_in_cutoff = (int64_t)st[1];
    return 0;
}

// Registers per thread of the library's main kernels as loaded on `device` (bench.py matches them against the ncu
// captures under profiles/ before quoting a DRAM-traffic figure): 0 = grid_build_tiled_kernel<NT,1,true,true>,
// 1 = interpolate_kernel<BLOCK>, 2 = interpolate_pair_kernel, 3 = reciprocal_rows_kernel<1>.
extern "C" int ceg_kernel_registers(int32_t device, int32_t which, int32_t* regs) {
    if (!regs) return fail("ceg_kernel_registers: null argument");
    if (use_device(device)) return 1;
    cudaFuncAttributes a;
    cudaError_t e;
    switch (which) {
        case 0: e = cudaFuncGetAttributes(&a, grid_build_tiled_kernel<CEG_BUILD_NT, 1, true, true>); break;
        case 1: e = cudaFuncGetAttributes(&a, interpolate_kernel<LAYOUT_BLOCK>); break;
        case 2: e = cudaFuncGetAttributes(&a, interpolate_pair_kernel); break;
        case 3: e = cudaFuncGetAttributes(&a, reciprocal_rows_kernel<1>); break;
        default: return fail("ceg_kernel_registers: unknown kernel");
    }
    CEG_CK(e);
    *regs = a.numRegs;
    return 0;
}

// Plain device buffers for callers that chain _dev entry points without a CUDA runtime of their own (the host layer's
// device-resident build -> gather -> upload hand-off); served from the library's allocation cache.
extern "C" int ceg_device_malloc(int32_t device, int64_t bytes, void** out) {
    if (!out || bytes <= 0) return fail("ceg_device_malloc: bad argument");
    if (use_device(device)) return 1;
    size_t got = 0;
    void* p = pool_acquire((size_t)bytes, &got);
    if (!p) {
        cudaError_t e = cudaMalloc(&p, (size_t)bytes);
        if (e != cudaSuccess) { cudaGetLastError(); pool_trim(); e = cudaMalloc(&p, (size_t)bytes); }
        CEG_CK(e);
        got = (size_t)bytes;
    }
    {
        std::lock_guard<std::mutex> lk(g_user_mu);
        g_user_bufs.push_back({p, got, device});
    }
    *out = p;
    return 0;
}
extern "C" int ceg_device_free(void* ptr) {
    if (!ptr) return 0;
    PoolEntry ent{nullptr, 0, 0};
    {
        std::lock_guard<std::mutex> lk(g_user_mu);
        for (size_t i = 0; i < g_user_bufs.size(); i++)
            if (g_user_bufs[i].p == ptr) { ent = g_user_bufs[i]; g_user_bufs.erase(g_user_bufs.begin() + i); break; }
    }
    if (!ent.p) return fail("ceg_device_free: not a ceg_device_malloc pointer");
    if (use_device(ent.device)) return 1;
    CEG_CK(cudaDeviceSynchronize());
    pool_release(ent.p, ent.bytes);
    return 0;
}
// device -> host copy of a ceg_device_malloc buffer (e.g. to write the `.grid` file of a grid that stays on the device)
extern "C" int ceg_device_download(int32_t device, const void* d_src, int64_t bytes, void* host_dst) {
    if (!d_src || !host_dst || bytes < 0) return fail("ceg_device_download: bad argument");
    if (use_device(device)) return 1;
    CEG_CK(cudaMemcpy(host_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ceg_device_synchronize(int32_t device) {
    if (use_device(device)) return 1;
    CEG_CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int ceg_probe_last_timing(ceg_probe_h h, double* out4) {
    if (!h || !out4) return fail("ceg_probe_last_timing: null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    for (int i = 0; i < 4; i++) out4[i] = h->last_timing[i];
    return 0;
}

// Gather x-slabs that live on several devices of this process into one full grid on `dst_device` (peer copies over
// NVLink; the reference exchanges them through `.grid` files, raspa.jl:430-438).  Part p holds planes [i0[p], i1[p])
// as [8][i1-i0][ny][nz] floats on devices[p]; d_full is [8][nx][ny][nz] on dst_device.
extern "C" int ceg_grid_gather(int32_t n_parts, const int32_t* devices, const float* const* d_slabs, const int32_t* i0,
                               const int32_t* i1, const ceg_gridspec_t* gs, int32_t dst_device, float* d_full) {
    if (n_parts <= 0 || !devices || !d_slabs || !i0 || !i1 || !gs || !d_full) return fail("ceg_grid_gather: null argument");
    const int64_t nx = gs->dims[0] + 1, plane = (int64_t)(gs->dims[1] + 1) * (gs->dims[2] + 1), npts = nx * plane;
    for (int p = 0; p < n_parts; p++)
        if (!d_slabs[p] || i0[p] < 0 || i1[p] > nx || i0[p] >= i1[p]) return fail("ceg_grid_gather: bad part");
    if (use_device(dst_device)) return 1;
    cudaStream_t s;
    CEG_CK(acquire_stream(&s));
    int rc = 0;
    for (int p = 0; p < n_parts && !rc; p++) {
        const int64_t cslab = (int64_t)(i1[p] - i0[p]) * plane;
        if (devices[p] != dst_device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, dst_device, devices[p]);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[p], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = fail_cuda(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__); break; }
                cudaGetLastError();
            }
        }
        for (int c = 0; c < 8 && !rc; c++) {   // (without peer access the runtime stages the copy through the host)
            cudaError_t e = cudaMemcpyPeerAsync(d_full + (int64_t)c * npts + (int64_t)i0[p] * plane, dst_device,
                                                d_slabs[p] + (int64_t)c * cslab, devices[p], sizeof(float) * cslab, s);
            if (e != cudaSuccess) rc = fail_cuda(e, "cudaMemcpyPeerAsync", __FILE__, __LINE__);
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    release_stream(s);
    if (!rc && e != cudaSuccess) rc = fail_cuda(e, "ceg_grid_gather", __FILE__, __LINE__);
    return rc;
}

// ================================================================================================
// interpolation
// ================================================================================================
static std::atomic<int> g_interp_layout{-1};   // -1 = automatic

extern "C" int ceg_set_interp_layout(int32_t layout) {
    if (layout < -1 || layout > LAYOUT_BLOCK) return fail("ceg_set_interp_layout: layout must be -1 (default), 0, 1, 2, 3 or 4");
    g_interp_layout.store(layout);
    return 0;
}

extern "C" int ceg_grid_layout(ceg_grid_h h, int32_t* layout) {
    if (!h || !layout) return fail("null argument");
    *layout = h->g.layout;
    return 0;
}

static int default_layout(size_t node_bytes) {
    return (node_bytes > ((size_t)96 << 20) && node_bytes < ((size_t)1536 << 20)) ? LAYOUT_BLOCK : LAYOUT_NODE;
}
static size_t record_bytes(int layout, int nx, int ny, int nz) {
    const size_t bricks = (size_t)((nx + 1) / 2) * ((ny + 1) / 2) * ((nz + 1) / 2);
    return layout == LAYOUT_BLOCK ? sizeof(Rec8) * 8 * bricks : sizeof(Rec8) * (size_t)nx * ny * nz;
}

// An interpolation grid whose records are still to be written — by ceg_grid_build_records on this device and/or on the
// other GPUs that hold the rest of the x range.  Layout: the one ceg_grid_upload would choose (NODE or BLOCK).
extern "C" int ceg_grid_create_empty(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, int32_t is_vdw,
                                     ceg_grid_h* out) {
    if (!unitcell || !gs || !out) return fail("ceg_grid_create_empty: null argument");
    for (int d = 0; d < 3; d++)
        if (gs->dims[d] < 1) return fail("ceg_grid_create_empty: bad dims");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_grid_s> h(new ceg_grid_s);
    h->device = device;
    h->g.uc = to_dev(*unitcell);
    h->g.gs = to_dev(*gs);
    h->g.is_vdw = is_vdw ? 1 : 0;
    const int nx = gs->dims[0] + 1, ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    int layout = g_interp_layout.load();
    if (layout != LAYOUT_NODE && layout != LAYOUT_BLOCK) layout = default_layout(sizeof(Rec8) * (size_t)nx * ny * nz);
    CEG_CK(acquire_stream(&h->stream));
    CEG_CK(h->d_nodes.reserve(record_bytes(layout, nx, ny, nz)));
    h->g.data = h->d_nodes.as<float>();
    h->g.layout = layout;
    *out = h.release();
    return 0;
}
extern "C" int ceg_grid_records(ceg_grid_h h, void** d_records, int64_t* bytes, int32_t* layout) {
    if (!h) return fail("ceg_grid_records: null handle");
    if (d_records) *d_records = h->d_nodes.p;
    if (bytes) *bytes = (int64_t)record_bytes(h->g.layout, h->g.gs.n[0], h->g.gs.n[1], h->g.gs.n[2]);
    if (layout) *layout = h->g.layout;
    return 0;
}

// grid: host pointer (from_device = false) or device pointer on `device` (true), [8][nx][ny][nz]; or `parts` (device,
// x-slabs of a sharded build, see GridSource)
static int grid_upload_impl(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, const float* grid,
                            bool from_device, double scale, int32_t is_vdw, ceg_grid_h* out, const GridSource* parts = nullptr) {
    if (!unitcell || !gs || (!grid && !parts) || !out) return fail("ceg_grid_upload: null argument");
    for (int d = 0; d < 3; d++)
        if (gs->dims[d] < 1) return fail("ceg_grid_upload: bad dims");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_grid_s> h(new ceg_grid_s);
    h->device = device;
    h->g.uc = to_dev(*unitcell);
    h->g.gs = to_dev(*gs);
    h->g.is_vdw = is_vdw ? 1 : 0;
    const int nx = gs->dims[0] + 1, ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    const int64_t npts = (int64_t)nx * ny * nz;
    const size_t node_bytes = sizeof(float) * 8 * (size_t)npts;
    CEG_CK(acquire_stream(&h->stream));
    // layout unless the caller asked otherwise: 2x2x2-node bricks once the grid no longer fits the L2, plain node
    // records below.  Measured on B200 (10^7 random queries): 34 MB grid NODE 0.384 / BLOCK 0.396 ms; 270 MB grid
    // NODE 0.700 / BLOCK 0.627 / PAIR 1.12 / QUAD 1.42 / CELL 1.57 ms; 907 MB grid NODE 1.359 / BLOCK 1.356 ms; 4 GB
    // grid (10^9-query sweep) NODE 174 / BLOCK 184 ms.  The
    // duplicating layouts lose to their own footprint; the bricks remove half of the 64-byte-granularity over-fetch
    // of a z pair without growing the grid.
    int layout = g_interp_layout.load();
    if (layout < 0) layout = default_layout(node_bytes);
    {
        DevBuf tmp, nodes;
        const bool direct = (layout == LAYOUT_NODE || layout == LAYOUT_BLOCK);    // one pass straight into the final layout
        DevBuf& node_dst = direct ? h->d_nodes : nodes;
        CEG_CK(node_dst.reserve(layout == LAYOUT_BLOCK ? record_bytes(layout, nx, ny, nz) : node_bytes));
        GridSource S;
        memset(&S, 0, sizeof S);
        if (parts) {
            S = *parts;
            CEG_CK(cudaDeviceSynchronize());     // the caller's producers (any stream) have finished writing the parts
        } else if (!from_device) {
            CEG_CK(tmp.reserve(node_bytes));
            CEG_CK(cudaMemcpyAsync(tmp.p, grid, node_bytes, cudaMemcpyHostToDevice, h->stream));
            S.whole = tmp.as<float>();
        } else {
            S.whole = grid;
            CEG_CK(cudaDeviceSynchronize());     // the caller's producer (any stream) has finished writing the grid
        }
        const int blocks = (int)std::min<int64_t>((npts + 255) / 256, sm_count_here() * 32);
        if (layout == LAYOUT_BLOCK)
            relayout_nodes_kernel<true><<<blocks, 256, 0, h->stream>>>(S, node_dst.as<Rec8>(), nx, ny, nz, scale);
        else
            relayout_nodes_kernel<false><<<blocks, 256, 0, h->stream>>>(S, node_dst.as<Rec8>(), nx, ny, nz, scale);
        count_launch();
        CEG_CK(cudaGetLastError());
        if (!direct) {
            const int corners = (layout == LAYOUT_PAIR) ? 2 : (layout == LAYOUT_QUAD) ? 4 : 8;
            CEG_CK(h->d_nodes.reserve(node_bytes * corners));
            expand_records_kernel<<<sm_count_here() * 32, 256, 0, h->stream>>>(nodes.as<Rec8>(), h->d_nodes.as<Rec8>(), nx, ny, nz, corners);
            count_launch();
            CEG_CK(cudaGetLastError());
        }
        CEG_CK(cudaStreamSynchronize(h->stream));
    }
    h->g.data = h->d_nodes.as<float>();
    h->g.layout = layout;
    *out = h.release();
    return 0;
}

// x-slabs of a sharded build, as they sit on this device after an all-gather, straight into an interpolation handle:
// part r holds planes [i0[r], i1[r]) in a [8][stride[r]][ny][nz] buffer (stride >= i1 - i0: padded gathers).
extern "C" int ceg_grid_upload_parts_dev(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, int32_t n_parts,
                                         const float* const* d_parts, const int32_t* i0, const int32_t* i1, const int64_t* stride,
                                         double scale, int32_t is_vdw, ceg_grid_h* out) {
    if (n_parts <= 0 || n_parts > MAX_GRID_PARTS || !d_parts || !i0 || !i1 || !gs)
        return fail("ceg_grid_upload_parts_dev: between 1 and 64 parts expected");
    if (!(scale > 0.0)) return fail("ceg_grid_upload_parts_dev: scale must be positive");
    GridSource S;
    memset(&S, 0, sizeof S);
    S.n_parts = n_parts;
    int next = 0;
    for (int r = 0; r < n_parts; r++) {
        if (!d_parts[r] || i0[r] != next || i1[r] <= i0[r]) return fail("ceg_grid_upload_parts_dev: parts must tile the x range in order");
        S.part[r] = d_parts[r]; S.i0[r] = i0[r]; S.i1[r] = i1[r];
        S.stride[r] = stride ? stride[r] : (int64_t)(i1[r] - i0[r]);
        if (S.stride[r] < i1[r] - i0[r]) return fail("ceg_grid_upload_parts_dev: stride shorter than the part");
        next = i1[r];
    }
    if (next != gs->dims[0] + 1) return fail("ceg_grid_upload_parts_dev: parts do not cover the grid");
    return grid_upload_impl(device, unitcell, gs, nullptr, true, scale, is_vdw, out, &S);
}

extern "C" int ceg_grid_upload(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, const float* grid,
                               int32_t is_vdw, ceg_grid_h* out) {
    return grid_upload_impl(device, unitcell, gs, grid, false, 1.0, is_vdw, out);
}

extern "C" int ceg_grid_upload_dev(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs, const float* d_grid,
                                   double scale, int32_t is_vdw, ceg_grid_h* out) {
    if (!(scale > 0.0)) return fail("ceg_grid_upload_dev: scale must be positive");
    return grid_upload_impl(device, unitcell, gs, d_grid, true, scale, is_vdw, out);
}

extern "C" int ceg_grid_free(ceg_grid_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    delete h;
    return 0;
}

static int interp_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, sm_count_here() * 64)); }

static cudaError_t launch_interpolate(const GridDev& g, int64_t n, const double* d_xyz, double* d_out, cudaStream_t s) {
    const int blocks = interp_blocks(n);
    switch (g.layout) {
        case LAYOUT_CELL: interpolate_kernel<LAYOUT_CELL><<<blocks, 256, 0, s>>>(g, n, d_xyz, d_out); break;
        case LAYOUT_QUAD: interpolate_kernel<LAYOUT_QUAD><<<blocks, 256, 0, s>>>(g, n, d_xyz, d_out); break;
        case LAYOUT_PAIR: interpolate_kernel<LAYOUT_PAIR><<<blocks, 256, 0, s>>>(g, n, d_xyz, d_out); break;
        case LAYOUT_BLOCK: interpolate_kernel<LAYOUT_BLOCK><<<blocks, 256, 0, s>>>(g, n, d_xyz, d_out); break;
        default: interpolate_kernel<LAYOUT_NODE><<<blocks, 256, 0, s>>>(g, n, d_xyz, d_out); break;
    }
    count_launch();
    return cudaGetLastError();
}

extern "C" int ceg_interpolate_dev(ceg_grid_h h, int64_t n, const double* d_xyz, double* d_out, void* stream) {
    if (!h || n < 0 || (n > 0 && (!d_xyz || !d_out))) return fail("ceg_interpolate_dev: bad argument");
    if (n == 0) return 0;
    if (use_device(h->device)) return 1;
    CEG_CK(launch_interpolate(h->g, n, d_xyz, d_out, (cudaStream_t)stream));
    return 0;
}

extern "C" int ceg_interpolate_sweep(ceg_grid_h h, int64_t first, int64_t n, uint64_t seed, const double* mat,
                                     double* sum, double* min, int64_t* n_blocked, int64_t n_keep, double* keep,
                                     float* kernel_ms) {
    if (!h || first < 0 || n < 0 || !mat || !sum || !min || !n_blocked || n_keep < 0 || (n_keep > 0 && !keep))
        return fail("ceg_interpolate_sweep: bad argument");
    *sum = 0.0; *min = INFINITY; *n_blocked = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    if (n == 0) return 0;
    n_keep = std::min(n_keep, n);
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    const int blocks = SWEEP_BLOCKS;    // fixed: the reduction order (and so the sum) depends neither on n nor on the device
    DevBuf d_mat, d_part, d_res, d_keep;
    CEG_CK(d_mat.reserve(sizeof(double) * 9));
    CEG_CK(d_part.reserve(sizeof(SweepPartial) * blocks));
    CEG_CK(d_res.reserve(sizeof(SweepPartial)));
    CEG_CK(d_keep.reserve(sizeof(double) * std::max<int64_t>(1, n_keep)));
    cudaStream_t s = h->stream;
    CEG_CK(cudaMemcpyAsync(d_mat.p, mat, sizeof(double) * 9, cudaMemcpyHostToDevice, s));
    cudaEvent_t e0, e1;
    CEG_CK(cudaEventCreate(&e0));
    CEG_CK(cudaEventCreate(&e1));
    cudaEventRecord(e0, s);
    switch (h->g.layout) {
        case LAYOUT_CELL: interpolate_sweep_kernel<LAYOUT_CELL><<<blocks, SWEEP_THREADS, 0, s>>>(h->g, first, n, seed, d_mat.as<double>(), n_keep, d_keep.as<double>(), d_part.as<SweepPartial>()); break;
        case LAYOUT_QUAD: interpolate_sweep_kernel<LAYOUT_QUAD><<<blocks, SWEEP_THREADS, 0, s>>>(h->g, first, n, seed, d_mat.as<double>(), n_keep, d_keep.as<double>(), d_part.as<SweepPartial>()); break;
        case LAYOUT_PAIR: interpolate_sweep_kernel<LAYOUT_PAIR><<<blocks, SWEEP_THREADS, 0, s>>>(h->g, first, n, seed, d_mat.as<double>(), n_keep, d_keep.as<double>(), d_part.as<SweepPartial>()); break;
        case LAYOUT_BLOCK: interpolate_sweep_kernel<LAYOUT_BLOCK><<<blocks, SWEEP_THREADS, 0, s>>>(h->g, first, n, seed, d_mat.as<double>(), n_keep, d_keep.as<double>(), d_part.as<SweepPartial>()); break;
        default: interpolate_sweep_kernel<LAYOUT_NODE><<<blocks, SWEEP_THREADS, 0, s>>>(h->g, first, n, seed, d_mat.as<double>(), n_keep, d_keep.as<double>(), d_part.as<SweepPartial>()); break;
    }
    count_launch();
    sweep_reduce_kernel<<<1, 32, 0, s>>>(blocks, d_part.as<SweepPartial>(), d_res.as<SweepPartial>());
    count_launch();
    cudaEventRecord(e1, s);
    cudaError_t err = cudaGetLastError();
    SweepPartial r{};
    if (err == cudaSuccess) err = cudaMemcpyAsync(&r, d_res.p, sizeof r, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess && n_keep > 0) err = cudaMemcpyAsync(keep, d_keep.p, sizeof(double) * n_keep, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    float ms = 0.f;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CEG_CK(err);
    *sum = r.sum; *min = r.min; *n_blocked = (int64_t)r.blocked;
    if (kernel_ms) *kernel_ms = ms;
    return 0;
}

// Host-pointer path: queries are streamed through the device in chunks on two streams, so that the
// upload of chunk c+1, the kernels of chunk c and the download of chunk c-1 overlap (PCIe is full
// duplex).  Every grid of `grids` is evaluated at the same positions, which are uploaded once.
static int interpolate_pipelined(int n_grids, ceg_grid_s* const* grids, int64_t n, const double* xyz, double* const* outs) {
    ceg_grid_s* h0 = grids[0];
    std::lock_guard<std::mutex> lk(h0->mu);
    if (use_device(h0->device)) return 1;
    if (!h0->stream2) CEG_CK(acquire_stream(&h0->stream2));
    const int64_t chunk = std::min<int64_t>(n, 1 << 20);
    for (int b = 0; b < 2; b++) {
        CEG_CK(h0->d_xyz[b].reserve(sizeof(double) * 3 * chunk));
        CEG_CK(h0->d_out[b].reserve(sizeof(double) * chunk * n_grids));
    }
    cudaStream_t st[2] = {h0->stream, h0->stream2};
    int c = 0;
    for (int64_t q0 = 0; q0 < n; q0 += chunk, c++) {
        const int64_t m = std::min(chunk, n - q0);
        const int b = c & 1;
        CEG_CK(cudaMemcpyAsync(h0->d_xyz[b].p, xyz + 3 * q0, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st[b]));
        for (int g = 0; g < n_grids; g++) {
            double* d_o = h0->d_out[b].as<double>() + (int64_t)g * chunk;
            CEG_CK(launch_interpolate(grids[g]->g, m, h0->d_xyz[b].as<double>(), d_o, st[b]));
            CEG_CK(cudaMemcpyAsync(outs[g] + q0, d_o, sizeof(double) * m, cudaMemcpyDeviceToHost, st[b]));
        }
    }
    CEG_CK(cudaStreamSynchronize(st[0]));
    CEG_CK(cudaStreamSynchronize(st[1]));
    return 0;
}

static int small_energy_points(const EnergyPointsParams& P, int device, ceg_ewald_h ewald, const double* rest_reim,
                               int64_t n_pos, int n_atoms, const int32_t* atomsidx, const double* charges, const double* xyz,
                               double energy_net_charges, double static_contribution, double* out_vdw, double* out_coulomb);

extern "C" int ceg_interpolate(ceg_grid_h h, int64_t n, const double* xyz, double* out) {
    if (!h || n < 0 || (n > 0 && (!xyz || !out))) return fail("ceg_interpolate: bad argument");
    if (n == 0) return 0;
    if (n <= 256) {     // a handful of points: the one-launch latency path (csrc/latency.cuh), one "atom" of charge 1 per point
        if (use_device(h->device)) return 1;
        EnergyPointsParams P;
        memset(&P, 0, sizeof P);
        if (h->g.is_vdw) { P.n_vdw = 1; P.vdw[0] = h->g; P.vdw_present[0] = 1; }
        else { P.coulomb = h->g; P.has_coulomb = 1; }
        const double one = 1.0;
        double other[256];
        const int rc = small_energy_points(P, h->device, nullptr, nullptr, n, 1, nullptr, &one, xyz, 0.0, 0.0,
                                           h->g.is_vdw ? out : other, h->g.is_vdw ? other : out);
        if (rc != 2) return rc;
    }
    ceg_grid_s* g[1] = {h};
    double* o[1] = {out};
    return interpolate_pipelined(1, g, n, xyz, o);
}

extern "C" int ceg_interpolate_multi(int32_t n_grids, const ceg_grid_h* grids, int64_t n, const double* xyz, double* const* outs) {
    if (n_grids <= 0 || n_grids > 16 || !grids || n < 0 || (n > 0 && (!xyz || !outs)))
        return fail("ceg_interpolate_multi: bad argument");
    for (int g = 0; g < n_grids; g++) {
        if (!grids[g] || (n > 0 && !outs[g])) return fail("ceg_interpolate_multi: null grid or output");
        if (grids[g]->device != grids[0]->device) return fail("ceg_interpolate_multi: grids live on different devices");
    }
    if (n == 0) return 0;
    return interpolate_pipelined(n_grids, grids, n, xyz, outs);
}

// ---- fused (VdW, Coulomb) pair grids --------------------------------------------------------------------------
extern "C" int ceg_grid_pair_create(ceg_grid_h vdw, ceg_grid_h coulomb, ceg_gridpair_h* out) {
    if (!vdw || !coulomb || !out) return fail("ceg_grid_pair_create: null argument");
    if (vdw->device != coulomb->device) return fail("ceg_grid_pair_create: grids live on different devices");
    if (!vdw->g.is_vdw || coulomb->g.is_vdw) return fail("ceg_grid_pair_create: expects (VdW grid, Coulomb grid)");
    for (ceg_grid_s* g : {vdw, coulomb})
        if (g->g.layout != LAYOUT_NODE && g->g.layout != LAYOUT_BLOCK)
            return fail("ceg_grid_pair_create: grids must use the node or brick layout");
    {   // field by field (the structs carry padding)
        const GridSpecDev &a = vdw->g.gs, &b = coulomb->g.gs;
        const CellDev &ca = vdw->g.uc, &cb = coulomb->g.uc;
        bool same = memcmp(a.dims, b.dims, sizeof a.dims) == 0 && memcmp(a.size, b.size, sizeof a.size) == 0 &&
                    memcmp(a.shift, b.shift, sizeof a.shift) == 0 && memcmp(a.delta, b.delta, sizeof a.delta) == 0 &&
                    memcmp(ca.mat, cb.mat, sizeof ca.mat) == 0 && memcmp(ca.invmat, cb.invmat, sizeof ca.invmat) == 0;
        if (!same) return fail("ceg_grid_pair_create: the two grids do not share one GridCoordinatesSetup");
    }
    if (use_device(vdw->device)) return 1;
    std::unique_ptr<ceg_gridpair_s> h(new ceg_gridpair_s);
    h->device = vdw->device;
    h->g.uc = vdw->g.uc;
    h->g.gs = vdw->g.gs;
    h->vdw = vdw;
    h->coul = coulomb;
    const int nx = vdw->g.gs.n[0], ny = vdw->g.gs.n[1], nz = vdw->g.gs.n[2];
    const size_t bricks = (size_t)((nx + 1) / 2) * ((ny + 1) / 2) * ((nz + 1) / 2);
    CEG_CK(acquire_stream(&h->stream));
    CEG_CK(h->d_nodes.reserve(sizeof(Rec8) * 16 * bricks));
    pair_records_kernel<<<sm_count_here() * 32, 256, 0, h->stream>>>(reinterpret_cast<const Rec8*>(vdw->g.data), vdw->g.layout,
                                                                     reinterpret_cast<const Rec8*>(coulomb->g.data),
                                                                     coulomb->g.layout, h->d_nodes.as<Rec8>(), nx, ny, nz);
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaStreamSynchronize(h->stream));
    h->g.data = h->d_nodes.as<float>();
    {
        std::lock_guard<std::mutex> lk(vdw->mu);
        vdw->pair = h.get();       // ceg_energy_points picks the fused records up when it is given this (VdW, Coulomb) couple
    }
    *out = h.release();
    return 0;
}

extern "C" int ceg_grid_pair_free(ceg_gridpair_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->vdw) {
        std::lock_guard<std::mutex> lk(h->vdw->mu);
        if (h->vdw->pair == h) h->vdw->pair = nullptr;
    }
    delete h;
    return 0;
}

extern "C" int ceg_interpolate_pair_dev(ceg_gridpair_h h, int64_t n, const double* d_xyz, double* d_out_vdw, double* d_out_coul,
                                        void* stream) {
    if (!h || n < 0 || (n > 0 && (!d_xyz || !d_out_vdw || !d_out_coul))) return fail("ceg_interpolate_pair_dev: bad argument");
    if (n == 0) return 0;
    if (use_device(h->device)) return 1;
    interpolate_pair_kernel<<<interp_blocks(n), 256, 0, (cudaStream_t)stream>>>(h->g, n, d_xyz, d_out_vdw, d_out_coul);
    count_launch();
    CEG_CK(cudaGetLastError());
    return 0;
}

// host pointers: chunks on two streams, upload / kernel / two downloads overlapped (as interpolate_pipelined)
extern "C" int ceg_interpolate_pair(ceg_gridpair_h h, int64_t n, const double* xyz, double* out_vdw, double* out_coul) {
    if (!h || n < 0 || (n > 0 && (!xyz || !out_vdw || !out_coul))) return fail("ceg_interpolate_pair: bad argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    if (!h->stream2) CEG_CK(acquire_stream(&h->stream2));
    const int64_t chunk = std::min<int64_t>(n, 1 << 20);
    for (int b = 0; b < 2; b++) {
        CEG_CK(h->d_xyz[b].reserve(sizeof(double) * 3 * chunk));
        CEG_CK(h->d_out[b].reserve(sizeof(double) * 2 * chunk));
    }
    cudaStream_t st[2] = {h->stream, h->stream2};
    int c = 0;
    for (int64_t q0 = 0; q0 < n; q0 += chunk, c++) {
        const int64_t m = std::min(chunk, n - q0);
        const int b = c & 1;
        double* d_v = h->d_out[b].as<double>();
        double* d_c = d_v + chunk;
        CEG_CK(cudaMemcpyAsync(h->d_xyz[b].p, xyz + 3 * q0, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st[b]));
        interpolate_pair_kernel<<<interp_blocks(m), 256, 0, st[b]>>>(h->g, m, h->d_xyz[b].as<double>(), d_v, d_c);
        count_launch();
        CEG_CK(cudaGetLastError());
        CEG_CK(cudaMemcpyAsync(out_vdw + q0, d_v, sizeof(double) * m, cudaMemcpyDeviceToHost, st[b]));
        CEG_CK(cudaMemcpyAsync(out_coul + q0, d_c, sizeof(double) * m, cudaMemcpyDeviceToHost, st[b]));
    }
    CEG_CK(cudaStreamSynchronize(st[0]));
    CEG_CK(cudaStreamSynchronize(st[1]));
    return 0;
}

// ================================================================================================
// Ewald
// ================================================================================================
extern "C" int ceg_ewald_create(int32_t device, const double* invmat, const int32_t kmax[3], int64_t n_k,
                                const int32_t* kvec, const double* kfactor, int64_t n_sites, const double* frac,
                                const double* charge, ceg_ewald_h* out, double* S_f_reim) {
    if (!invmat || !kmax || !out || n_k < 0 || n_sites < 0) return fail("ceg_ewald_create: bad argument");
    if (n_k > 0 && (!kvec || !kfactor)) return fail("ceg_ewald_create: null k-list");
    if (n_sites > 0 && (!frac || !charge)) return fail("ceg_ewald_create: null sites");
    for (int64_t k = 0; k < n_k; k++)
        if (kvec[3 * k] < 0 || kvec[3 * k] > kmax[0] || std::abs(kvec[3 * k + 1]) > kmax[1] || std::abs(kvec[3 * k + 2]) > kmax[2])
            return fail("ceg_ewald_create: k-vector outside kmax");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_ewald_s> h(new ceg_ewald_s);
    h->device = device;
    memcpy(h->invmat, invmat, sizeof h->invmat);
    for (int d = 0; d < 3; d++) h->kmax[d] = kmax[d];
    h->n_k = n_k;
    CEG_CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CEG_CK(h->d_kv.reserve(sizeof(KVecDev) * std::max<int64_t>(1, n_k)));
    if (n_k > 0) {
        DevBuf d_kvec, d_frac, d_q, d_part, d_S;
        DevBuf& d_kf = h->d_kf;
        const int n_blocks = (int)std::max<int64_t>(1, (n_sites + SF_SITES_PER_BLOCK - 1) / SF_SITES_PER_BLOCK);
        CEG_CK(d_kvec.reserve(sizeof(int) * 3 * n_k));
        CEG_CK(d_kf.reserve(sizeof(double) * n_k));
        CEG_CK(d_frac.reserve(sizeof(double) * 3 * std::max<int64_t>(1, n_sites)));
        CEG_CK(d_q.reserve(sizeof(double) * std::max<int64_t>(1, n_sites)));
        CEG_CK(d_part.reserve(sizeof(double2) * n_k * n_blocks));
        CEG_CK(d_S.reserve(sizeof(double) * 2 * n_k));
        CEG_CK(cudaMemcpyAsync(d_kvec.p, kvec, sizeof(int) * 3 * n_k, cudaMemcpyHostToDevice, h->stream));
        CEG_CK(cudaMemcpyAsync(d_kf.p, kfactor, sizeof(double) * n_k, cudaMemcpyHostToDevice, h->stream));
        if (n_sites > 0) {
            CEG_CK(cudaMemcpyAsync(d_frac.p, frac, sizeof(double) * 3 * n_sites, cudaMemcpyHostToDevice, h->stream));
            CEG_CK(cudaMemcpyAsync(d_q.p, charge, sizeof(double) * n_sites, cudaMemcpyHostToDevice, h->stream));
        }
        dim3 grid((unsigned)((n_k + SF_THREADS - 1) / SF_THREADS), (unsigned)n_blocks);
        structure_factor_kernel<<<grid, SF_THREADS, 0, h->stream>>>(n_k, d_kvec.as<int>(), n_sites, d_frac.as<double>(),
                                                                   d_q.as<double>(), d_part.as<double2>());
        count_launch();
        CEG_CK(cudaGetLastError());
        structure_factor_reduce_kernel<<<(unsigned)((n_k + 127) / 128), 128, 0, h->stream>>>(
            n_k, n_blocks, d_part.as<double2>(), d_kf.as<double>(), d_kvec.as<int>(), h->d_kv.as<KVecDev>(), d_S.as<double>());
        count_launch();
        CEG_CK(cudaGetLastError());
        if (S_f_reim) CEG_CK(cudaMemcpyAsync(S_f_reim, d_S.p, sizeof(double) * 2 * n_k, cudaMemcpyDeviceToHost, h->stream));
        // rows of the k-list (ewald.jl:213-236: one contiguous ix range per (jy, jz)); any other list keeps the
        // warp-per-placement kernel
        const int ny = 2 * kmax[1] + 1, nz = 2 * kmax[2] + 1;
        std::vector<RowDev> rows((size_t)ny * nz, RowDev{0, 0, 0});
        bool ok = true;
        for (int64_t k = 0; k < n_k && ok; k++) {
            RowDev& r = rows[(size_t)(kvec[3 * k + 1] + kmax[1]) * nz + (kvec[3 * k + 2] + kmax[2])];
            if (r.count == 0) { r.ix0 = kvec[3 * k]; r.count = 1; r.offset = k; }
            else if (k == r.offset + r.count && kvec[3 * k] == r.ix0 + r.count) r.count++;
            else ok = false;
        }
        h->rows_ok = ok;
        h->sum_kf = 0.0;
        for (int64_t k = 0; k < n_k; k++) h->sum_kf += kfactor[k];
        if (ok) {
            CEG_CK(h->d_rows.reserve(sizeof(RowDev) * rows.size()));
            CEG_CK(h->d_ksf.reserve(sizeof(double2) * n_k));
            CEG_CK(cudaMemcpyAsync(h->d_rows.p, rows.data(), sizeof(RowDev) * rows.size(), cudaMemcpyHostToDevice, h->stream));
            scale_structure_factor_kernel<<<(unsigned)((n_k + 127) / 128), 128, 0, h->stream>>>(
                n_k, d_kf.as<double>(), d_S.as<double2>(), h->d_ksf.as<double2>());
            count_launch();
            CEG_CK(cudaGetLastError());
        }
        CEG_CK(cudaStreamSynchronize(h->stream));
    }
    *out = h.release();
    return 0;
}

// The same handle from a structure factor that already exists on the host (EwaldFramework.StoreRigidChargeFramework of
// an unchanged initialize_ewald, ewald.jl:264-271): nothing is summed on the device.
extern "C" int ceg_ewald_create_from_sf(int32_t device, const double* invmat, const int32_t kmax[3], int64_t n_k,
                                        const int32_t* kvec, const double* kfactor, const double* S_f_reim, ceg_ewald_h* out) {
    if (!invmat || !kmax || !out || n_k < 0) return fail("ceg_ewald_create_from_sf: bad argument");
    if (n_k > 0 && (!kvec || !kfactor || !S_f_reim)) return fail("ceg_ewald_create_from_sf: null k-list");
    for (int64_t k = 0; k < n_k; k++)
        if (kvec[3 * k] < 0 || kvec[3 * k] > kmax[0] || std::abs(kvec[3 * k + 1]) > kmax[1] || std::abs(kvec[3 * k + 2]) > kmax[2])
            return fail("ceg_ewald_create_from_sf: k-vector outside kmax");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_ewald_s> h(new ceg_ewald_s);
    h->device = device;
    memcpy(h->invmat, invmat, sizeof h->invmat);
    for (int d = 0; d < 3; d++) h->kmax[d] = kmax[d];
    h->n_k = n_k;
    CEG_CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CEG_CK(h->d_kv.reserve(sizeof(KVecDev) * std::max<int64_t>(1, n_k)));
    if (n_k > 0) {
        std::vector<KVecDev> kv((size_t)n_k);
        std::vector<double2> ksf((size_t)n_k);
        const int nz = 2 * kmax[2] + 1, ny = 2 * kmax[1] + 1;
        std::vector<RowDev> rows((size_t)ny * nz, RowDev{0, 0, 0});
        bool ok = true;
        h->sum_kf = 0.0;
        for (int64_t k = 0; k < n_k; k++) {
            kv[k].kf = kfactor[k]; kv[k].sf_re = S_f_reim[2 * k]; kv[k].sf_im = S_f_reim[2 * k + 1];
            kv[k].ix = kvec[3 * k]; kv[k].jy = kvec[3 * k + 1]; kv[k].jz = kvec[3 * k + 2]; kv[k]._pad = 0;
            ksf[k] = make_double2(kfactor[k] * S_f_reim[2 * k], kfactor[k] * S_f_reim[2 * k + 1]);
            h->sum_kf += kfactor[k];
            if (ok) {
                RowDev& r = rows[(size_t)(kvec[3 * k + 1] + kmax[1]) * nz + (kvec[3 * k + 2] + kmax[2])];
                if (r.count == 0) { r.ix0 = kvec[3 * k]; r.count = 1; r.offset = k; }
                else if (k == r.offset + r.count && kvec[3 * k] == r.ix0 + r.count) r.count++;
                else ok = false;
            }
        }
        h->rows_ok = ok;
        CEG_CK(h->d_kf.reserve(sizeof(double) * n_k));
        CEG_CK(cudaMemcpyAsync(h->d_kv.p, kv.data(), sizeof(KVecDev) * n_k, cudaMemcpyHostToDevice, h->stream));
        CEG_CK(cudaMemcpyAsync(h->d_kf.p, kfactor, sizeof(double) * n_k, cudaMemcpyHostToDevice, h->stream));
        if (ok) {
            CEG_CK(h->d_rows.reserve(sizeof(RowDev) * rows.size()));
            CEG_CK(h->d_ksf.reserve(sizeof(double2) * n_k));
            CEG_CK(cudaMemcpyAsync(h->d_rows.p, rows.data(), sizeof(RowDev) * rows.size(), cudaMemcpyHostToDevice, h->stream));
            CEG_CK(cudaMemcpyAsync(h->d_ksf.p, ksf.data(), sizeof(double2) * n_k, cudaMemcpyHostToDevice, h->stream));
        }
        CEG_CK(cudaStreamSynchronize(h->stream));
    }
    *out = h.release();
    return 0;
}

extern "C" int ceg_ewald_free(ceg_ewald_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    delete h;
    return 0;
}

// q_host (optional): the charges on the host; molecules of up to RECIP_QV atoms then carry them by value in the kernel
// parameters instead of reading d_q
static int launch_reciprocal(ceg_ewald_s* h, int64_t n_pos, int n_atoms, const double* d_xyz, const double* d_q,
                             double enc, double stat, double* d_out, const uint8_t* d_skip, int accumulate,
                             cudaStream_t s, const double2* d_rest = nullptr, const double* q_host = nullptr) {
    // Small batches (single Monte-Carlo steps) do not fill the GPU with one thread per placement, and that thread walks
    // the whole k-list alone; the warp-per-placement kernel spreads the k-list over 32 lanes and returns sooner.
    constexpr int64_t ROWS_MIN_PLACEMENTS = 16384;
    if (h->rows_ok && n_atoms <= RECIP_QV && n_pos >= ROWS_MIN_PLACEMENTS) {
        RowsParams R;
        R.n_atoms = n_atoms;
        memcpy(R.invmat, h->invmat, sizeof R.invmat);
        R.ky = h->kmax[1]; R.kz = h->kmax[2];
        R.rows = h->d_rows.as<RowDev>();
        R.kf = h->d_kf.as<double>();
        R.ksf = h->d_ksf.as<double2>();
        R.sum_kf = h->sum_kf;
        if (d_rest) {   // single_contribution_ewald: the framework sums are replaced by `rest`
            CEG_CK(h->d_ksf_rest.reserve(sizeof(double2) * h->n_k));
            scale_structure_factor_kernel<<<(unsigned)((h->n_k + 127) / 128), 128, 0, s>>>(h->n_k, R.kf, d_rest,
                                                                                         h->d_ksf_rest.as<double2>());
            count_launch();
            CEG_CK(cudaGetLastError());
            R.ksf = h->d_ksf_rest.as<double2>();
        }
        R.n_pos = n_pos; R.xyz = d_xyz; R.q = d_q;
        if (q_host) { R.q = nullptr; for (int a = 0; a < n_atoms; a++) R.qv[a] = q_host[a]; }
        R.energy_net_charges = enc; R.static_contribution = stat;
        R.out = d_out; R.skip = d_skip; R.accumulate = accumulate;
        const unsigned blocks = (unsigned)((n_pos + ROWS_THREADS - 1) / ROWS_THREADS);
        const int groups = n_atoms <= 8 ? 2 : n_atoms <= 16 ? 4 : 8;             // lanes per placement beyond 4 atoms
        const unsigned gblocks = (unsigned)((n_pos * groups + ROWS_THREADS - 1) / ROWS_THREADS);
        if (n_atoms <= ROWS_MAX_ATOMS) {
            switch (n_atoms) {
                case 1: reciprocal_rows_kernel<1><<<blocks, ROWS_THREADS, 0, s>>>(R); break;
                case 2: reciprocal_rows_kernel<2><<<blocks, ROWS_THREADS, 0, s>>>(R); break;
                case 3: reciprocal_rows_kernel<3><<<blocks, ROWS_THREADS, 0, s>>>(R); break;
                default: reciprocal_rows_kernel<4><<<blocks, ROWS_THREADS, 0, s>>>(R); break;
            }
        } else if (groups == 2) {
            reciprocal_rows_group_kernel<2><<<gblocks, ROWS_THREADS, 0, s>>>(R);
        } else if (groups == 4) {
            reciprocal_rows_group_kernel<4><<<gblocks, ROWS_THREADS, 0, s>>>(R);
        } else {
            reciprocal_rows_group_kernel<8><<<gblocks, ROWS_THREADS, 0, s>>>(R);
        }
        count_launch();
        CEG_CK(cudaGetLastError());
        return 0;
    }
    RecipParams P;
    P.rest = d_rest;
    memcpy(P.invmat, h->invmat, sizeof P.invmat);
    P.kx = h->kmax[0]; P.ky = h->kmax[1]; P.kz = h->kmax[2];
    P.n_k = h->n_k;
    P.kv = h->d_kv.as<KVecDev>();
    P.n_pos = n_pos; P.n_atoms = n_atoms;
    P.xyz = d_xyz; P.q = d_q;
    if (q_host && n_atoms <= RECIP_QV) { P.q = nullptr; for (int a = 0; a < n_atoms; a++) P.qv[a] = q_host[a]; }
    P.energy_net_charges = enc; P.static_contribution = stat;
    P.out = d_out; P.skip = d_skip; P.accumulate = accumulate;
    const size_t per_atom = (size_t)(P.kx + 1) + (2 * P.ky + 1) + (2 * P.kz + 1);
    const size_t smem = sizeof(double2) * per_atom * n_atoms * RECIP_WARPS;
    if (smem > 200 * 1024) return fail("ceg_ewald_reciprocal: molecule too large for the shared-memory phase tables");
    if (smem > 48 * 1024)
        CEG_CK(cudaFuncSetAttribute(reciprocal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n_pos + RECIP_WARPS - 1) / RECIP_WARPS, sm_count_here() * 16));
    reciprocal_kernel<<<blocks, RECIP_WARPS * 32, smem, s>>>(P);
    count_launch();
    CEG_CK(cudaGetLastError());
    return 0;
}

extern "C" int ceg_ewald_reciprocal_dev(ceg_ewald_h h, int64_t n_pos, int32_t n_atoms, const double* d_xyz,
                                        const double* q_host, double energy_net_charges, double static_contribution,
                                        double* d_out, void* stream) {
    if (!h || n_pos < 0 || n_atoms <= 0 || !q_host || (n_pos > 0 && (!d_xyz || !d_out)))
        return fail("ceg_ewald_reciprocal_dev: bad argument");
    if (n_pos == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_atoms > RECIP_QV) {
        // large molecules: handle-owned charge buffer — calls with different charges must then share one stream
        CEG_CK(h->d_q.reserve(sizeof(double) * n_atoms));
        CEG_CK(cudaMemcpyAsync(h->d_q.p, q_host, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, s));
    }
    return launch_reciprocal(h, n_pos, n_atoms, d_xyz, h->d_q.as<double>(), energy_net_charges, static_contribution,
                             d_out, nullptr, 0, s, nullptr, q_host);
}

static int small_energy_points(const EnergyPointsParams& P, int device, ceg_ewald_h ewald, const double* rest_reim,
                               int64_t n_pos, int n_atoms, const int32_t* atomsidx, const double* charges, const double* xyz,
                               double energy_net_charges, double static_contribution, double* out_vdw, double* out_coulomb);
// the reciprocal-space term alone through the one-launch latency path (no grids, no block mask)
static int small_reciprocal(ceg_ewald_h h, const double* rest_reim, int64_t n_pos, int n_atoms, const double* xyz, const double* q,
                            double energy_net_charges, double static_contribution, double* out) {
    EnergyPointsParams P;
    memset(&P, 0, sizeof P);
    return small_energy_points(P, h->device, h, rest_reim, n_pos, n_atoms, nullptr, q, xyz, energy_net_charges,
                               static_contribution, nullptr, out);
}

extern "C" int ceg_ewald_reciprocal(ceg_ewald_h h, int64_t n_pos, int32_t n_atoms, const double* xyz, const double* q,
                                    double energy_net_charges, double static_contribution, double* out) {
    if (!h || n_pos < 0 || n_atoms <= 0 || !q || (n_pos > 0 && (!xyz || !out)))
        return fail("ceg_ewald_reciprocal: bad argument");
    if (n_pos == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    {   // single Monte-Carlo steps: one launch, mapped I/O (csrc/latency.cuh)
        const int rc = small_reciprocal(h, nullptr, n_pos, n_atoms, xyz, q, energy_net_charges, static_contribution, out);
        if (rc != 2) return rc;
    }
    CEG_CK(h->d_xyz.reserve(sizeof(double) * 3 * n_atoms * n_pos));
    CEG_CK(h->d_out.reserve(sizeof(double) * n_pos));
    CEG_CK(h->d_q.reserve(sizeof(double) * n_atoms));
    CEG_CK(cudaMemcpyAsync(h->d_xyz.p, xyz, sizeof(double) * 3 * n_atoms * n_pos, cudaMemcpyHostToDevice, h->stream));
    CEG_CK(cudaMemcpyAsync(h->d_q.p, q, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, h->stream));
    if (launch_reciprocal(h, n_pos, n_atoms, h->d_xyz.as<double>(), h->d_q.as<double>(), energy_net_charges,
                          static_contribution, h->d_out.as<double>(), nullptr, 0, h->stream))
        return 1;
    CEG_CK(cudaMemcpyAsync(out, h->d_out.p, sizeof(double) * n_pos, cudaMemcpyDeviceToHost, h->stream));
    CEG_CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// single_contribution_ewald (ewald.jl:704-738) for many trial placements of ONE molecule:
// out = 2 sum_k kf Re(rest conj S_single) + sum_k kf |S_single|^2, where `rest` = framework + all other
// adsorbates (StoreRigidChargeFramework + sums[:,1] - sums[:,ij+1]) is supplied by the caller.
extern "C" int ceg_ewald_single_contribution(ceg_ewald_h h, const double* rest_reim, int64_t n_pos, int32_t n_atoms,
                                             const double* xyz, const double* q, double* out) {
    if (!h || !rest_reim || n_pos < 0 || n_atoms <= 0 || !q || (n_pos > 0 && (!xyz || !out)))
        return fail("ceg_ewald_single_contribution: bad argument");
    if (n_pos == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    {   // single Monte-Carlo steps: one launch, `rest` read from mapped host memory (csrc/latency.cuh)
        const int rc = small_reciprocal(h, rest_reim, n_pos, n_atoms, xyz, q, 0.0, 0.0, out);
        if (rc != 2) return rc;
    }
    DevBuf d_rest;
    CEG_CK(d_rest.reserve(sizeof(double) * 2 * std::max<int64_t>(1, h->n_k)));
    CEG_CK(h->d_xyz.reserve(sizeof(double) * 3 * n_atoms * n_pos));
    CEG_CK(h->d_out.reserve(sizeof(double) * n_pos));
    CEG_CK(h->d_q.reserve(sizeof(double) * n_atoms));
    if (h->n_k > 0)
        CEG_CK(cudaMemcpyAsync(d_rest.p, rest_reim, sizeof(double) * 2 * h->n_k, cudaMemcpyHostToDevice, h->stream));
    CEG_CK(cudaMemcpyAsync(h->d_xyz.p, xyz, sizeof(double) * 3 * n_atoms * n_pos, cudaMemcpyHostToDevice, h->stream));
    CEG_CK(cudaMemcpyAsync(h->d_q.p, q, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, h->stream));
    // host = 2*(fa + 0), ads = aa + 0  ->  2*rest_single + single_single (ewald.jl:737)
    if (launch_reciprocal(h, n_pos, n_atoms, h->d_xyz.as<double>(), h->d_q.as<double>(), 0.0, 0.0, h->d_out.as<double>(),
                          nullptr, 0, h->stream, d_rest.as<double2>()))
        return 1;
    CEG_CK(cudaMemcpyAsync(out, h->d_out.p, sizeof(double) * n_pos, cudaMemcpyDeviceToHost, h->stream));
    CEG_CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ================================================================================================
// guest-guest pair energies (energy.jl:339-427)
// ================================================================================================
// Device-resident guest environment (SimulationStep.positions / atoms of the other guests, energy.jl:43-60): created
// once, evaluated for many batches of trial placements, updated in place after accepted moves.
struct ceg_guest_s {
    int device = 0;
    std::mutex mu;
    cudaStream_t stream = nullptr;
    CellDev cell;
    double cutoff2 = 0;
    int n_species = 0;
    int64_t n_env = 0;
    DevBuf d_tab, d_exyz, d_esp, d_emol, d_msp, d_txyz, d_out;
    cudaEvent_t ev[2] = {nullptr, nullptr};      // around the kernel of the last ceg_guest_env_energy
    float last_kernel_ms = 0.f;
    ~ceg_guest_s() {
        for (auto e : ev) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

extern "C" int ceg_guest_env_create(int32_t device, const ceg_cell_t* cell, double cutoff2, int32_t n_species,
                                    const ceg_pair_energy_t* table, int64_t n_env, const double* env_xyz,
                                    const int32_t* env_species, const int32_t* env_mol, ceg_guest_h* out) {
    if (!cell || n_species <= 0 || !table || n_env < 0 || !out || (n_env > 0 && (!env_xyz || !env_species || !env_mol)))
        return fail("ceg_guest_env_create: bad argument");
    if (n_env > (1ll << 30)) return fail("ceg_guest_env_create: too many environment atoms");
    for (int64_t i = 0; i < n_env; i++)
        if (env_species[i] < 0 || env_species[i] >= n_species) return fail("ceg_guest_env_create: species out of range");
    for (int i = 0; i < n_species * n_species; i++)
        if (table[i].flags < 0) return fail("Undefined interaction");
    if (use_device(device)) return 1;
    std::vector<PairEnergyDev> tab((size_t)n_species * n_species);
    for (size_t i = 0; i < tab.size(); i++) {
        const ceg_pair_energy_t& s = table[i];
        PairEnergyDev& d = tab[i];
        d.flags = s.flags; d._pad = 0;
        d.hs_r2 = s.hs_r * s.hs_r;
        d.lj_eps = s.lj_eps; d.lj_sig2 = s.lj_sigma * s.lj_sigma; d.lj_shift = s.lj_shift;
        d.bA = s.buck_A; d.bB = s.buck_B; d.bC = s.buck_C; d.b_shift = s.buck_shift;
        d.ew_alpha = s.ewald_alpha; d.ew_qq = s.ewald_qq; d.coul_qq = s.coulomb_qq;
    }
    std::unique_ptr<ceg_guest_s> h(new ceg_guest_s);
    h->device = device;
    h->cell = to_dev(*cell);
    h->cutoff2 = cutoff2;
    h->n_species = n_species;
    h->n_env = n_env;
    CEG_CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CEG_CK(h->d_tab.reserve(sizeof(PairEnergyDev) * tab.size()));
    CEG_CK(h->d_exyz.reserve(sizeof(double) * 3 * std::max<int64_t>(1, n_env)));
    CEG_CK(h->d_esp.reserve(sizeof(int) * std::max<int64_t>(1, n_env)));
    CEG_CK(h->d_emol.reserve(sizeof(int) * std::max<int64_t>(1, n_env)));
    CEG_CK(cudaMemcpyAsync(h->d_tab.p, tab.data(), sizeof(PairEnergyDev) * tab.size(), cudaMemcpyHostToDevice, h->stream));
    if (n_env > 0) {
        CEG_CK(cudaMemcpyAsync(h->d_exyz.p, env_xyz, sizeof(double) * 3 * n_env, cudaMemcpyHostToDevice, h->stream));
        CEG_CK(cudaMemcpyAsync(h->d_esp.p, env_species, sizeof(int) * n_env, cudaMemcpyHostToDevice, h->stream));
        CEG_CK(cudaMemcpyAsync(h->d_emol.p, env_mol, sizeof(int) * n_env, cudaMemcpyHostToDevice, h->stream));
    }
    CEG_CK(cudaStreamSynchronize(h->stream));
    *out = h.release();
    return 0;
}

extern "C" int ceg_guest_env_free(ceg_guest_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    delete h;
    return 0;
}

// positions of environment atoms [first, first + n) after an accepted move
extern "C" int ceg_guest_env_update(ceg_guest_h h, int64_t first, int64_t n, const double* xyz) {
    if (!h || first < 0 || n < 0 || first + n > h->n_env || (n > 0 && !xyz)) return fail("ceg_guest_env_update: bad argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    CEG_CK(cudaMemcpyAsync(h->d_exyz.as<double>() + 3 * first, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    CEG_CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int ceg_guest_env_energy(ceg_guest_h h, int64_t n_trials, int32_t n_atoms, const int32_t* mol_species,
                                    int32_t skip_mol, const double* trial_xyz, double* out) {
    if (!h || n_trials < 0 || n_atoms <= 0 || !mol_species || (n_trials > 0 && (!trial_xyz || !out)))
        return fail("ceg_guest_env_energy: bad argument");
    for (int a = 0; a < n_atoms; a++)
        if (mol_species[a] < 0 || mol_species[a] >= h->n_species) return fail("ceg_guest_env_energy: species out of range");
    if (n_trials == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (use_device(h->device)) return 1;
    cudaStream_t s = h->stream;
    CEG_CK(h->d_msp.reserve(sizeof(int) * n_atoms));
    CEG_CK(h->d_txyz.reserve(sizeof(double) * 3 * n_atoms * n_trials));
    CEG_CK(h->d_out.reserve(sizeof(double) * n_trials));
    CEG_CK(cudaMemcpyAsync(h->d_msp.p, mol_species, sizeof(int) * n_atoms, cudaMemcpyHostToDevice, s));
    CEG_CK(cudaMemcpyAsync(h->d_txyz.p, trial_xyz, sizeof(double) * 3 * n_atoms * n_trials, cudaMemcpyHostToDevice, s));
    GuestParams P;
    P.cell = h->cell; P.cutoff2 = h->cutoff2; P.n_species = h->n_species; P.table = h->d_tab.as<PairEnergyDev>();
    P.cut_lo = h->cutoff2 - GUEST_BAND; P.cut_hi = h->cutoff2 + GUEST_BAND;
    P.fast_ok = (2.0 * std::sqrt(h->cutoff2) <= perp_width_min(h->cell.mat) * (1.0 - 1e-9)) ? 1 : 0;
    P.n_env = (int)h->n_env; P.env_xyz = h->d_exyz.as<double>(); P.env_species = h->d_esp.as<int>(); P.env_mol = h->d_emol.as<int>();
    P.n_trials = n_trials; P.n_atoms = n_atoms; P.mol_species = h->d_msp.as<int>(); P.skip_mol = skip_mol;
    P.trial_xyz = h->d_txyz.as<double>(); P.out = h->d_out.as<double>();
    for (auto& e : h->ev) if (!e) CEG_CK(cudaEventCreate(&e));
    CEG_CK(cudaEventRecord(h->ev[0], s));
    if (P.fast_ok) {
        const int tpt = std::min<int>(n_atoms, 32);                  // threads sharing one trial placement
        const int per_cta = GUEST_THREADS / tpt;
        const unsigned nb = (unsigned)((n_trials + per_cta - 1) / per_cta);
        if (h->cell.mat[1] == 0.0 && h->cell.mat[2] == 0.0 && h->cell.mat[5] == 0.0)
            guest_energy_fast_kernel<true><<<nb, GUEST_THREADS, 0, s>>>(P, tpt);
        else
            guest_energy_fast_kernel<false><<<nb, GUEST_THREADS, 0, s>>>(P, tpt);
    } else {
        guest_energy_kernel<<<(unsigned)((n_trials + 127) / 128), 128, 0, s>>>(P);
    }
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaEventRecord(h->ev[1], s));
    CEG_CK(cudaMemcpyAsync(out, h->d_out.p, sizeof(double) * n_trials, cudaMemcpyDeviceToHost, s));
    CEG_CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&h->last_kernel_ms, h->ev[0], h->ev[1]);
    return 0;
}

// Device time (ms, CUDA events on the handle's stream) of the kernel of the last ceg_guest_env_energy call.
extern "C" int ceg_guest_env_last_kernel_ms(ceg_guest_h h, double* ms) {
    if (!h || !ms) return fail("ceg_guest_env_last_kernel_ms: null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *ms = h->last_kernel_ms;
    return 0;
}

// Stateless form: environment uploaded, evaluated and released in one call.
extern "C" int ceg_guest_energy(int32_t device, const ceg_cell_t* cell, double cutoff2, int32_t n_species,
                                const ceg_pair_energy_t* table, int64_t n_env, const double* env_xyz,
                                const int32_t* env_species, const int32_t* env_mol, int64_t n_trials, int32_t n_atoms,
                                const int32_t* mol_species, int32_t skip_mol, const double* trial_xyz, double* out) {
    if (n_trials < 0 || n_atoms <= 0 || !mol_species || (n_trials > 0 && (!trial_xyz || !out)))
        return fail("ceg_guest_energy: bad argument");
    ceg_guest_h h = nullptr;
    if (ceg_guest_env_create(device, cell, cutoff2, n_species, table, n_env, env_xyz, env_species, env_mol, &h)) return 1;
    const int rc = ceg_guest_env_energy(h, n_trials, n_atoms, mol_species, skip_mol, trial_xyz, out);
    ceg_guest_env_free(h);
    return rc;
}

// ================================================================================================
// block mask + fused energy_point
// ================================================================================================
extern "C" int ceg_block_upload(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs,
                                const uint8_t* mask, ceg_block_h* out) {
    if (!unitcell || !gs || !mask || !out) return fail("ceg_block_upload: null argument");
    if (use_device(device)) return 1;
    std::unique_ptr<ceg_block_s> h(new ceg_block_s);
    h->device = device;
    h->b.uc = to_dev(*unitcell);
    h->b.gs = to_dev(*gs);
    const int64_t npts = (int64_t)(gs->dims[0] + 1) * (gs->dims[1] + 1) * (gs->dims[2] + 1);
    CEG_CK(h->d_mask.reserve(npts));
    CEG_CK(cudaMemcpy(h->d_mask.p, mask, npts, cudaMemcpyHostToDevice));
    h->b.mask = h->d_mask.as<uint8_t>();
    *out = h.release();
    return 0;
}

// parse_blockfile's projection of blocking spheres on the grid nodes (coordinates.jl:112-167).
extern "C" int ceg_block_from_spheres(int32_t device, const ceg_cell_t* unitcell, const ceg_gridspec_t* gs,
                                      int32_t n_spheres, const double* centers, const double* radius2, uint8_t* mask) {
    if (!unitcell || !gs || !mask || n_spheres < 0 || (n_spheres > 0 && (!centers || !radius2)))
        return fail("ceg_block_from_spheres: bad argument");
    if (use_device(device)) return 1;
    const int64_t npts = (int64_t)(gs->dims[0] + 1) * (gs->dims[1] + 1) * (gs->dims[2] + 1);
    DevBuf d_c, d_r, d_m;
    CEG_CK(d_c.reserve(sizeof(double) * 3 * std::max(1, n_spheres)));
    CEG_CK(d_r.reserve(sizeof(double) * std::max(1, n_spheres)));
    CEG_CK(d_m.reserve(npts));
    if (n_spheres > 0) {
        CEG_CK(cudaMemcpy(d_c.p, centers, sizeof(double) * 3 * n_spheres, cudaMemcpyHostToDevice));
        CEG_CK(cudaMemcpy(d_r.p, radius2, sizeof(double) * n_spheres, cudaMemcpyHostToDevice));
    }
    block_spheres_kernel<<<sm_count_here() * 8, 256>>>(to_dev(*unitcell), to_dev(*gs), n_spheres, d_c.as<double>(), d_r.as<double>(),
                                          d_m.as<uint8_t>());
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaMemcpy(mask, d_m.p, npts, cudaMemcpyDeviceToHost));
    return 0;
}

// BlockFile(g::EnergyGrid) (grids.jl:188-204) from the value channel [nx][ny][nz] of a VdW grid.
extern "C" int ceg_block_from_values(int32_t device, const ceg_gridspec_t* gs, const float* values, float threshold,
                                     uint8_t* mask) {
    if (!gs || !values || !mask) return fail("ceg_block_from_values: bad argument");
    if (use_device(device)) return 1;
    const int nx = gs->dims[0] + 1, ny = gs->dims[1] + 1, nz = gs->dims[2] + 1;
    const int64_t npts = (int64_t)nx * ny * nz;
    DevBuf d_v, d_m;
    CEG_CK(d_v.reserve(sizeof(float) * npts));
    CEG_CK(d_m.reserve(npts));
    CEG_CK(cudaMemcpy(d_v.p, values, sizeof(float) * npts, cudaMemcpyHostToDevice));
    block_from_values_kernel<<<sm_count_here() * 8, 256>>>(d_v.as<float>(), threshold, nx, ny, nz, d_m.as<uint8_t>());
    count_launch();
    CEG_CK(cudaGetLastError());
    CEG_CK(cudaMemcpy(mask, d_m.p, npts, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int ceg_block_free(ceg_block_h h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    delete h;
    return 0;
}

// Per-thread pinned staging block (mapped into the device address space) and stream for small
// ceg_energy_points calls (one device per thread at a time).
struct SmallStage {
    static constexpr size_t BYTES = 1u << 20;
    void* host = nullptr;
    void* host_dev = nullptr;          // device view of `host`
    unsigned long long gen = 0;
    cudaStream_t stream = nullptr;
    int device = -1;
    bool ready(int dev) {
        if (dev < 0) return false;
        if (device != dev) {                       // (re)create for this device
            if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
            if (host) { cudaFreeHost(host); host = nullptr; }
            if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { stream = nullptr; cudaGetLastError(); return false; }
            device = dev;
        }
        if (!host) {
            if (cudaHostAlloc(&host, BYTES, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { host = nullptr; cudaGetLastError(); return false; }
            if (cudaHostGetDevicePointer(&host_dev, host, 0) != cudaSuccess) { cudaFreeHost(host); host = nullptr; cudaGetLastError(); return false; }
            memset(host, 0, BYTES);                 // result slots carry generation numbers: none matches before its call
        }
        return true;
    }
    ~SmallStage() {                                 // thread exit; the context may already be gone: ignore errors
        if (host) cudaFreeHost(host);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
    }
};
static SmallStage& small_stage() {
    static thread_local SmallStage st;
    return st;
}
// Latency path (single Monte-Carlo steps): ONE kernel launch per call.  Small inputs travel inside the kernel
// parameters, larger ones and all outputs sit in a pinned block mapped into the device address space (the kernel
// writes the energies over PCIe itself) and the host polls the result slots — no copies, no stream synchronisation,
// no completion counter (csrc/latency.cuh).  Serves ceg_energy_points, ceg_ewald_reciprocal and
// ceg_ewald_single_contribution (`rest`: the structure factor the placement interacts with instead of the framework's).
// Returns 0 done, 1 error, 2 not applicable (the caller takes its batch path).
static int small_energy_points(const EnergyPointsParams& P, int device, ceg_ewald_h ewald, const double* rest_reim,
                               int64_t n_pos, int n_atoms, const int32_t* atomsidx, const double* charges, const double* xyz,
                               double energy_net_charges, double static_contribution, double* out_vdw, double* out_coulomb) {
    constexpr int64_t LATENCY_MAX_POS = 256;
    if (n_pos > LATENCY_MAX_POS || n_atoms > SMALL_MAX_ATOMS) return 2;
    const bool recip = ewald && ewald->n_k > 0;
    if (rest_reim && !recip) return 2;
    SmallStage& stg = small_stage();
    const size_t n_idx8 = ((size_t)n_atoms + 1) / 2;
    const size_t n_rest = rest_reim ? (size_t)2 * ewald->n_k : 0;
    const size_t need = sizeof(double) * (8 + n_idx8 + (size_t)n_atoms + (size_t)3 * n_atoms * n_pos + (size_t)4 * n_pos + 4 + n_rest);
    size_t smem = 0;
    if (recip) {
        smem = sizeof(double2) * (size_t)n_atoms * ((size_t)(ewald->kmax[0] + 1) + (2 * ewald->kmax[1] + 1) + (2 * ewald->kmax[2] + 1));
        if (smem > 40 * 1024) return 2;
    }
    if (need > SmallStage::BYTES || !stg.ready(device)) return 2;
    double* hb = static_cast<double*>(stg.host);
    double* db = static_cast<double*>(stg.host_dev);
    const size_t o_idx = 8, o_q = o_idx + n_idx8, o_xyz = o_q + n_atoms;
    const size_t o_out = (o_xyz + (size_t)3 * n_atoms * n_pos + 1) & ~(size_t)1;     // 16-byte aligned result slots
    const size_t o_rest = o_out + (size_t)4 * n_pos;                                  // 16-byte aligned as well
    SmallParams S;
    S.E = P;
    S.E.n_pos = n_pos;
    S.E.n_atoms = n_atoms;
    S.n_idx8 = (int)n_idx8;
    const size_t n_in = n_idx8 + (size_t)n_atoms + (size_t)3 * n_atoms * n_pos;
    S.n_inline = n_in <= (size_t)SMALL_INLINE_DOUBLES ? (int)n_in : 0;
    double* dst = S.n_inline ? S.inl : hb + o_idx;      // by value in the launch, or through the mapped block
    dst[n_idx8 - 1] = 0.0;                              // padding word of an odd atom count
    if (atomsidx) memcpy(dst, atomsidx, sizeof(int32_t) * n_atoms);
    else memset(dst, 0, sizeof(double) * n_idx8);       // no grids: species index 0 for every atom
    memcpy(dst + n_idx8, charges, sizeof(double) * n_atoms);
    memcpy(dst + n_idx8 + n_atoms, xyz, sizeof(double) * 3 * n_atoms * n_pos);
    S.in = db + o_idx;
    S.has_ewald = recip ? 1 : 0;
    S.rest = nullptr;
    if (recip) {
        memcpy(S.invmat, ewald->invmat, sizeof S.invmat);
        S.kx = ewald->kmax[0]; S.ky = ewald->kmax[1]; S.kz = ewald->kmax[2];
        S.n_k = ewald->n_k;
        S.kv = ewald->d_kv.as<KVecDev>();
        if (rest_reim) {
            memcpy(hb + o_rest, rest_reim, sizeof(double) * n_rest);
            S.rest = reinterpret_cast<const double2*>(db + o_rest);
        }
    } else {
        S.n_k = 0; S.kv = nullptr; S.kx = S.ky = S.kz = 0;
    }
    S.energy_net_charges = energy_net_charges;
    S.static_contribution = static_contribution;
    S.out = reinterpret_cast<double2*>(db + o_out);
    S.gen = ++stg.gen;
    energy_points_small_kernel<<<(unsigned)n_pos, SMALL_THREADS, smem, stg.stream>>>(S);
    count_launch();
    CEG_CK(cudaGetLastError());
    // every result slot is {value, generation}, written by one 16-byte store: a slot that shows this call's
    // generation holds this call's value
    const volatile unsigned long long* slot = reinterpret_cast<const volatile unsigned long long*>(hb + o_out);
    unsigned long long spins = 0;
    for (int64_t i = 0; i < n_pos; i++) {
        while (slot[4 * i + 1] != S.gen || slot[4 * i + 3] != S.gen) {
            if ((++spins & 0xfffff) == 0) {            // every ~1M polls: has the launch failed?
                cudaError_t q = cudaStreamQuery(stg.stream);
                if (q != cudaErrorNotReady && q != cudaSuccess) return fail_cuda(q, "energy_points_small_kernel", __FILE__, __LINE__);
                if (q == cudaSuccess && (slot[4 * i + 1] != S.gen || slot[4 * i + 3] != S.gen))
                    return fail("latency kernel finished without completing");   // cannot happen; do not hang
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        if (out_vdw) out_vdw[i] = hb[o_out + 4 * i];
        out_coulomb[i] = hb[o_out + 4 * i + 2];
    }
    return 0;
}

// Median round trip of ONE empty launch on the latency path's machinery (same stream, same mapped flag, same polling
// loop): what any single-launch evaluation pays on this box before its kernel does any work.
extern "C" int ceg_measure_launch_roundtrip(int32_t device, int32_t reps, double* median_us) {
    if (!median_us || reps < 1) return fail("ceg_measure_launch_roundtrip: bad argument");
    if (use_device(device)) return 1;
    SmallStage& stg = small_stage();
    if (!stg.ready(device)) return fail("ceg_measure_launch_roundtrip: staging block unavailable");
    volatile unsigned long long* hflag = reinterpret_cast<volatile unsigned long long*>(stg.host);
    std::vector<double> t((size_t)reps);
    for (int r = -10; r < reps; r++) {
        const unsigned long long gen = ++stg.gen;
        const auto t0 = std::chrono::steady_clock::now();
        flag_only_kernel<<<1, 32, 0, stg.stream>>>(reinterpret_cast<volatile unsigned long long*>(stg.host_dev), gen);
        CEG_CK(cudaGetLastError());
        unsigned long long spins = 0;
        while (*hflag != gen)
            if ((++spins & 0xfffff) == 0 && cudaStreamQuery(stg.stream) != cudaErrorNotReady && *hflag != gen)
                return fail("ceg_measure_launch_roundtrip: kernel finished without raising the flag");
        if (r >= 0) t[r] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    CEG_CK(cudaStreamSynchronize(stg.stream));
    std::sort(t.begin(), t.end());
    *median_us = t[t.size() / 2];
    return 0;
}

static int P_device_of(const ceg_grid_h* vdw, int n_vdw, ceg_grid_h coulomb, ceg_ewald_h ewald, ceg_block_h block) {
    int device = -1;
    for (int g = 0; g < n_vdw; g++)
        if (vdw && vdw[g]) device = vdw[g]->device;
    if (coulomb) device = coulomb->device;
    if (ewald) device = ewald->device;
    if (block) device = block->device;
    return device;
}

// Shared front end of ceg_energy_points / ceg_energy_lattice: validates the handles and fills the grid part of
// the kernel parameters.  Returns -1 when nothing is resident (all grids are zero grids), 0 on success, 1 on error.
static int energy_points_setup(int32_t n_vdw, const ceg_grid_h* vdw, ceg_grid_h coulomb, ceg_ewald_h ewald,
                               ceg_block_h block, int32_t n_atoms, const int32_t* atomsidx, EnergyPointsParams& P) {
    if (n_vdw < 0 || n_vdw > EP_MAX_VDW) return fail("ceg_energy_points: too many distinct VdW grids");
    int device = -1;
    for (int g = 0; g < n_vdw; g++)
        if (vdw && vdw[g]) device = vdw[g]->device;
    if (coulomb) device = coulomb->device;
    if (ewald) device = ewald->device;
    if (block) device = block->device;
    if (device < 0) return -1;
    for (int a = 0; a < n_atoms; a++)
        if (atomsidx[a] < 0 || atomsidx[a] >= std::max(1, n_vdw)) return fail("ceg_energy_points: atomsidx out of range");
    if (use_device(device)) return 1;
    memset(&P, 0, sizeof P);
    P.n_vdw = n_vdw;
    for (int g = 0; g < n_vdw; g++) {
        P.vdw_present[g] = (vdw && vdw[g]) ? 1 : 0;
        if (P.vdw_present[g]) {
            if (vdw[g]->device != device) return fail("ceg_energy_points: handles live on different devices");
            P.vdw[g] = vdw[g]->g;
        }
    }
    P.has_coulomb = coulomb ? 1 : 0;
    if (coulomb) P.coulomb = coulomb->g;
    for (int g = 0; g < n_vdw; g++) {      // fused {VdW, Coulomb} records registered for this couple of handles?
        P.pair_data[g] = nullptr;
        if (P.vdw_present[g] && coulomb) {
            std::lock_guard<std::mutex> lk(vdw[g]->mu);
            if (vdw[g]->pair && vdw[g]->pair->coul == coulomb) P.pair_data[g] = vdw[g]->pair->g.data;
        }
    }
    P.has_block = block ? 1 : 0;
    if (block) P.block = block->b;
    P.n_atoms = n_atoms;
    return 0;
}

// Device part shared by both entry points: positions already on the device (d_xyz, n_pos x n_atoms x 3);
// d_v / d_c receive the VdW and Coulomb energies.  Everything is enqueued on `s`.
static int energy_points_enqueue(EnergyPointsParams P, ceg_ewald_h ewald, bool has_coulomb, int64_t n_pos, int32_t n_atoms,
                                 const double* d_xyz, double enc, double stat, double* d_v, double* d_c, uint8_t* d_blk,
                                 cudaStream_t s) {
    P.n_pos = n_pos;
    P.xyz = d_xyz;
    P.out_vdw = d_v;
    P.out_coulomb = d_c;
    P.blocked = d_blk;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n_pos + 127) / 128, sm_count_here() * 32));
    energy_points_kernel<<<blocks, 128, 0, s>>>(P);
    count_launch();
    CEG_CK(cudaGetLastError());
    if (has_coulomb && ewald && ewald->n_k > 0) {
        std::lock_guard<std::mutex> lk(ewald->mu);
        if (launch_reciprocal(ewald, n_pos, n_atoms, d_xyz, P.charges, enc, stat, d_c, d_blk, 1, s)) return 1;
    }
    return 0;
}

extern "C" int ceg_energy_points(int32_t n_vdw, const ceg_grid_h* vdw, ceg_grid_h coulomb, ceg_ewald_h ewald,
                                 ceg_block_h block, int64_t n_pos, int32_t n_atoms, const int32_t* atomsidx,
                                 const double* charges, const double* xyz, double energy_net_charges,
                                 double static_contribution, double* out_vdw, double* out_coulomb) {
    if (n_pos < 0 || n_atoms <= 0 || !atomsidx || !charges || (n_pos > 0 && (!xyz || !out_vdw || !out_coulomb)))
        return fail("ceg_energy_points: bad argument");
    if (n_pos == 0) return 0;
    EnergyPointsParams P;
    const int st = energy_points_setup(n_vdw, vdw, coulomb, ewald, block, n_atoms, atomsidx, P);
    if (st > 0) return 1;
    if (st < 0) {   // nothing resident: all grids are zero grids
        for (int64_t i = 0; i < n_pos; i++) { out_vdw[i] = 0.0; out_coulomb[i] = 0.0; }
        return 0;
    }
    // Latency path (single Monte-Carlo steps): ONE kernel launch per call (small_energy_points below the staging block).
    {
        const int rc = small_energy_points(P, P_device_of(vdw, n_vdw, coulomb, ewald, block), (coulomb && ewald) ? ewald : nullptr,
                                           nullptr, n_pos, n_atoms, atomsidx, charges, xyz, energy_net_charges,
                                           static_contribution, out_vdw, out_coulomb);
        if (rc != 2) return rc;
    }
    // Small batches (single Monte-Carlo steps): latency is what matters.  Inputs are packed into one pinned, thread-local
    // staging block and uploaded with ONE copy, the two outputs come back with ONE copy, and the stream is kept per
    // thread — 5 pageable copies and a stream create/destroy per call cost more than the kernels.
    const size_t n_idx_d = ((size_t)n_atoms + 1) / 2;                       // int32 indices, in units of doubles
    const size_t n_in_d = n_idx_d + (size_t)n_atoms + (size_t)3 * n_atoms * n_pos;
    const size_t n_out_d = (size_t)2 * n_pos;
    const bool small = (n_in_d + n_out_d) * sizeof(double) <= SmallStage::BYTES;
    SmallStage& stage = small_stage();
    cudaStream_t s = nullptr;
    bool own_stream = false;
    if (small && stage.ready(P_device_of(vdw, n_vdw, coulomb, ewald, block))) s = stage.stream;
    if (!s) {
        CEG_CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        own_stream = true;
    }
    DevBuf d_in, d_xyz, d_out, d_blk, d_idx, d_q;
    int rc = 0;
    do {
#define CK2(call) { cudaError_t _e = (call); if (_e != cudaSuccess) { rc = fail_cuda(_e, #call, __FILE__, __LINE__); break; } }
        CK2(d_out.reserve(sizeof(double) * n_out_d));
        CK2(d_blk.reserve(n_pos));
        double* d_v = d_out.as<double>();
        double* d_c = d_v + n_pos;
        const double* d_pos = nullptr;
        if (!own_stream) {
            double* hin = static_cast<double*>(stage.host);
            memcpy(hin, atomsidx, sizeof(int32_t) * n_atoms);
            memcpy(hin + n_idx_d, charges, sizeof(double) * n_atoms);
            memcpy(hin + n_idx_d + n_atoms, xyz, sizeof(double) * 3 * n_atoms * n_pos);
            CK2(d_in.reserve(sizeof(double) * n_in_d));
            CK2(cudaMemcpyAsync(d_in.p, hin, sizeof(double) * n_in_d, cudaMemcpyHostToDevice, s));
            P.atomsidx = reinterpret_cast<const int*>(d_in.as<double>());
            P.charges = d_in.as<double>() + n_idx_d;
            d_pos = d_in.as<double>() + n_idx_d + n_atoms;
        } else {
            CK2(d_idx.reserve(sizeof(int) * n_atoms));
            CK2(d_q.reserve(sizeof(double) * n_atoms));
            CK2(d_xyz.reserve(sizeof(double) * 3 * n_atoms * n_pos));
            CK2(cudaMemcpyAsync(d_idx.p, atomsidx, sizeof(int) * n_atoms, cudaMemcpyHostToDevice, s));
            CK2(cudaMemcpyAsync(d_q.p, charges, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, s));
            CK2(cudaMemcpyAsync(d_xyz.p, xyz, sizeof(double) * 3 * n_atoms * n_pos, cudaMemcpyHostToDevice, s));
            P.atomsidx = d_idx.as<int>();
            P.charges = d_q.as<double>();
            d_pos = d_xyz.as<double>();
        }
        if (energy_points_enqueue(P, ewald, coulomb != nullptr, n_pos, n_atoms, d_pos, energy_net_charges,
                                  static_contribution, d_v, d_c, d_blk.as<uint8_t>(), s)) { rc = 1; break; }
        if (!own_stream) {
            double* hout = static_cast<double*>(stage.host) + n_in_d;
            CK2(cudaMemcpyAsync(hout, d_v, sizeof(double) * n_out_d, cudaMemcpyDeviceToHost, s));
            CK2(cudaStreamSynchronize(s));
            memcpy(out_vdw, hout, sizeof(double) * n_pos);
            memcpy(out_coulomb, hout + n_pos, sizeof(double) * n_pos);
        } else {
            CK2(cudaMemcpyAsync(out_vdw, d_v, sizeof(double) * n_pos, cudaMemcpyDeviceToHost, s));
            CK2(cudaMemcpyAsync(out_coulomb, d_c, sizeof(double) * n_pos, cudaMemcpyDeviceToHost, s));
            CK2(cudaStreamSynchronize(s));
        }
    } while (0);
    if (rc) cudaStreamSynchronize(s);   // nothing may still use the scratch buffers when they go back to the pool
    if (own_stream) cudaStreamDestroy(s);
    return rc;
}

// energy_grid (grids.jl:346-424) with the lattice generated on the device: point (iA, iB, iC) of rotation r has its
// atoms at (iA stepA + iB stepB) + iC stepC + rotpos[r][a] (the reference's left-to-right sum, grids.jl:396, each
// product and sum rounded).  out[r][iA][iB][iC] = vdw + coulomb of energy_point (1e100 when blocked).
extern "C" int ceg_energy_lattice(int32_t n_vdw, const ceg_grid_h* vdw, ceg_grid_h coulomb, ceg_ewald_h ewald,
                                  ceg_block_h block, const int32_t* nums, const double* steps, int32_t n_rot,
                                  int32_t n_atoms, const double* rotpos, const int32_t* atomsidx, const double* charges,
                                  double energy_net_charges, double static_contribution, double* out) {
    if (!nums || !steps || n_rot <= 0 || n_atoms <= 0 || !rotpos || !atomsidx || !charges || !out)
        return fail("ceg_energy_lattice: bad argument");
    if (nums[0] <= 0 || nums[1] <= 0 || nums[2] <= 0) return fail("ceg_energy_lattice: empty lattice");
    const int64_t n_lat = (int64_t)nums[0] * nums[1] * nums[2];
    EnergyPointsParams P;
    const int st = energy_points_setup(n_vdw, vdw, coulomb, ewald, block, n_atoms, atomsidx, P);
    if (st > 0) return 1;
    if (st < 0) {
        for (int64_t i = 0; i < n_lat * n_rot; i++) out[i] = 0.0;
        return 0;
    }
    cudaStream_t s;
    CEG_CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    DevBuf d_idx, d_q, d_rot, d_xyz, d_v, d_c, d_blk;
    int rc = 0;
    do {
        CK2(d_idx.reserve(sizeof(int) * n_atoms));
        CK2(d_q.reserve(sizeof(double) * n_atoms));
        CK2(d_rot.reserve(sizeof(double) * 3 * n_atoms * n_rot));
        CK2(d_xyz.reserve(sizeof(double) * 3 * n_atoms * n_lat));
        CK2(d_v.reserve(sizeof(double) * n_lat));
        CK2(d_c.reserve(sizeof(double) * n_lat));
        CK2(d_blk.reserve(n_lat));
        CK2(cudaMemcpyAsync(d_idx.p, atomsidx, sizeof(int) * n_atoms, cudaMemcpyHostToDevice, s));
        CK2(cudaMemcpyAsync(d_q.p, charges, sizeof(double) * n_atoms, cudaMemcpyHostToDevice, s));
        CK2(cudaMemcpyAsync(d_rot.p, rotpos, sizeof(double) * 3 * n_atoms * n_rot, cudaMemcpyHostToDevice, s));
        P.atomsidx = d_idx.as<int>();
        P.charges = d_q.as<double>();
        LatticeParams L;
        for (int d = 0; d < 3; d++) L.nums[d] = nums[d];
        for (int d = 0; d < 9; d++) L.steps[d] = steps[d];
        L.n_atoms = n_atoms;
        L.n_lat = n_lat;
        const int blocks = (int)std::min<int64_t>((n_lat + 255) / 256, sm_count_here() * 32);
        for (int r = 0; r < n_rot && !rc; r++) {      // one rotation per pass bounds the scratch to one lattice
            L.rotpos = d_rot.as<double>() + (size_t)3 * n_atoms * r;
            lattice_positions_kernel<<<blocks, 256, 0, s>>>(L, d_xyz.as<double>());
            count_launch();
            CK2(cudaGetLastError());
            if (energy_points_enqueue(P, ewald, coulomb != nullptr, n_lat, n_atoms, d_xyz.as<double>(), energy_net_charges,
                                      static_contribution, d_v.as<double>(), d_c.as<double>(), d_blk.as<uint8_t>(), s)) { rc = 1; break; }
            add_energies_kernel<<<blocks, 256, 0, s>>>(n_lat, d_v.as<double>(), d_c.as<double>());
            count_launch();
            CK2(cudaGetLastError());
            CK2(cudaMemcpyAsync(out + (size_t)n_lat * r, d_v.p, sizeof(double) * n_lat, cudaMemcpyDeviceToHost, s));
        }
        if (!rc) CK2(cudaStreamSynchronize(s));
#undef CK2
    } while (0);
    if (rc) cudaStreamSynchronize(s);   // nothing may still use the scratch buffers when they go back to the pool
    cudaStreamDestroy(s);
    return rc;
}
