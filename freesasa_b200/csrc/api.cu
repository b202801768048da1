// freesasa_b200/csrc/api.cu — contexts, device scratch management and the C ABI (include/fsb200.h).
//
// The host side here is deliberately thin: validate, upload, enqueue the fixed kernel sequence
// (cells.cu, integrate.cu), read back one status block together with the results, and only in the
// rare large-neighbourhood case run a second pass.  There is no CPU implementation of the hot path
// in this library: if no sm_100 device is usable every compute entry point fails with a message.
#include "../../include/fsb200.h"
#include "engine.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

using namespace fsb200;

namespace {

// probe directions of the buried-atom certificate: one vector per antipodal pair
const float kCertDirs[kCertPairs][3] = {
#include "cert_dirs.inc"
};

thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return FSB200_FAIL;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = need + need / 4 + 64;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct fsb200_ctx {
    int device = 0;
    int precision = FSB200_FP32;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int *h_status = nullptr;  // pinned, kCtrCount ints
    // The cell-list build (3 memsets + 9 kernels) is captured once per distinct Workspace and replayed as ONE
    // CUDA graph launch: the device timeline no longer depends on how fast the host can issue 12 calls
    // (matters for small structures and on a busy host).
    cudaGraphExec_t graph_exec = nullptr;
    Workspace graph_ws, last_ws;
    int graph_launches = 0;
    unsigned char *h_stage = nullptr;  // pinned staging for host-pointer calls: inputs then outputs
    size_t h_stage_cap = 0;
    int grid_ctas[2][2] = {{0, 0}, {0, 0}};
    fsb200_stats stats{};
    std::mutex lock;

    // inputs / outputs when the caller hands us host memory
    DevBuf<double> in_xyz, in_radii, out_sasa;
    DevBuf<int> out_nn;
    // workspace
    DevBuf<int> offsets, cell_of, cell_start, cell_fill, slot_atom, perm, scan_tmp, counters, overflow;
    DevBuf<unsigned long long> bounds;
    DevBuf<GridDesc> grid;
    DevBuf<double4> atoms;
    DevBuf<Item> items;
    DevBuf<unsigned char> scratch;
    // Shrake-Rupley test points of the last resolution used
    // probe directions of the buried-atom certificate (uploaded once per context)
    DevBuf<float4> cert_points;
    bool use_certificate = true;
    int sr_points = 0;
    DevBuf<float4> points_f;
    DevBuf<double> points_d;
    int last_n = 0;  // atoms of the last device call (for unpermute)
};

namespace {

// Golden-spiral unit vectors, generated with the recurrence of the reference (src/sasa_sr.c:56-90:
// z and the longitude are ACCUMULATED) so that the fp64 re-check sees bit-identical points.  The points
// are then handed to the device PATCH-ORDERED: the number of exposed points does not depend on the order
// in which they are tested, and 32 consecutive points that form a compact patch of the sphere are usually
// hidden by the same one or two neighbours, which lets a warp (lane = point) leave its neighbour loop early.
// Ordering: latitude bands of roughly square patches (the spiral already runs from z = +1 to -1, so a band
// is an index range), sorted by longitude inside a band, alternating direction from band to band.
void make_test_points(int n, std::vector<double> &pd, std::vector<float4> &pf)
{
    std::vector<double> raw(3 * (size_t)n);
    const double dlong = M_PI * (3 - std::sqrt(5.0)), dz = 2.0 / n;
    double longitude = 0, z = 1 - dz / 2;
    for (int k = 0; k < n; ++k) {
        const double r = std::sqrt(1 - z * z);
        raw[3 * k] = std::cos(longitude) * r;
        raw[3 * k + 1] = std::sin(longitude) * r;
        raw[3 * k + 2] = z;
        z -= dz;
        longitude += dlong;
    }
    const int patches = (n + 31) / 32;
    const double side = std::sqrt(4.0 * M_PI / patches);           // angular size of a square patch
    const int bands = std::max(1, (int)std::lround(M_PI / side));
    std::vector<int> order(n);
    for (int k = 0; k < n; ++k) order[k] = k;
    for (int b = 0; b < bands; ++b) {
        const int lo = (int)((long long)n * b / bands), hi = (int)((long long)n * (b + 1) / bands);
        auto lon = [&](int k) { return std::atan2(raw[3 * k + 1], raw[3 * k]); };
        if (b % 2 == 0) std::sort(order.begin() + lo, order.begin() + hi, [&](int i, int j) { return lon(i) < lon(j); });
        else std::sort(order.begin() + lo, order.begin() + hi, [&](int i, int j) { return lon(i) > lon(j); });
    }
    pd.resize(3 * (size_t)n);
    pf.resize(n);
    for (int k = 0; k < n; ++k) {
        const int src = order[k];
        for (int a = 0; a < 3; ++a) pd[3 * k + a] = raw[3 * src + a];
        pf[k] = make_float4((float)pd[3 * k], (float)pd[3 * k + 1], (float)pd[3 * k + 2], 0.f);
    }
}

int ensure_workspace(fsb200_ctx *c, int n, int n_struct, Workspace &ws)
{
    const size_t cells = (size_t)kCellsPerAtomCap * n + (size_t)kCellsSlack * n_struct;
    if (cells + 1 > 0x7fffffffull) return fail("problem too large: %d atoms in %d structures", n, n_struct);
    CU(c->offsets.ensure((size_t)n_struct + 1));
    CU(c->bounds.ensure(7 * (size_t)n_struct));
    CU(c->grid.ensure(n_struct));
    CU(c->cell_of.ensure(n));
    CU(c->cell_start.ensure(cells + 1));
    CU(c->cell_fill.ensure(cells));
    CU(c->slot_atom.ensure(n));
    CU(c->atoms.ensure(n));
    CU(c->perm.ensure(n));
    CU(c->items.ensure(n));
    CU(c->scan_tmp.ensure(cells / 2048 + 2));
    CU(c->counters.ensure(kCtrCount));
    CU(c->overflow.ensure(n));
    ws.n = n;
    ws.n_struct = n_struct;
    ws.offsets = c->offsets.p;
    ws.bounds = c->bounds.p;
    ws.grid = c->grid.p;
    ws.total_cells_cap = (int)cells;
    ws.cell_of = c->cell_of.p;
    ws.cell_start = c->cell_start.p;
    ws.cell_fill = c->cell_fill.p;
    ws.slot_atom = c->slot_atom.p;
    ws.atoms = c->atoms.p;
    ws.perm = c->perm.p;
    ws.items = c->items.p;
    ws.scan_tmp = c->scan_tmp.p;
    ws.counters = c->counters.p;
    ws.overflow = c->overflow.p;
    return FSB200_SUCCESS;
}

int ensure_points(fsb200_ctx *c, int n_points, cudaStream_t stream)
{
    if (c->sr_points == n_points) return FSB200_SUCCESS;
    std::vector<double> pd;
    std::vector<float4> pf;
    make_test_points(n_points, pd, pf);
    CU(c->points_d.ensure(pd.size()));
    CU(c->points_f.ensure(pf.size()));
    CU(cudaMemcpyAsync(c->points_d.p, pd.data(), pd.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(c->points_f.p, pf.data(), pf.size() * sizeof(float4), cudaMemcpyHostToDevice, stream));
    CU(cudaStreamSynchronize(stream));  // the host vectors die here
    c->sr_points = n_points;
    return FSB200_SUCCESS;
}

// Host-pointer calls receive pageable memory (the caller's malloc'd arrays).  Copying it chunk-wise into a
// pinned staging buffer and sending each chunk with its own async DMA overlaps the CPU copy of chunk k+1
// with the transfer of chunk k and avoids the driver's internal pageable path (measured: ~2x less host
// overhead per call on the 100k-atom benchmark, 4x on the 1024-structure batch).
constexpr size_t kStageChunk = 1u << 20;
constexpr size_t kStageMax = 1ull << 30;   // beyond this fall back to plain pageable copies

// ---- host copy pool ------------------------------------------------------------------------------------
// Filling the pinned staging buffer is a plain memcpy from the caller's pageable arrays; at 3-200 MB per call
// one core's ~15 GB/s is the largest host-side cost of an end-to-end call.  Three persistent helper threads
// (created on first use, parked on a condition variable, never joined: the pool is deliberately leaked so that
// process exit does not wait on them) plus the calling thread copy 256 KB pieces in parallel.
constexpr size_t kParallelCopyBytes = 8u << 20;   // below this the single-threaded chunked path already hides the copy behind the DMA
constexpr size_t kCopyGroup = 4u << 20;            // bytes copied by the pool between two DMA submissions
constexpr size_t kCopyPiece = 256u << 10;

class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool *pool = new CopyPool(3);
        return *pool;
    }
    // fn(t) for t in [0, n_tasks), on the pool's threads and the caller's; returns when all are done
    void run(int n_tasks, const std::function<void(int)> &fn)
    {
        std::lock_guard<std::mutex> one_job_at_a_time(job_lock_);
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_tasks_ = n_tasks;
            next_.store(0);
            active_ = (int)workers_.size();
            ++generation_;
        }
        wake_.notify_all();
        for (int t; (t = next_.fetch_add(1)) < n_tasks;) fn(t);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return active_ == 0; });
        fn_ = nullptr;
    }

private:
    explicit CopyPool(int n)
    {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
        for (auto &w : workers_) w.detach();
    }
    void loop()
    {
        int seen = 0;
        for (;;) {
            const std::function<void(int)> *fn;
            int n;
            {
                std::unique_lock<std::mutex> lk(m_);
                wake_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                fn = fn_;
                n = n_tasks_;
            }
            for (int t; (t = next_.fetch_add(1)) < n;) (*fn)(t);
            std::lock_guard<std::mutex> g(m_);
            if (--active_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_lock_;
    std::condition_variable wake_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_tasks_ = 0, active_ = 0, generation_ = 0;
};

struct CopyJob {
    void *dst;
    const void *src;
    size_t bytes;
};

// all jobs, cut into kCopyPiece pieces, spread over the pool
void parallel_copy(const std::vector<CopyJob> &jobs)
{
    struct Piece { unsigned char *d; const unsigned char *s; size_t n; };
    std::vector<Piece> pieces;
    size_t total = 0;
    for (const CopyJob &j : jobs) {
        total += j.bytes;
        for (size_t o = 0; o < j.bytes; o += kCopyPiece)
            pieces.push_back({static_cast<unsigned char *>(j.dst) + o, static_cast<const unsigned char *>(j.src) + o,
                              j.bytes - o < kCopyPiece ? j.bytes - o : kCopyPiece});
    }
    if (total < kParallelCopyBytes) {
        for (const Piece &p : pieces) std::memcpy(p.d, p.s, p.n);
        return;
    }
    CopyPool::get().run((int)pieces.size(), [&](int t) { std::memcpy(pieces[t].d, pieces[t].s, pieces[t].n); });
}

int ensure_stage(fsb200_ctx *c, size_t bytes)
{
    if (bytes <= c->h_stage_cap) return FSB200_SUCCESS;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    c->h_stage_cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CU(cudaMallocHost((void **)&c->h_stage, want));
    c->h_stage_cap = want;
    return FSB200_SUCCESS;
}

// pageable src -> pinned stage (at stage_off) -> device dst, chunked
int staged_h2d(fsb200_ctx *c, void *dst_dev, const void *src, size_t bytes, size_t stage_off, cudaStream_t st)
{
    const unsigned char *s8 = static_cast<const unsigned char *>(src);
    for (size_t o = 0; o < bytes; o += kStageChunk) {
        const size_t m = bytes - o < kStageChunk ? bytes - o : kStageChunk;
        std::memcpy(c->h_stage + stage_off + o, s8 + o, m);
        CU(cudaMemcpyAsync(static_cast<unsigned char *>(dst_dev) + o, c->h_stage + stage_off + o, m, cudaMemcpyHostToDevice, st));
    }
    return FSB200_SUCCESS;
}

struct Request {
    int alg, resolution;
    double probe;
    int n, n_struct;
    const int *h_offsets;  // n_struct+1, or nullptr for one structure
    const double *d_xyz, *d_radii;
    double *d_out;
    int *d_nn;             // optional
    int shard_index, shard_count;
    cudaStream_t stream;
};

// Enqueue the whole pipeline.  `after_enqueue` (may be null) lets the host-buffer entry points queue
// their result download before the one synchronisation of the call.
template <typename F>
int run_pipeline(fsb200_ctx *c, const Request &rq, F after_enqueue)
{
    if (rq.alg != FSB200_LEE_RICHARDS && rq.alg != FSB200_SHRAKE_RUPLEY) return fail("unknown algorithm %d", rq.alg);
    if (rq.n <= 0) return fail("no atoms");
    if (rq.resolution <= 0) return fail("invalid resolution %d, must be > 0", rq.resolution);
    if (!std::isfinite(rq.probe) || rq.probe < 0) return fail("invalid probe radius %f", rq.probe);
    if (rq.shard_count < 1 || rq.shard_index < 0 || rq.shard_index >= rq.shard_count) return fail("invalid shard %d of %d", rq.shard_index, rq.shard_count);
    cudaStream_t st = rq.stream;
    Workspace ws;
    std::memset(&ws, 0, sizeof ws);  // the struct doubles as the graph cache key: no indeterminate padding
    ws.n_struct = 1;
    if (ensure_workspace(c, rq.n, rq.n_struct, ws)) return FSB200_FAIL;
    ws.xyz = rq.d_xyz;
    ws.radii = rq.d_radii;
    ws.probe = rq.probe;
    if (rq.n_struct > 1)
        CU(cudaMemcpyAsync(ws.offsets, rq.h_offsets, sizeof(int) * ((size_t)rq.n_struct + 1), cudaMemcpyHostToDevice, st));

    IntegrateArgs ia;
    std::memset(&ia, 0, sizeof ia);
    ia.alg = rq.alg;
    ia.resolution = rq.resolution;
    ia.precision = c->precision;
    ia.shard_begin = fsb200_shard_begin(rq.n, rq.shard_index, rq.shard_count);
    ia.shard_end = fsb200_shard_end(rq.n, rq.shard_index, rq.shard_count);
    ia.sorted_output = rq.shard_count > 1;
    ia.out = rq.d_out;
    ia.nn_out = rq.d_nn;
    if (rq.alg == FSB200_SHRAKE_RUPLEY) {
        if (ensure_points(c, rq.resolution, st)) return FSB200_FAIL;
        ia.points_f = c->points_f.p;
        ia.points_d = c->points_d.p;
    }
    ia.cert_points = c->use_certificate ? c->cert_points.p : nullptr;
    int &ctas = c->grid_ctas[rq.alg][c->precision];
    if (ctas == 0) ctas = integrate_grid_ctas(rq.alg, c->precision, c->device);
    ia.grid_ctas = ctas;

    // Cell-list build: 3 memsets + 9 small kernels whose arguments depend only on the workspace -> captured
    // once per distinct Workspace and replayed as ONE graph launch.  The integration kernel is launched
    // directly so that plain CUDA events can bracket it (events recorded inside a graph cannot be timed).
    int launches = 0;
    CU(cudaEventRecord(c->ev[0], st));
    bool replayed = false;
    if (c->graph_exec && std::memcmp(&c->graph_ws, &ws, sizeof ws) == 0) {
        replayed = cudaGraphLaunch(c->graph_exec, st) == cudaSuccess;
        launches = c->graph_launches;
    } else if (std::memcmp(&c->last_ws, &ws, sizeof ws) != 0) {
        // first sighting of this workspace: plain launches; capturing pays off only for repeated shapes
    } else if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
               (c->graph_exec ? (cudaGraphExecDestroy(c->graph_exec), c->graph_exec = nullptr, true) : true) &&
               cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        launches = launch_cell_build(ws, st);
        cudaGraph_t graph = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
        if (e_end == cudaSuccess && graph && cudaGraphInstantiate(&c->graph_exec, graph, 0) == cudaSuccess) {
            c->graph_ws = ws;
            c->graph_launches = launches;
            replayed = cudaGraphLaunch(c->graph_exec, st) == cudaSuccess;
        } else {
            c->graph_exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
    }
    if (!replayed) launches = launch_cell_build(ws, st);  // plain stream launches (new shape, legacy stream, or capture refused)
    c->last_ws = ws;
    CU(cudaEventRecord(c->ev[1], st));
    launches += launch_integrate(ws, ia, st);
    CU(cudaEventRecord(c->ev[2], st));
    CU(cudaMemcpyAsync(c->h_status, ws.counters, sizeof(int) * kCtrCount, cudaMemcpyDeviceToHost, st));
    if (after_enqueue(st)) return FSB200_FAIL;
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    c->last_n = rq.n;

    fsb200_stats &s = c->stats;
    s.n_atoms = rq.n;
    s.n_structures = rq.n_struct;
    s.n_items = c->h_status[kCtrItems] + c->h_status[kCtrItemsBack];
    s.n_overflow = c->h_status[kCtrOverflow];
    s.n_certified = c->h_status[kCtrCertified];
    s.max_neighbours = 0;
    if (c->h_status[kCtrBadInput]) {
        g_launches += launches;
        return fail("non-finite coordinate or radius in input");
    }
    if (c->h_status[kCtrStalled]) {
        g_launches += launches;
        const int *d = c->h_status + 7;
        return fail("internal error: the integration kernel's tile ring stalled (results discarded) "
                    "[cur %d | slot0 claim %08x gathered %d n*2+dead %d | slot1 claim %08x gathered %d n*2+dead %d | bars %d | queue %d of %d]",
                    d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], c->h_status[kCtrItems]);
    }
    float overflow_ms = 0.f;
    if (cudaEventElapsedTime(&s.device_ms, c->ev[0], c->ev[2]) != cudaSuccess) s.device_ms = -1.f;
    if (cudaEventElapsedTime(&s.integrate_ms, c->ev[1], c->ev[2]) != cudaSuccess) s.integrate_ms = -1.f;
    cudaGetLastError();
    if (s.n_overflow > 0) {
        // Large neighbourhoods (more than kNbCap neighbours): second pass with the lists in global memory.
        const int cap = ((c->h_status[kCtrMaxCand] + 7) / 8) * 8;
        s.max_neighbours = c->h_status[kCtrMaxCand];
        const int warps = overflow_warps(s.n_overflow);
        CU(c->scratch.ensure(overflow_scratch_bytes(warps, cap, c->precision)));
        CU(cudaEventRecord(c->ev[0], st));  // ev[0]/ev[2] of the first pass were consumed above
        launches += launch_overflow(ws, ia, s.n_overflow, cap, c->scratch.p, st);
        CU(cudaEventRecord(c->ev[3], st));
        if (after_enqueue(st)) return FSB200_FAIL;
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        cudaEventElapsedTime(&overflow_ms, c->ev[0], c->ev[3]);
    }
    s.device_ms += overflow_ms;
    s.kernel_launches = launches;
    g_launches += launches;
    return FSB200_SUCCESS;
}

// ---- context pool for the context-free entry points ---------------------------------------------------
std::mutex g_pool_lock;
std::vector<fsb200_ctx *> g_pool;  // idle contexts (any device)

fsb200_ctx *pool_acquire()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        fail("no CUDA device available: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> g(g_pool_lock);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i]->device == dev) {
                fsb200_ctx *c = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                return c;
            }
    }
    return fsb200_ctx_create(dev);
}

void pool_release(fsb200_ctx *c)
{
    std::lock_guard<std::mutex> g(g_pool_lock);
    g_pool.push_back(c);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *fsb200_last_error(void) { return g_error; }
const char *fsb200_version(void) { return "fsb200 0.1 (sm_100a)"; }
unsigned long long fsb200_launch_count(void) { return g_launches.load(); }

int fsb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int fsb200_available(void)
{
    const int n = fsb200_device_count();
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) return 1;
    }
    return 0;
}

int fsb200_shard_begin(int n_total, int shard_index, int shard_count)
{
    return (int)(((long long)n_total * shard_index) / shard_count);
}
int fsb200_shard_end(int n_total, int shard_index, int shard_count)
{
    return (int)(((long long)n_total * (shard_index + 1)) / shard_count);
}

fsb200_ctx *fsb200_ctx_create(int device)
{
    int n = fsb200_device_count();
    if (device < 0 || device >= n) {
        fail("CUDA device %d not available (%d visible): the engine has no CPU path", device, n);
        return nullptr;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) {
        fail("device %d has compute capability %d.x; this library contains sm_100a code only", device, major);
        return nullptr;
    }
    DeviceGuard guard(device);
    fsb200_ctx *c = new fsb200_ctx();
    c->device = device;
    {   // FSB200_PRECISION=fp64: the drop-in entry points (which have no precision argument) use the all-fp64 kernels
        const char *env = getenv("FSB200_PRECISION");
        if (env && (strcmp(env, "fp64") == 0 || strcmp(env, "FP64") == 0 || strcmp(env, "double") == 0)) c->precision = FSB200_FP64;
    }
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < 4; ++k) ok = cudaEventCreate(&c->ev[k]) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&c->h_status, sizeof(int) * kCtrCount) == cudaSuccess;
    if (ok) {  // the certificate's probe set: kCertPairs antipodal pairs (cert_dirs.inc)
        std::vector<float4> pf(kCertPairs);
        for (int k = 0; k < kCertPairs; ++k) pf[k] = make_float4(kCertDirs[k][0], kCertDirs[k][1], kCertDirs[k][2], 0.f);
        ok = c->cert_points.ensure(pf.size()) == cudaSuccess &&
             cudaMemcpy(c->cert_points.p, pf.data(), pf.size() * sizeof(float4), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        fail("could not initialise context on device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        fsb200_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

void fsb200_ctx_destroy(fsb200_ctx *c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->in_xyz.release(); c->in_radii.release(); c->out_sasa.release(); c->out_nn.release();
    c->offsets.release(); c->cell_of.release(); c->cell_start.release(); c->cell_fill.release();
    c->slot_atom.release(); c->perm.release(); c->scan_tmp.release(); c->counters.release();
    c->overflow.release(); c->bounds.release(); c->grid.release(); c->atoms.release();
    c->items.release(); c->scratch.release(); c->points_f.release(); c->points_d.release(); c->cert_points.release();
    for (int k = 0; k < 4; ++k)
        if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int fsb200_ctx_set_precision(fsb200_ctx *c, int precision)
{
    if (!c) return fail("null context");
    if (precision != FSB200_FP32 && precision != FSB200_FP64) return fail("unknown precision %d", precision);
    c->precision = precision;
    return FSB200_SUCCESS;
}

int fsb200_ctx_set_certificate(fsb200_ctx *c, int on)
{
    if (!c) return fail("null context");
    c->use_certificate = on != 0;
    return FSB200_SUCCESS;
}

int fsb200_ctx_stats(const fsb200_ctx *c, fsb200_stats *out)
{
    if (!c || !out) return fail("null argument");
    *out = c->stats;
    return FSB200_SUCCESS;
}

int fsb200_ctx_calc_batch(fsb200_ctx *c, int alg, int n_struct, const int *n_atoms, const double *const *xyz,
                          const double *const *radii, double *const *sasa, double probe, int resolution)
{
    if (!c) return fail("null context");
    if (n_struct <= 0 || !n_atoms || !xyz || !radii || !sasa) return fail("invalid batch arguments");
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    std::vector<int> off((size_t)n_struct + 1, 0);
    for (int k = 0; k < n_struct; ++k) {
        if (n_atoms[k] <= 0 || !xyz[k] || !radii[k] || !sasa[k]) return fail("structure %d is empty or has a null array", k);
        if ((long long)off[k] + n_atoms[k] > 0x3fffffffll) return fail("batch too large");
        off[k + 1] = off[k] + n_atoms[k];
    }
    const int n = off[n_struct];
    const auto t_begin = std::chrono::steady_clock::now();
    CU(c->in_xyz.ensure(3 * (size_t)n));
    CU(c->in_radii.ensure(n));
    CU(c->out_sasa.ensure(n));
    cudaStream_t st = c->stream;
    const size_t in_bytes = 32 * (size_t)n, out_bytes = 8 * (size_t)n;
    const bool staged = in_bytes + out_bytes <= kStageMax;
    const bool threaded = staged && in_bytes >= kParallelCopyBytes;  // big enough to wake the copy pool
    if (threaded) {
        // The PCIe link (~13 GB/s measured on the pool's hosts) is slower than four cores copying, so the
        // upload is pipelined: the pool fills kCopyGroup bytes of the staging buffer, their DMA is submitted,
        // and the pool moves on while that DMA runs.
        if (ensure_stage(c, in_bytes + out_bytes)) return FSB200_FAIL;
        for (int region = 0; region < 2; ++region) {
            const size_t unit = region == 0 ? 24 : 8;      // bytes per atom in this region
            unsigned char *h_base = c->h_stage + (region == 0 ? 0 : 24 * (size_t)n);
            unsigned char *d_base = region == 0 ? reinterpret_cast<unsigned char *>(c->in_xyz.p) : reinterpret_cast<unsigned char *>(c->in_radii.p);
            std::vector<CopyJob> group;
            size_t group_begin = 0, cursor = 0;            // byte offsets inside the region
            auto flush = [&]() -> int {
                if (cursor == group_begin) return FSB200_SUCCESS;
                parallel_copy(group);
                CU(cudaMemcpyAsync(d_base + group_begin, h_base + group_begin, cursor - group_begin, cudaMemcpyHostToDevice, st));
                group.clear();
                group_begin = cursor;
                return FSB200_SUCCESS;
            };
            for (int k = 0; k < n_struct; ++k) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(region == 0 ? (const void *)xyz[k] : (const void *)radii[k]);
                size_t left = unit * (size_t)n_atoms[k];
                while (left > 0) {
                    const size_t room = kCopyGroup - (cursor - group_begin);
                    const size_t m = left < room ? left : room;
                    group.push_back({h_base + cursor, src, m});
                    cursor += m;
                    src += m;
                    left -= m;
                    if (cursor - group_begin >= kCopyGroup && flush()) return FSB200_FAIL;
                }
            }
            if (flush()) return FSB200_FAIL;
        }
    } else if (staged) {
        if (ensure_stage(c, in_bytes + out_bytes)) return FSB200_FAIL;
        for (int k = 0; k < n_struct; ++k) {
            if (staged_h2d(c, c->in_xyz.p + 3 * (size_t)off[k], xyz[k], sizeof(double) * 3 * (size_t)n_atoms[k], 24 * (size_t)off[k], st)) return FSB200_FAIL;
            if (staged_h2d(c, c->in_radii.p + off[k], radii[k], sizeof(double) * (size_t)n_atoms[k], 24 * (size_t)n + 8 * (size_t)off[k], st)) return FSB200_FAIL;
        }
    } else {
        for (int k = 0; k < n_struct; ++k) {
            CU(cudaMemcpyAsync(c->in_xyz.p + 3 * (size_t)off[k], xyz[k], sizeof(double) * 3 * (size_t)n_atoms[k], cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(c->in_radii.p + off[k], radii[k], sizeof(double) * (size_t)n_atoms[k], cudaMemcpyHostToDevice, st));
        }
    }
    const auto t_staged = std::chrono::steady_clock::now();
    Request rq{alg, resolution, probe, n, n_struct, off.data(), c->in_xyz.p, c->in_radii.p, c->out_sasa.p, nullptr, 0, 1, st};
    double *h_out = staged ? reinterpret_cast<double *>(c->h_stage + in_bytes) : nullptr;
    auto download = [&](cudaStream_t s) -> int {
        if (staged) {
            CU(cudaMemcpyAsync(h_out, c->out_sasa.p, out_bytes, cudaMemcpyDeviceToHost, s));
        } else {
            for (int k = 0; k < n_struct; ++k)
                CU(cudaMemcpyAsync(sasa[k], c->out_sasa.p + off[k], sizeof(double) * (size_t)n_atoms[k], cudaMemcpyDeviceToHost, s));
        }
        return FSB200_SUCCESS;
    };
    const int rc = run_pipeline(c, rq, download);
    if (rc == FSB200_SUCCESS && threaded) {
        std::vector<CopyJob> jobs;
        jobs.reserve(n_struct);
        for (int k = 0; k < n_struct; ++k) jobs.push_back({sasa[k], h_out + off[k], sizeof(double) * (size_t)n_atoms[k]});
        parallel_copy(jobs);
    } else if (rc == FSB200_SUCCESS && staged)
        for (int k = 0; k < n_struct; ++k) std::memcpy(sasa[k], h_out + off[k], sizeof(double) * (size_t)n_atoms[k]);
    const auto t_end = std::chrono::steady_clock::now();
    c->stats.host_stage_ms = std::chrono::duration<float, std::milli>(t_staged - t_begin).count();
    c->stats.host_total_ms = std::chrono::duration<float, std::milli>(t_end - t_begin).count();
    return rc;
}

int fsb200_ctx_calc(fsb200_ctx *c, int alg, double *sasa, const double *xyz, const double *radii, int n, double probe,
                    int resolution)
{
    if (!sasa || !xyz || !radii) return fail("null array");
    return fsb200_ctx_calc_batch(c, alg, 1, &n, &xyz, &radii, &sasa, probe, resolution);
}

int fsb200_ctx_neighbour_counts(fsb200_ctx *c, int *counts, const double *xyz, const double *radii, int n, double probe)
{
    if (!c || !counts || !xyz || !radii || n <= 0) return fail("invalid arguments");
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    CU(c->in_xyz.ensure(3 * (size_t)n));
    CU(c->in_radii.ensure(n));
    CU(c->out_sasa.ensure(n));
    CU(c->out_nn.ensure(n));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->in_xyz.p, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->in_radii.p, radii, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    Request rq{FSB200_LEE_RICHARDS, 1, probe, n, 1, nullptr, c->in_xyz.p, c->in_radii.p, c->out_sasa.p, c->out_nn.p, 0, 1, st};
    auto download = [&](cudaStream_t s) -> int {
        CU(cudaMemcpyAsync(counts, c->out_nn.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
        return FSB200_SUCCESS;
    };
    return run_pipeline(c, rq, download);
}

int fsb200_ctx_calc_device(fsb200_ctx *c, int alg, const double *d_xyz, const double *d_radii, int n_total, int n_struct,
                           const int *offsets, double probe, int resolution, int shard_index, int shard_count,
                           double *d_sasa, void *stream)
{
    if (!c || !d_xyz || !d_radii || !d_sasa) return fail("null argument");
    if (n_struct < 1 || (n_struct > 1 && !offsets)) return fail("offsets required for %d structures", n_struct);
    if (n_struct > 1 && shard_count > 1) return fail("sharding applies to a single replicated structure");
    if (n_struct > 1 && (offsets[0] != 0 || offsets[n_struct] != n_total)) return fail("offsets do not span the atoms");
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    Request rq{alg, resolution, probe, n_total, n_struct, offsets, d_xyz, d_radii, d_sasa, nullptr, shard_index, shard_count, st};
    return run_pipeline(c, rq, [](cudaStream_t) { return FSB200_SUCCESS; });
}

int fsb200_ctx_unpermute(fsb200_ctx *c, const double *d_sorted, double *d_out, int n_total, void *stream)
{
    if (!c || !d_sorted || !d_out) return fail("null argument");
    if (n_total != c->last_n) return fail("unpermute: %d atoms but the last call had %d", n_total, c->last_n);
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    g_launches += launch_unpermute(c->perm.p, d_sorted, d_out, n_total, st);
    CU(cudaStreamSynchronize(st));
    return FSB200_SUCCESS;
}

int fsb200_cert_directions(double *out)
{
    if (!out) return fail("fsb200_cert_directions: null output");
    for (int k = 0; k < kCertPairs; ++k)
        for (int a = 0; a < 3; ++a) {
            out[3 * k + a] = (double)kCertDirs[k][a];
            out[3 * (k + kCertPairs) + a] = -(double)kCertDirs[k][a];
        }
    return FSB200_SUCCESS;
}

int fsb200_test_points(int n_points, double *out)
{
    if (n_points <= 0 || !out) return fail("invalid arguments");
    std::vector<double> pd;
    std::vector<float4> pf;
    make_test_points(n_points, pd, pf);
    std::memcpy(out, pd.data(), pd.size() * sizeof(double));
    return FSB200_SUCCESS;
}

// ---- context-free drop-in entry points ------------------------------------------------------------------
// Large batches through the context-free entry point are cut into contiguous sub-batches that two pooled contexts
// (two streams, two sets of device scratch, two pinned staging buffers) work through from two host threads: while one
// sub-batch is being integrated, the next one is being staged and uploaded and the previous one downloaded, so the
// PCIe transfers (a third to a half of an end-to-end call on the 1024-structure configuration) hide behind the kernels.
// Structures are independent, results do not depend on the split.
constexpr long long kOverlapMinAtoms = 400000;   // below this one pass is cheaper than a second thread
constexpr long long kOverlapChunkAtoms = 320000; // sub-batch size: ~10 MB up, ~2.5 MB down, a few ms of kernel

int fsb200_calc_batch(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                      double *const *sasa, double probe, int resolution)
{
    long long total = 0;
    if (n_struct > 0 && n_atoms)
        for (int k = 0; k < n_struct; ++k) total += n_atoms[k] > 0 ? n_atoms[k] : 0;
    if (n_struct < 4 || total < kOverlapMinAtoms || !xyz || !radii || !sasa) {
        fsb200_ctx *c = pool_acquire();
        if (!c) return FSB200_FAIL;
        const int rc = fsb200_ctx_calc_batch(c, alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution);
        pool_release(c);
        return rc;
    }
    const long long chunk_atoms = kOverlapChunkAtoms;
    const int n_workers = 2;   // measured on the 1024-structure configuration: 2, 3 or 4 workers and sub-batches of 100k..640k
                               // atoms all land within 35.6..37.8 ms (one pass: 57 ms, kernels alone: 25.7 ms)
    std::vector<int> begin;   // first structure of each sub-batch, plus the end
    long long acc = 0;
    begin.push_back(0);
    for (int k = 0; k < n_struct; ++k) {
        acc += n_atoms[k] > 0 ? n_atoms[k] : 0;
        if (acc >= chunk_atoms && k + 1 < n_struct) {
            begin.push_back(k + 1);
            acc = 0;
        }
    }
    begin.push_back(n_struct);
    const int n_chunks = (int)begin.size() - 1;
    constexpr int kMaxWorkers = 4;
    fsb200_ctx *ctx[kMaxWorkers] = {nullptr, nullptr, nullptr, nullptr};
    for (int w = 0; w < n_workers; ++w) {
        ctx[w] = pool_acquire();
        if (!ctx[w]) {
            for (int v = 0; v < w; ++v) pool_release(ctx[v]);
            return FSB200_FAIL;
        }
    }
    std::atomic<int> next{0};
    int rc[kMaxWorkers] = {FSB200_SUCCESS, FSB200_SUCCESS, FSB200_SUCCESS, FSB200_SUCCESS};
    char err[kMaxWorkers][sizeof g_error + 48] = {"", "", "", ""};
    auto work = [&](int w) {
        for (int k; rc[w] == FSB200_SUCCESS && (k = next.fetch_add(1)) < n_chunks;) {
            const int b = begin[k], cnt = begin[k + 1] - b;
            rc[w] = fsb200_ctx_calc_batch(ctx[w], alg, cnt, n_atoms + b, xyz + b, radii + b, sasa + b, probe, resolution);
            if (rc[w] != FSB200_SUCCESS) {
                snprintf(err[w], sizeof err[w], "structures %d..%d: %s", b, b + cnt - 1, g_error);   // g_error is thread-local
                next.store(n_chunks);                                                                 // stop the other worker too
            }
        }
    };
    std::vector<std::thread> helpers;
    for (int w = 1; w < n_workers; ++w) helpers.emplace_back(work, w);
    work(0);
    for (auto &h : helpers) h.join();
    for (int w = 0; w < n_workers; ++w) pool_release(ctx[w]);
    for (int w = 0; w < n_workers; ++w)
        if (rc[w] != FSB200_SUCCESS) return fail("%.500s", err[w]);
    return FSB200_SUCCESS;
}

int fsb200_lr(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_slices)
{
    if (!sasa || !xyz || !radii) return fail("null array");
    return fsb200_calc_batch(FSB200_LEE_RICHARDS, 1, &n, &xyz, &radii, &sasa, probe, n_slices);
}

int fsb200_sr(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_points)
{
    if (!sasa || !xyz || !radii) return fail("null array");
    return fsb200_calc_batch(FSB200_SHRAKE_RUPLEY, 1, &n, &xyz, &radii, &sasa, probe, n_points);
}

}  // extern "C"
