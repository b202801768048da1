// freesasa_b200/csrc/api.cu — contexts, device scratch management and the C ABI (include/fsb200.h).
//
// The host side here is deliberately thin: validate, upload, enqueue the fixed kernel sequence
// (cells.cu, integrate.cu), read back one status block together with the results, and only in the
// rare large-neighbourhood case run a second pass.  There is no CPU implementation of the hot path
// in this library: if no sm_100 device is usable every compute entry point fails with a message.
#include "../../include/fsb200.h"
#include "engine.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

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
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace fsb200;

namespace {

// probe directions of the buried-atom certificate: one vector per antipodal pair
const float kCertDirs[kCertPairs][3] = {
#include "cert_dirs.inc"
};

thread_local char g_error[512] = "";
thread_local fsb200_stats g_last_stats{};   // of the last context-free call on this thread (fsb200_last_stats)
std::atomic<unsigned long long> g_launches{0};

int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return FSB200_FAIL;
}

// NVTX range for one phase of a call (upload / cell build / integrate / download / multi-device)
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

// C++ exceptions must not cross the C ABI (std::bad_alloc from a vector, std::system_error from a thread that cannot be
// started): every extern "C" entry point that allocates runs its body through this.
template <typename F> int guarded(F body)
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return fail("out of host memory");
    } catch (const std::exception &e) {
        return fail("internal error: %s", e.what());
    } catch (...) {
        return fail("internal error: unknown exception");
    }
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
    cudaEvent_t ev_upload = nullptr;   // multi-device: "my slice of the inputs is on the device" (created on first use)
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
    // split pipeline of the fp32 L&R path: task-record pool, record lists, control block
    DevBuf<unsigned char> todo_pool;
    DevBuf<unsigned long long> todo_list;
    DevBuf<TodoCtl> todo_ctl;
    bool split_pipeline = true;
    size_t pool_bytes_per_atom = 1024;   // FSB200_POOL_BYTES_PER_ATOM: test hook (a tiny pool forces the in-kernel fallback)
    // Shrake-Rupley test points of the last resolution used
    // probe directions of the buried-atom certificate (uploaded once per context)
    DevBuf<float4> cert_points;
    bool use_certificate = true;
    int sr_points = 0;
    DevBuf<float4> points_f;
    DevBuf<double> points_d;
    int last_n = 0;  // atoms of the last device call (for unpermute)
    bool perm_from_device_call = false;   // the permutation in `perm` belongs to the last fsb200_ctx_calc_device[_async] call
    unsigned long long generation = 0;    // bumped by every pipeline run on this context (stamps `perm`)
    // a call that has been enqueued (fsb200_ctx_calc_device_async) but not finished (fsb200_ctx_finish)
    struct Pending {
        bool active = false;
        int alg = 0, n = 0, n_struct = 0, launches = 0;
        cudaStream_t stream = nullptr;
        Workspace ws;
        IntegrateArgs ia;
    } pending;
    // mirrors of the output on other GPUs (fsb200_ctx_set_peer_outputs), applied to device-resident calls
    int barrier_epoch = 0;
    DevBuf<int> barrier_status;   // one int, 0 = fine
    bool peer_skip_zero = false;   // peers zero their buffers themselves: areas that are exactly 0 are not stored remotely
    int n_peer_out = 0;
    double *peer_out[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
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

// A host thread that is created once and then parked on a condition variable between jobs.  The multi-device entry point
// and the overlapped batch path used to start fresh std::threads per call; a new thread pays for its CUDA thread state on
// its first runtime call (cudaSetDevice: 0.1-1 ms), which on eight GPUs was most of a 3 ms call.
class ParkedThread {
public:
    ~ParkedThread()
    {
        if (th_.joinable()) {
            {
                std::lock_guard<std::mutex> g(m_);
                quit_ = true;
            }
            cv_.notify_all();
            th_.join();
        }
    }
    void start(std::function<void()> job)   // may throw std::system_error (no thread could be created)
    {
        if (!th_.joinable()) th_ = std::thread([this] { loop(); });
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = std::move(job);
            pending_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return done_; });
    }

private:
    void loop()
    {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return pending_ || quit_; });
                if (quit_) return;
                job = std::move(job_);
                pending_ = false;
            }
            try {
                job();
            } catch (...) {
            }
            {
                std::lock_guard<std::mutex> g(m_);
                done_ = true;
            }
            cv_.notify_all();
        }
    }
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::function<void()> job_;
    bool pending_ = false, done_ = true, quit_ = false;
};

// worker d of the multi-device entry point (leaked on purpose: process exit must not wait for parked threads)
ParkedThread &device_worker(int d)
{
    static ParkedThread *workers[FSB200_MAX_DEVICES] = {};
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    if (!workers[d]) workers[d] = new ParkedThread();
    return *workers[d];
}
std::mutex g_multi_call_lock;   // one multi-device call at a time (it uses every device and every worker anyway)

struct CopyJob {
    void *dst;
    const void *src;
    size_t bytes;
};

// Worker threads of the multi-device entry point copy with their own core: the pool serialises its jobs (one at a time),
// which would queue the staging copies of all GPUs behind each other (measured on 8 GPUs: 7.5 ms per device instead of 4).
thread_local bool g_copy_inline = false;

// all jobs, cut into kCopyPiece pieces, spread over the pool
void parallel_copy(const std::vector<CopyJob> &jobs)
{
    if (g_copy_inline) {
        for (const CopyJob &j : jobs) std::memcpy(j.dst, j.src, j.bytes);
        return;
    }
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
    int sorted_output = 0;   // 1: d_out[sorted position]; 0: d_out[caller index]
    bool device_call = false;
    int n_peer_out = 0;
    double *const *peer_out = nullptr;
    int owner_slice = 0;     // > 0: results partitioned by caller index over peer_out[] (IntegrateArgs::owner_slice)
    bool peer_skip_zero = false;
};

// A call is two halves.  enqueue_pipeline() validates and puts the whole device sequence on the stream — cell-list build
// (one CUDA graph launch), the persistent integration kernel, the download of the status block — and returns without
// waiting; the caller may queue more work behind it (its result download, a collective, peer signalling).
// finish_pipeline() is the ONE synchronisation of the call: it waits, reads the status block, and only in the rare
// large-neighbourhood case runs the second pass (returning kSecondPass so that work queued in between can be redone).
constexpr int kSecondPass = 1;

int enqueue_pipeline(fsb200_ctx *c, const Request &rq)
{
    if (c->pending.active) return fail("a call is still pending on this context: call fsb200_ctx_finish() first");
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
    ws.shard_begin = fsb200_shard_begin(rq.n, rq.shard_index, rq.shard_count);
    ws.shard_end = fsb200_shard_end(rq.n, rq.shard_index, rq.shard_count);
    if (rq.n_struct > 1)
        CU(cudaMemcpyAsync(ws.offsets, rq.h_offsets, sizeof(int) * ((size_t)rq.n_struct + 1), cudaMemcpyHostToDevice, st));

    IntegrateArgs ia;
    std::memset(&ia, 0, sizeof ia);
    ia.alg = rq.alg;
    ia.resolution = rq.resolution;
    ia.precision = c->precision;
    ia.shard_begin = ws.shard_begin;
    ia.shard_end = ws.shard_end;
    ia.sorted_output = rq.sorted_output;
    ia.out = rq.d_out;
    ia.nn_out = rq.d_nn;
    ia.n_peer_out = rq.n_peer_out;
    for (int q = 0; q < rq.n_peer_out; ++q) ia.peer_out[q] = rq.peer_out[q];
    ia.owner_slice = rq.owner_slice;
    ia.peer_skip_zero = rq.peer_skip_zero ? 1 : 0;
    if (rq.alg == FSB200_SHRAKE_RUPLEY) {
        if (ensure_points(c, rq.resolution, st)) return FSB200_FAIL;
        ia.points_f = c->points_f.p;
        ia.points_d = c->points_d.p;
    }
    ia.cert_points = c->use_certificate ? c->cert_points.p : nullptr;
    if (FSB200_SPLIT && c->split_pipeline && rq.alg == FSB200_LEE_RICHARDS && c->precision == FSB200_FP32) {
        // Task records of the atoms that have to be integrated: 1 KB per owned atom covers ~50 % of the atoms at the
        // largest record (96 neighbours); if a call needs more, the remaining atoms are integrated inside k_integrate with
        // the same arithmetic in the same order (bit-identical), only slower.
        const size_t owned = (size_t)(ws.shard_end - ws.shard_begin);
        size_t bytes = owned * c->pool_bytes_per_atom + (c->pool_bytes_per_atom == 1024 ? (4u << 20) : 4096);
        if (bytes > (16ull << 30)) bytes = 16ull << 30;
        CU(c->todo_pool.ensure(bytes));
        CU(c->todo_list.ensure(2 * (size_t)rq.n + 2));
        CU(c->todo_ctl.ensure(1));
        ia.todo_pool = c->todo_pool.p;
        ia.todo_cap = bytes;
        ia.todo_list = c->todo_list.p;
        ia.todo_redo_base = rq.n + 1;
        ia.todo_ctl = c->todo_ctl.p;
    }
    int &ctas = c->grid_ctas[rq.alg][c->precision];
    if (ctas == 0) ctas = integrate_grid_ctas(rq.alg, c->precision, c->device);
    ia.grid_ctas = ctas;

    // Cell-list build: 3 memsets + 9 small kernels whose arguments depend only on the workspace -> captured
    // once per distinct Workspace and replayed as ONE graph launch.  The integration kernel is launched
    // directly so that plain CUDA events can bracket it (events recorded inside a graph cannot be timed).
    int launches = 0;
    CU(cudaEventRecord(c->ev[0], st));
    {
        Range r("fsb200:cell_build");
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
    }
    CU(cudaEventRecord(c->ev[1], st));
    {
        Range r("fsb200:integrate");
        launches += launch_integrate(ws, ia, st);
    }
    CU(cudaEventRecord(c->ev[2], st));
    CU(cudaMemcpyAsync(c->h_status, ws.counters, sizeof(int) * kCtrCount, cudaMemcpyDeviceToHost, st));
    CU(cudaGetLastError());
    fsb200_ctx::Pending &p = c->pending;
    p.active = true;
    p.alg = rq.alg;
    p.n = rq.n;
    p.n_struct = rq.n_struct;
    p.launches = launches;
    p.stream = st;
    p.ws = ws;
    p.ia = ia;
    ++c->generation;
    c->last_n = rq.n;
    c->perm_from_device_call = rq.device_call;
    return FSB200_SUCCESS;
}

// `after_second_pass(stream)` re-queues whatever the caller had queued behind the first pass (its download).
template <typename F>
int finish_pipeline(fsb200_ctx *c, F after_second_pass)
{
    fsb200_ctx::Pending &p = c->pending;
    if (!p.active) return fail("fsb200_ctx_finish: no call is pending on this context");
    p.active = false;
    cudaStream_t st = p.stream;
    int launches = p.launches;
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());

    fsb200_stats &s = c->stats;
    s.n_atoms = p.n;
    s.n_structures = p.n_struct;
    s.n_items = c->h_status[kCtrItems] + c->h_status[kCtrItemsBack];
    s.n_overflow = c->h_status[kCtrOverflow];
    s.n_certified = c->h_status[kCtrCertified];
    s.n_marginal = c->h_status[kCtrMarginal];
    s.max_neighbours = 0;
    if (c->h_status[kCtrBadInput]) {
        g_launches += launches;
        return fail("non-finite coordinate or radius in input (or a coordinate range that overflows a double)");
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
    int rc = FSB200_SUCCESS;
    if (s.n_overflow > 0) {
        // Large neighbourhoods (more than kNbCap neighbours): second pass with the lists in global memory.
        Range r("fsb200:overflow_pass");
        const int cap = ((c->h_status[kCtrMaxCand] + 7) / 8) * 8;
        s.max_neighbours = c->h_status[kCtrMaxCand];
        const int warps = overflow_warps(s.n_overflow);
        CU(c->scratch.ensure(overflow_scratch_bytes(warps, cap, c->precision)));
        CU(cudaEventRecord(c->ev[0], st));  // ev[0]/ev[2] of the first pass were consumed above
        launches += launch_overflow(p.ws, p.ia, s.n_overflow, cap, c->scratch.p, st);
        CU(cudaEventRecord(c->ev[3], st));
        if (after_second_pass(st)) return FSB200_FAIL;
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        cudaEventElapsedTime(&overflow_ms, c->ev[0], c->ev[3]);
        rc = kSecondPass;
    }
    s.device_ms += overflow_ms;
    s.kernel_launches = launches;
    g_launches += launches;
    return rc;
}

// Both halves back to back.  `after_enqueue` (may do nothing) lets the host-buffer entry points queue their result
// download before the one synchronisation of the call; it runs again after a second pass.
template <typename F>
int run_pipeline(fsb200_ctx *c, const Request &rq, F after_enqueue)
{
    if (enqueue_pipeline(c, rq)) return FSB200_FAIL;
    if (after_enqueue(rq.stream)) {
        c->pending.active = false;
        cudaStreamSynchronize(rq.stream);
        return FSB200_FAIL;
    }
    const int rc = finish_pipeline(c, after_enqueue);
    return rc == kSecondPass ? FSB200_SUCCESS : rc;
}

// ---- peer barrier over NVLink ---------------------------------------------------------------------------------
// flags[r] is rank r's flag array (world ints, peer-mapped).  A rank signals by storing the epoch into slot `rank` of EVERY
// rank's array (st.release.sys: everything this GPU wrote before — the peer stores of the integration kernel that ran
// earlier on the same stream — is visible to whoever acquires the flag) and then waits until every slot of ITS OWN array
// has reached the epoch (ld.acquire.sys).  One warp, lane = peer.  The wait is bounded: a dead peer turns into an error
// code in *status (rank + 1), never into a hung GPU.
struct PeerFlags {
    int *flags[kMaxPeers + 1];
};

__global__ void k_peer_barrier(PeerFlags pf, int rank, int world, int epoch, int *status)
{
    const int r = threadIdx.x;
    if (r < world) {
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[r] + rank), "r"(epoch) : "memory");
        const int *mine = pf.flags[rank] + r;
        bool ok = false;
        for (long long spin = 0; spin < (1ll << 24); ++spin) {    // >= 3 s with the back-off below
            int v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (v - epoch >= 0) {
                ok = true;
                break;
            }
            __nanosleep(200);
        }
        if (!ok) atomicCAS(status, 0, r + 1);
    }
}

// ---- context pool for the context-free entry points ---------------------------------------------------
// Idle contexts keep their device scratch (that is the point of the pool) but NOT unbounded host memory: a context whose
// pinned staging buffer grew beyond kPoolStageKeep gives it back when it returns to the pool, at most kPoolMaxIdle idle
// contexts are kept per device (surplus ones are destroyed), and fsb200_trim() releases everything that is idle.
std::mutex g_pool_lock;
std::vector<fsb200_ctx *> g_pool;  // idle contexts (any device)
constexpr size_t kPoolStageKeep = 256ull << 20;
constexpr int kPoolMaxIdle = 4;

// FREESASA_B200_DEVICE=<k>: the device the context-free entry points use when the caller has not chosen one with
// cudaSetDevice (i.e. the thread's current device is 0).
int default_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev == 0) {
        static const int env_dev = [] {
            const char *e = getenv("FREESASA_B200_DEVICE");
            return e && *e ? atoi(e) : 0;
        }();
        if (env_dev > 0) dev = env_dev;
    }
    return dev;
}

fsb200_ctx *pool_acquire(int dev = -1)
{
    if (dev < 0) dev = default_device();
    if (dev < 0) {
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
    if (c->h_stage_cap > kPoolStageKeep) {   // one huge call must not pin gigabytes of host memory for the life of the process
        cudaFreeHost(c->h_stage);
        c->h_stage = nullptr;
        c->h_stage_cap = 0;
    }
    c->n_peer_out = 0;
    {
        std::lock_guard<std::mutex> g(g_pool_lock);
        int same = 0;
        for (fsb200_ctx *o : g_pool) same += o->device == c->device;
        if (same < kPoolMaxIdle) {
            g_pool.push_back(c);
            return;
        }
    }
    fsb200_ctx_destroy(c);
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
    fsb200_ctx *c = new (std::nothrow) fsb200_ctx();
    if (!c) {
        fail("out of host memory");
        return nullptr;
    }
    c->device = device;
    {   // FSB200_PIPELINE=fused: everything inside the one persistent kernel (round 1's layout; for A/B measurements)
        const char *env = getenv("FSB200_PIPELINE");
        if (env && strcmp(env, "fused") == 0) c->split_pipeline = false;
        const char *pool = getenv("FSB200_POOL_BYTES_PER_ATOM");
        if (pool && atoi(pool) > 0) c->pool_bytes_per_atom = (size_t)atoi(pool);
    }
    {   // FSB200_PRECISION=fp64: the drop-in entry points (which have no precision argument) use the all-fp64 kernels
        const char *env = getenv("FSB200_PRECISION");
        if (env && (strcmp(env, "fp64") == 0 || strcmp(env, "FP64") == 0 || strcmp(env, "double") == 0)) c->precision = FSB200_FP64;
    }
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < 4; ++k) ok = cudaEventCreate(&c->ev[k]) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&c->h_status, sizeof(int) * kCtrCount) == cudaSuccess;
    if (ok) {  // the certificate's probe set: kCertPairs antipodal pairs (cert_dirs.inc)
        float4 pf[kCertPairs];
        for (int k = 0; k < kCertPairs; ++k) pf[k] = make_float4(kCertDirs[k][0], kCertDirs[k][1], kCertDirs[k][2], 0.f);
        ok = c->cert_points.ensure(kCertPairs) == cudaSuccess &&
             cudaMemcpy(c->cert_points.p, pf, sizeof pf, cudaMemcpyHostToDevice) == cudaSuccess;
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
    c->items.release(); c->scratch.release(); c->barrier_status.release();
    c->todo_pool.release(); c->todo_list.release(); c->todo_ctl.release(); c->points_f.release(); c->points_d.release(); c->cert_points.release();
    for (int k = 0; k < 4; ++k)
        if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
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

static int ctx_calc_batch_impl(fsb200_ctx *c, int alg, int n_struct, const int *n_atoms, const double *const *xyz,
                               const double *const *radii, double *const *sasa, double probe, int resolution);

int fsb200_ctx_calc_batch(fsb200_ctx *c, int alg, int n_struct, const int *n_atoms, const double *const *xyz,
                          const double *const *radii, double *const *sasa, double probe, int resolution)
{
    return guarded([&]() -> int { return ctx_calc_batch_impl(c, alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution); });
}

static int ctx_calc_batch_impl(fsb200_ctx *c, int alg, int n_struct, const int *n_atoms, const double *const *xyz,
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
    return guarded([&]() -> int {
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
    });
}

static int device_request(fsb200_ctx *c, Request &rq, int alg, const double *d_xyz, const double *d_radii, int n_total,
                          int n_struct, const int *offsets, double probe, int resolution, int shard_index, int shard_count,
                          double *d_sasa, void *stream)
{
    if (!c || !d_xyz || !d_radii || !d_sasa) return fail("null argument");
    if (n_struct < 1 || (n_struct > 1 && !offsets)) return fail("offsets required for %d structures", n_struct);
    if (n_struct > 1 && shard_count > 1) return fail("sharding applies to a single replicated structure");
    if (n_struct > 1 && (offsets[0] != 0 || offsets[n_struct] != n_total)) return fail("offsets do not span the atoms");
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    rq = Request{alg, resolution, probe, n_total, n_struct, offsets, d_xyz, d_radii, d_sasa, nullptr, shard_index, shard_count, st};
    rq.device_call = true;
    // with mirrors (fsb200_ctx_set_peer_outputs) every buffer receives the caller-order areas of this shard directly;
    // without, a shard writes its range of the SORTED order (gather the ranges, then fsb200_ctx_unpermute)
    rq.n_peer_out = c->n_peer_out;
    rq.peer_out = c->peer_out;
    rq.peer_skip_zero = c->peer_skip_zero;
    rq.sorted_output = (shard_count > 1 && c->n_peer_out == 0) ? 1 : 0;
    return FSB200_SUCCESS;
}

int fsb200_ctx_calc_device(fsb200_ctx *c, int alg, const double *d_xyz, const double *d_radii, int n_total, int n_struct,
                           const int *offsets, double probe, int resolution, int shard_index, int shard_count,
                           double *d_sasa, void *stream)
{
    return guarded([&]() -> int {
        Request rq{};
        if (device_request(c, rq, alg, d_xyz, d_radii, n_total, n_struct, offsets, probe, resolution, shard_index, shard_count, d_sasa, stream))
            return FSB200_FAIL;
        std::lock_guard<std::mutex> g(c->lock);
        DeviceGuard guard(c->device);
        return run_pipeline(c, rq, [](cudaStream_t) { return FSB200_SUCCESS; });
    });
}

int fsb200_ctx_calc_device_async(fsb200_ctx *c, int alg, const double *d_xyz, const double *d_radii, int n_total, int n_struct,
                                 const int *offsets, double probe, int resolution, int shard_index, int shard_count,
                                 double *d_sasa, void *stream)
{
    return guarded([&]() -> int {
        Request rq{};
        if (device_request(c, rq, alg, d_xyz, d_radii, n_total, n_struct, offsets, probe, resolution, shard_index, shard_count, d_sasa, stream))
            return FSB200_FAIL;
        std::lock_guard<std::mutex> g(c->lock);
        DeviceGuard guard(c->device);
        return enqueue_pipeline(c, rq);
    });
}

int fsb200_ctx_finish(fsb200_ctx *c)
{
    if (!c) return fail("null context");
    return guarded([&]() -> int {
        std::lock_guard<std::mutex> g(c->lock);
        DeviceGuard guard(c->device);
        return finish_pipeline(c, [](cudaStream_t) { return FSB200_SUCCESS; });
    });
}

int fsb200_ctx_set_peer_outputs(fsb200_ctx *c, int n_peers, double *const *d_peer_sasa)
{
    if (!c) return fail("null context");
    if (n_peers < 0 || n_peers > kMaxPeers) return fail("at most %d peer outputs", kMaxPeers);
    if (n_peers > 0 && !d_peer_sasa) return fail("null peer array");
    std::lock_guard<std::mutex> g(c->lock);
    if (c->pending.active) return fail("a call is still pending on this context");
    for (int q = 0; q < n_peers; ++q) {
        if (!d_peer_sasa[q]) return fail("peer output %d is null", q);
        c->peer_out[q] = d_peer_sasa[q];
    }
    c->n_peer_out = n_peers;
    return FSB200_SUCCESS;
}

int fsb200_ctx_set_peer_zero_skipping(fsb200_ctx *c, int on)
{
    if (!c) return fail("null context");
    std::lock_guard<std::mutex> g(c->lock);
    c->peer_skip_zero = on != 0;
    return FSB200_SUCCESS;
}

unsigned long long fsb200_ctx_generation(const fsb200_ctx *c) { return c ? c->generation : 0ull; }

int fsb200_ctx_unpermute(fsb200_ctx *c, const double *d_sorted, double *d_out, int n_total, void *stream)
{
    if (!c || !d_sorted || !d_out) return fail("null argument");
    std::lock_guard<std::mutex> g(c->lock);
    // `perm` is that of the LAST pipeline run on this context: refuse if that was not a device-resident call of this size
    // (a host-pointer call or fsb200_ctx_neighbour_counts in between silently replaces the permutation)
    if (c->pending.active) return fail("unpermute: a call is still pending on this context (fsb200_ctx_finish first)");
    if (!c->perm_from_device_call) return fail("unpermute: the last call on this context was not fsb200_ctx_calc_device");
    if (n_total != c->last_n) return fail("unpermute: %d atoms but the last call had %d", n_total, c->last_n);
    DeviceGuard guard(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    g_launches += launch_unpermute(c->perm.p, d_sorted, d_out, n_total, st);
    CU(cudaStreamSynchronize(st));
    return FSB200_SUCCESS;
}

// ---- CUDA IPC + peer barrier: the one-process-per-GPU form of the fused all-gather -------------------------------
int fsb200_ipc_alloc(int device, unsigned long long bytes, void **d_ptr, unsigned char handle[64])
{
    if (!d_ptr || !handle || bytes == 0) return fail("invalid arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    DeviceGuard guard(device);
    void *p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        cudaFree(p);
        return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    CU(cudaDeviceSynchronize());
    std::memcpy(handle, &h, 64);
    *d_ptr = p;
    return FSB200_SUCCESS;
}

int fsb200_ipc_open(int device, const unsigned char handle[64], void **d_ptr)
{
    if (!d_ptr || !handle) return fail("invalid arguments");
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FSB200_SUCCESS;
}

int fsb200_ipc_close(int device, void *d_ptr)
{
    DeviceGuard guard(device);
    CU(cudaIpcCloseMemHandle(d_ptr));
    return FSB200_SUCCESS;
}

int fsb200_ipc_free(int device, void *d_ptr)
{
    DeviceGuard guard(device);
    CU(cudaFree(d_ptr));
    return FSB200_SUCCESS;
}

int fsb200_ctx_peer_barrier(fsb200_ctx *c, int rank, int world, int *const *d_flags, void *stream)
{
    if (!c || !d_flags) return fail("null argument");
    if (world < 1 || world > kMaxPeers + 1 || rank < 0 || rank >= world) return fail("invalid rank %d of %d", rank, world);
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    PeerFlags pf;
    for (int r = 0; r < world; ++r) {
        if (!d_flags[r]) return fail("flag array of rank %d is null", r);
        pf.flags[r] = d_flags[r];
    }
    const int epoch = ++c->barrier_epoch;
    if (!c->barrier_status.p) {
        CU(c->barrier_status.ensure(1));
        CU(cudaMemsetAsync(c->barrier_status.p, 0, sizeof(int), st));
    }
    k_peer_barrier<<<1, 32, 0, st>>>(pf, rank, world, epoch, c->barrier_status.p);
    g_launches += 1;
    CU(cudaGetLastError());
    return FSB200_SUCCESS;
}

int fsb200_ctx_peer_barrier_status(fsb200_ctx *c)
{
    if (!c) return fail("null context");
    std::lock_guard<std::mutex> g(c->lock);
    DeviceGuard guard(c->device);
    int status = 0;
    CU(cudaMemcpy(&status, c->barrier_status.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (status != 0) return fail("peer barrier timed out waiting for rank %d (a peer died or never reached the barrier)", status - 1);
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
    return guarded([&]() -> int {
        std::vector<double> pd;
        std::vector<float4> pf;
        make_test_points(n_points, pd, pf);
        std::memcpy(out, pd.data(), pd.size() * sizeof(double));
        return FSB200_SUCCESS;
    });
}

// ---- context-free drop-in entry points ------------------------------------------------------------------
// Large batches through the context-free entry point are cut into contiguous sub-batches that two pooled contexts
// (two streams, two sets of device scratch, two pinned staging buffers) work through from two host threads: while one
// sub-batch is being integrated, the next one is being staged and uploaded and the previous one downloaded, so the
// PCIe transfers (a third to a half of an end-to-end call on the 1024-structure configuration) hide behind the kernels.
// Structures are independent, results do not depend on the split.
constexpr long long kOverlapMinAtoms = 400000;   // below this one pass is cheaper than a second thread
constexpr long long kOverlapChunkAtoms = 320000; // sub-batch size: ~10 MB up, ~2.5 MB down, a few ms of kernel

static int calc_batch_impl(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                           double *const *sasa, double probe, int resolution);

int fsb200_calc_batch(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                      double *const *sasa, double probe, int resolution)
{
    return guarded([&]() -> int { return calc_batch_impl(alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution); });
}

static int calc_batch_impl(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                           double *const *sasa, double probe, int resolution)
{
    long long total = 0;
    if (n_struct > 0 && n_atoms)
        for (int k = 0; k < n_struct; ++k) total += n_atoms[k] > 0 ? n_atoms[k] : 0;
    if (n_struct < 4 || total < kOverlapMinAtoms || !xyz || !radii || !sasa) {
        fsb200_ctx *c = pool_acquire();
        if (!c) return FSB200_FAIL;
        const int rc = fsb200_ctx_calc_batch(c, alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution);
        g_last_stats = c->stats;
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
    static_assert(kMaxWorkers >= 2, "");
    thread_local ParkedThread helper;   // n_workers == 2: this thread + one parked helper of its own
    const bool copy_inline = g_copy_inline;
    helper.start([&work, copy_inline] {
        g_copy_inline = copy_inline;
        work(1);
    });
    work(0);
    helper.wait();
    g_last_stats = ctx[0]->stats;   // the last sub-batch of worker 0; totals below
    g_last_stats.n_atoms = (int)total;
    g_last_stats.n_structures = n_struct;
    for (int w = 0; w < n_workers; ++w) pool_release(ctx[w]);
    for (int w = 0; w < n_workers; ++w)
        if (rc[w] != FSB200_SUCCESS) return fail("%.500s", err[w]);
    return FSB200_SUCCESS;
}

// ---- several GPUs behind one C call --------------------------------------------------------------------------
// One host thread per device, one pooled context each; no NCCL (one process): the exchange steps are peer copies and
// peer stores over NVLink.
//   * n_struct == 1 (one huge structure, config C5): the inputs are REPLICATED, the outputs PARTITIONED.  Device d
//     uploads only its 1/N slice of xyz / radii over its own PCIe link, then every device pulls the other slices from
//     its peers (an all-gather of the inputs over NVLink, N-1 peer copies per device); every device builds the identical
//     deterministic cell list, integrates only its contiguous share of the cell-sorted order, and stores each area
//     straight into the result slice of the device that OWNS that part of the caller's array (peer store over NVLink
//     from the integration epilogue: the exchange of the outputs costs no launch and no extra pass); every device then
//     downloads its contiguous slice over its own PCIe link.
//     Every atom sees ALL its neighbours — this is the whole-structure SASA, not the per-chain quantity of the
//     reference's --separate-chains (src/structure.c:955-1081).
//   * n_struct > 1 (independent structures, config C4): structures are dealt to the devices by longest-processing-time
//     on their atom counts; every device runs fsb200_calc_batch() on its share (uploads / downloads overlapped with the
//     kernels on two contexts), results land directly in the caller's per-structure arrays.
namespace {

struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n, count = 0, generation = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void drop()   // a participant that could not be started
    {
        std::unique_lock<std::mutex> lk(m);
        if (--n > 0 && count == n) {
            count = 0;
            ++generation;
            cv.notify_all();
        }
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const int g = generation;
        if (++count >= n) {
            count = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != g; });
        }
    }
};

fsb200_multi_stats g_multi_stats{};
std::mutex g_multi_stats_lock;

bool enable_peer_uncached(int from, int to)
{
    DeviceGuard guard(from);
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) {
        cudaGetLastError();
        return false;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    cudaGetLastError();
    return e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
}

bool enable_peer(int from, int to)
{
    static std::mutex m;
    static bool known[FSB200_MAX_DEVICES][FSB200_MAX_DEVICES], ok[FSB200_MAX_DEVICES][FSB200_MAX_DEVICES];
    std::lock_guard<std::mutex> g(m);
    if (from < FSB200_MAX_DEVICES && to < FSB200_MAX_DEVICES && known[from][to]) return ok[from][to];
    const bool result = enable_peer_uncached(from, to);
    if (from < FSB200_MAX_DEVICES && to < FSB200_MAX_DEVICES) {
        known[from][to] = true;
        ok[from][to] = result;
    }
    return result;
}

double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int multi_replicated(int alg, int n, const double *xyz, const double *radii, double *sasa, double probe, int resolution,
                     const std::vector<int> &devs)
{
    const int N = (int)devs.size();
    if (N > kMaxPeers) return fail("fsb200_calc_multi: one structure can be spread over at most %d devices", kMaxPeers);
    const int slice = (n + N - 1) / N;   // caller indices [o * slice, (o + 1) * slice) of the result live on device o
    for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b)
            if (a != b && !enable_peer(devs[a], devs[b]))
                return fail("fsb200_calc_multi: devices %d and %d cannot access each other's memory (no NVLink / PCIe P2P)", devs[a], devs[b]);
    std::vector<fsb200_ctx *> ctx(N, nullptr);
    for (int d = 0; d < N; ++d) {
        ctx[d] = pool_acquire(devs[d]);
        if (!ctx[d]) {
            for (int e = 0; e < d; ++e) pool_release(ctx[e]);
            return FSB200_FAIL;
        }
    }
    std::vector<cudaEvent_t> uploaded(N, nullptr);
    std::vector<int> rc(N, FSB200_SUCCESS);
    std::vector<std::string> err(N);
    std::atomic<bool> failed{false};
    HostBarrier barrier(N);
    fsb200_multi_stats ms{};
    ms.n_devices = N;
    ms.n_atoms = n;
    ms.n_structures = 1;
    const auto t_begin = std::chrono::steady_clock::now();

    auto work = [&](int d) {
        fsb200_ctx *c = ctx[d];
        std::lock_guard<std::mutex> lock(c->lock);
        g_copy_inline = true;
        cudaSetDevice(c->device);
        cudaStream_t st = c->stream;
        auto check = [&](int r) {
            if (r != FSB200_SUCCESS && rc[d] == FSB200_SUCCESS) {
                rc[d] = r;
                err[d] = g_error;
                failed.store(true);
            }
        };
        auto cu = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess) check(fail("%s failed on device %d: %s", what, c->device, cudaGetErrorString(e)));
        };
        const int a0 = fsb200_shard_begin(n, d, N), a1 = fsb200_shard_end(n, d, N), cnt = a1 - a0;
        {   // phase 1: my slice of the inputs, pageable -> my pinned staging -> my device, over my own PCIe link
            Range r("fsb200:multi:upload_slice");
            cu(c->in_xyz.ensure(3 * (size_t)n), "cudaMalloc");
            cu(c->in_radii.ensure(n), "cudaMalloc");
            cu(c->out_sasa.ensure(slice), "cudaMalloc");
            if (!c->ev_upload) cu(cudaEventCreateWithFlags(&c->ev_upload, cudaEventDisableTiming), "cudaEventCreate");
            uploaded[d] = c->ev_upload;
            if (rc[d] == FSB200_SUCCESS && ensure_stage(c, 32 * (size_t)(cnt > 0 ? cnt : 1) + 8 * (size_t)slice)) check(FSB200_FAIL);
            if (rc[d] == FSB200_SUCCESS && cnt > 0) {
                check(staged_h2d(c, c->in_xyz.p + 3 * (size_t)a0, xyz + 3 * (size_t)a0, 24 * (size_t)cnt, 0, st));
                check(staged_h2d(c, c->in_radii.p + a0, radii + a0, 8 * (size_t)cnt, 24 * (size_t)cnt, st));
            }
            if (uploaded[d]) cu(cudaEventRecord(uploaded[d], st), "cudaEventRecord");
        }
        barrier.wait();   // every slice is on its way, every buffer exists
        if (!failed.load()) {
            // phase 2: all-gather of the inputs over NVLink — pull every peer's slice once its upload has completed
            Range r("fsb200:multi:gather_inputs");
            for (int k = 1; k < N; ++k) {
                const int p = (d + k) % N;   // staggered, so that not everybody pulls from device 0 first
                const int b0 = fsb200_shard_begin(n, p, N), b1 = fsb200_shard_end(n, p, N);
                if (b1 <= b0) continue;
                cu(cudaStreamWaitEvent(st, uploaded[p], 0), "cudaStreamWaitEvent");
                cu(cudaMemcpyPeerAsync(c->in_xyz.p + 3 * (size_t)b0, c->device, ctx[p]->in_xyz.p + 3 * (size_t)b0, ctx[p]->device,
                                       24 * (size_t)(b1 - b0), st), "cudaMemcpyPeerAsync");
                cu(cudaMemcpyPeerAsync(c->in_radii.p + b0, c->device, ctx[p]->in_radii.p + b0, ctx[p]->device, 8 * (size_t)(b1 - b0), st),
                   "cudaMemcpyPeerAsync");
            }
        }
        if (d == 0) ms.upload_ms = (float)ms_since(t_begin);
        if (!failed.load()) {
            // phase 3: cell list (replicated) + my share of the atoms; every area goes straight to its owner (peer stores)
            Range r("fsb200:multi:integrate_shard");
            // the result is partitioned by caller index: owner o holds [o * slice, (o + 1) * slice); its buffer is addressed
            // through a base shifted by -o * slice so that the kernel can index every owner's buffer with the caller index
            double *owners[kMaxPeers];
            for (int o = 0; o < N; ++o) owners[o] = ctx[o]->out_sasa.p - (size_t)o * slice;
            Request rq{alg, resolution, probe, n, 1, nullptr, c->in_xyz.p, c->in_radii.p, c->out_sasa.p, nullptr, d, N, st};
            rq.sorted_output = 0;
            rq.n_peer_out = N;
            rq.peer_out = owners;
            rq.owner_slice = slice;
            check(run_pipeline(c, rq, [](cudaStream_t) { return FSB200_SUCCESS; }));
            if (rc[d] == FSB200_SUCCESS && d < FSB200_MAX_DEVICES) {
                ms.integrate_ms[d] = c->stats.integrate_ms;
                ms.device_ms[d] = c->stats.device_ms;
            }
        }
        barrier.wait();   // every shard has delivered its areas to their owners
        if (d == 0) ms.compute_ms = (float)ms_since(t_begin) - ms.upload_ms;
        const int o0 = d * slice, o1 = o0 + slice < n ? o0 + slice : n;
        if (!failed.load() && o1 > o0) {
            // phase 4: every device downloads ITS slice of the result over its own PCIe link, every thread copies its part out
            Range r("fsb200:multi:download_slice");
            double *h_out = reinterpret_cast<double *>(c->h_stage + 32 * (size_t)(cnt > 0 ? cnt : 1));
            cu(cudaMemcpyAsync(h_out, c->out_sasa.p, 8 * (size_t)(o1 - o0), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
            cu(cudaStreamSynchronize(st), "cudaStreamSynchronize");
            if (rc[d] == FSB200_SUCCESS) std::memcpy(sasa + o0, h_out, 8 * (size_t)(o1 - o0));
        }
    };
    {
        std::lock_guard<std::mutex> one_call(g_multi_call_lock);
        int started = 1;
        for (int d = 1; d < N; ++d) {
            try {
                device_worker(d).start([&work, d] { work(d); });
                ++started;
            } catch (...) {   // cannot start a worker thread: the call fails, but the started ones must not wait for it
                failed.store(true);
                rc[d] = FSB200_FAIL;
                err[d] = "cannot start a host thread";
                for (int e = d; e < N; ++e) barrier.drop();
                break;
            }
        }
        work(0);
        for (int d = 1; d < started; ++d) device_worker(d).wait();
    }
    ms.total_ms = (float)ms_since(t_begin);
    ms.download_ms = ms.total_ms - ms.upload_ms - ms.compute_ms;
    // (every stream has been synchronised by its own thread after the last peer copy out of any context: run_pipeline and
    //  the download both end with cudaStreamSynchronize, and all pulls are queued before a device's own kernels)
    for (int d = 0; d < N; ++d) pool_release(ctx[d]);
    {
        std::lock_guard<std::mutex> g(g_multi_stats_lock);
        g_multi_stats = ms;
    }
    for (int d = 0; d < N; ++d)
        if (rc[d] != FSB200_SUCCESS) return fail("device %d: %.480s", devs[d], err[d].c_str());
    return FSB200_SUCCESS;
}

int multi_batch(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                double *const *sasa, double probe, int resolution, const std::vector<int> &devs)
{
    const int N = (int)devs.size();
    // longest-processing-time first on the atom counts (deterministic)
    std::vector<int> order(n_struct);
    for (int k = 0; k < n_struct; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n_atoms[a] > n_atoms[b]; });
    std::vector<long long> load(N, 0);
    std::vector<std::vector<int>> mine(N);
    for (int k : order) {
        int best = 0;
        for (int d = 1; d < N; ++d)
            if (load[d] < load[best]) best = d;
        mine[best].push_back(k);
        load[best] += n_atoms[k] > 0 ? n_atoms[k] : 0;
    }
    std::vector<int> rc(N, FSB200_SUCCESS);
    std::vector<std::string> err(N);
    fsb200_multi_stats ms{};
    ms.n_devices = N;
    ms.n_structures = n_struct;
    for (int d = 0; d < N; ++d) ms.n_atoms += (int)load[d];
    const auto t_begin = std::chrono::steady_clock::now();
    auto work = [&](int d) {
        std::vector<int> &idx = mine[d];
        if (idx.empty()) return;
        std::sort(idx.begin(), idx.end());
        Range r("fsb200:multi:batch_share");
        g_copy_inline = true;
        cudaSetDevice(devs[d]);
        const size_t m = idx.size();
        std::vector<int> cnt(m);
        std::vector<const double *> px(m), pr(m);
        std::vector<double *> ps(m);
        for (size_t k = 0; k < m; ++k) {
            cnt[k] = n_atoms[idx[k]];
            px[k] = xyz[idx[k]];
            pr[k] = radii[idx[k]];
            ps[k] = sasa[idx[k]];
        }
        const auto t0 = std::chrono::steady_clock::now();
        rc[d] = fsb200_calc_batch(alg, (int)m, cnt.data(), px.data(), pr.data(), ps.data(), probe, resolution);
        if (rc[d] != FSB200_SUCCESS) err[d] = g_error;
        if (d < FSB200_MAX_DEVICES) ms.device_ms[d] = (float)ms_since(t0);   // wall time of this device's share
    };
    {
        std::lock_guard<std::mutex> one_call(g_multi_call_lock);
        int started = 1;
        bool all_started = true;
        for (int d = 1; d < N && all_started; ++d) {
            try {
                device_worker(d).start([&work, d] { work(d); });
                ++started;
            } catch (...) {   // cannot start a worker thread: the shares of the started ones complete, the call fails
                all_started = false;
                rc[d] = FSB200_FAIL;
                err[d] = "cannot start a host thread";
            }
        }
        const int prev = default_device();
        if (all_started) work(0);
        g_copy_inline = false;
        if (prev >= 0) cudaSetDevice(prev);
        for (int d = 1; d < started; ++d) device_worker(d).wait();
    }
    ms.total_ms = (float)ms_since(t_begin);
    ms.compute_ms = ms.total_ms;
    {
        std::lock_guard<std::mutex> g(g_multi_stats_lock);
        g_multi_stats = ms;
    }
    for (int d = 0; d < N; ++d)
        if (rc[d] != FSB200_SUCCESS) return fail("device %d: %.480s", devs[d], err[d].c_str());
    return FSB200_SUCCESS;
}

// the one-device case of fsb200_calc_multi: the ordinary batch call, timed like the others
int single_device(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                  double *const *sasa, double probe, int resolution)
{
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = fsb200_calc_batch(alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution);
    fsb200_multi_stats ms{};
    ms.n_devices = 1;
    ms.n_structures = n_struct;
    for (int k = 0; k < n_struct; ++k) ms.n_atoms += n_atoms[k] > 0 ? n_atoms[k] : 0;
    ms.total_ms = ms.compute_ms = (float)ms_since(t0);
    ms.device_ms[0] = g_last_stats.device_ms;
    ms.integrate_ms[0] = g_last_stats.integrate_ms;
    std::lock_guard<std::mutex> g(g_multi_stats_lock);
    g_multi_stats = ms;
    return rc;
}

}  // namespace

int fsb200_calc_multi(int alg, int n_struct, const int *n_atoms, const double *const *xyz, const double *const *radii,
                      double *const *sasa, double probe, int resolution, int n_devices)
{
    return guarded([&]() -> int {
        if (n_struct <= 0 || !n_atoms || !xyz || !radii || !sasa) return fail("invalid batch arguments");
        std::vector<int> devs;
        const int visible = fsb200_device_count();
        for (int d = 0; d < visible; ++d) {
            int major = 0;
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) devs.push_back(d);
        }
        if (devs.empty()) return fail("no sm_100 device visible: the engine has no CPU path");
        if (n_devices <= 0) n_devices = (int)devs.size();
        if (n_devices > (int)devs.size()) return fail("%d devices requested, %d usable", n_devices, (int)devs.size());
        if (n_devices > FSB200_MAX_DEVICES) return fail("at most %d devices", FSB200_MAX_DEVICES);
        devs.resize(n_devices);
        Range r("fsb200:calc_multi");
        if (n_struct == 1) {
            if (n_atoms[0] <= 0 || !xyz[0] || !radii[0] || !sasa[0]) return fail("structure 0 is empty or has a null array");
            // too small to be worth a second GPU: one device, the ordinary path
            if (n_devices == 1 || n_atoms[0] < 16384 * n_devices) {
                DeviceGuard guard(devs[0]);
                return single_device(alg, 1, n_atoms, xyz, radii, sasa, probe, resolution);
            }
            return multi_replicated(alg, n_atoms[0], xyz[0], radii[0], sasa[0], probe, resolution, devs);
        }
        if (n_devices == 1) {
            DeviceGuard guard(devs[0]);
            return single_device(alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution);
        }
        for (int k = 0; k < n_struct; ++k)
            if (n_atoms[k] <= 0 || !xyz[k] || !radii[k] || !sasa[k]) return fail("structure %d is empty or has a null array", k);
        return multi_batch(alg, n_struct, n_atoms, xyz, radii, sasa, probe, resolution, devs);
    });
}

int fsb200_lr_multi(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_slices, int n_devices)
{
    if (!sasa || !xyz || !radii) return fail("null array");
    return fsb200_calc_multi(FSB200_LEE_RICHARDS, 1, &n, &xyz, &radii, &sasa, probe, n_slices, n_devices);
}

int fsb200_sr_multi(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_points, int n_devices)
{
    if (!sasa || !xyz || !radii) return fail("null array");
    return fsb200_calc_multi(FSB200_SHRAKE_RUPLEY, 1, &n, &xyz, &radii, &sasa, probe, n_points, n_devices);
}

int fsb200_last_stats(fsb200_stats *out)
{
    if (!out) return fail("null argument");
    *out = g_last_stats;
    return FSB200_SUCCESS;
}

int fsb200_get_multi_stats(fsb200_multi_stats *out)
{
    if (!out) return fail("null argument");
    std::lock_guard<std::mutex> g(g_multi_stats_lock);
    *out = g_multi_stats;
    return FSB200_SUCCESS;
}

int fsb200_trim(void)
{
    std::vector<fsb200_ctx *> idle;
    {
        std::lock_guard<std::mutex> g(g_pool_lock);
        idle.swap(g_pool);
    }
    for (fsb200_ctx *c : idle) fsb200_ctx_destroy(c);
    return (int)idle.size();
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
