// freesasa_b200/csrc/engine.cuh — shared declarations of the B200 SASA engine (sm_100a only).
//
// Pipeline of one call (all device-driven: no host round trip between the upload and the
// download; every buffer is sized from the atom count alone):
//
//   bounds   per-structure min/max/maxR            (replaces reference src/nb.c:43-72,242-254)
//   grid     per-structure cell edge + dimensions  (src/nb.c:55-71,543)
//   count    cell id per atom, histogram            (src/nb.c:133-175)
//   scan     exclusive scan of the histogram  ->  cell_start
//   scatter  atoms to their cell's slot range
//   reorder  deterministic order inside each cell, pack {x,y,z,R} as double4, emit work items
//   integrate  persistent kernel: per work item, TMA-stage the 27-cell neighbourhood into shared
//            memory, one warp per atom: exact fp64 neighbour test (src/nb.c:483-491), then
//            Lee-Richards slices (src/sasa_lr.c:270-408) or Shrake-Rupley test points
//            (src/sasa_sr.c:276-338) in the atom-local frame
//
// No global adjacency is ever materialised (the reference's nb_list is ~3.5 KB/atom).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fsb200 {

// Build-time experiment knobs.  The defaults ARE the product; tests/tools/ab_variants.py builds variants side by side
// (libfsb200_<name>.so) to measure one change at a time on the GPU box.
#ifndef FSB200_WARPS
#define FSB200_WARPS 14           // warps per CTA of k_integrate (16 at a 64-register cap: slower)
#endif
#ifndef FSB200_FUSED_VOTES
#define FSB200_FUSED_VOTES 1      // slice loop: one REDUX.OR instead of 1 + K votes (same speed within 1 %)
#endif
#ifndef FSB200_EXACT_SLICES
#define FSB200_EXACT_SLICES 1     // fp32 L&R: slices with a near-tangent circle pair are redone in fp64
#endif

#ifndef FSB200_NEAR_FLOOR
#define FSB200_NEAR_FLOOR 3.0e-6f // marginal band of the fp32 L&R path: q = min(|N|,|D|)/max(|N|,|D|) below this ...
#endif
#ifndef FSB200_NEAR_SCALE
#define FSB200_NEAR_SCALE 1.26e-6f // ... or below this x (slice thickness x Ri)^2 (thick slices need a wider band)
#endif
#ifndef FSB200_RING_ATOMICS
#define FSB200_RING_ATOMICS 1     // ring protocol through acquire/release atomics (0: round 1's volatile polls, for A/B timing)
#endif
#ifndef FSB200_SPLIT
#define FSB200_SPLIT 1            // fp32 L&R: k_integrate prepares, k_slices integrates chunks of slices (and the marginal ones in fp64), k_finish sums
#endif
#ifndef FSB200_CHUNK
#define FSB200_CHUNK 16           // slices per task of k_slices (C2: 8: 0.523 ms, 16: 0.517 ms; 4: slower)
#endif
#ifndef FSB200_SLICE_CTAS
#define FSB200_SLICE_CTAS 4       // CTAs of k_slices per SM the compiler has to make room for (register cap)
#endif
#ifndef FSB200_SLICES_K
#define FSB200_SLICES_K 3          // k_slices: instantiations of the slice loop by record groups: 1 = K 3 only; 2 = K 2 + 3; 3 = K 1 + 2 + 3
#endif
#ifndef FSB200_RING_SLOTS
#define FSB200_RING_SLOTS 2        // tiles in flight per CTA (3 x 512-atom tiles: no gain; 4 x 384: slower)
#endif
#ifndef FSB200_TILE_CAP
#define FSB200_TILE_CAP 640        // atoms per tile; larger neighbourhoods are read from global memory
#endif

constexpr int kWarpsPerCta = FSB200_WARPS;  // warps claim atoms dynamically from the CTA's ring of staged tiles
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kRingSlots = FSB200_RING_SLOTS; // tiles in flight per CTA
constexpr int kItemAtoms = 16;              // one work item = up to 16 consecutive atoms of one cell
constexpr int kTileCap = FSB200_TILE_CAP;    // atoms of the 27-cell neighbourhood staged in smem (32 B each)
constexpr int kNbCap = 160;                 // per-warp neighbour list capacity in smem
constexpr int kCertPoints = 128;           // probe directions of the buried-atom certificate: kCertPairs antipodal pairs
constexpr int kCertPairs = kCertPoints / 2;
constexpr int kCellsPerAtomCap = 2;         // grid budget: cells <= 2*n_k + 64 per structure
constexpr int kCellsSlack = 64;
constexpr int kMaxPeers = 8;                // other buffers (usually on other GPUs) one call can mirror its areas into

// Per-structure uniform grid, computed on the device.
struct GridDesc {
    double lo[3];
    double edge;      // >= 2*max(R) (slightly padded); grown if the box would need too many cells
    int dim[3];
    int cell_base;    // first global cell id of this structure (= 2*atom_begin + 64*sid)
    int atom_begin;
    int atom_end;
};

// One unit of work for the integration kernel: <= kItemAtoms consecutive atoms of one cell.
struct Item {
    int sid;      // structure
    int cell;     // cell id local to the structure
    int first;    // first sorted position
    int count;    // atoms in this item
};

// Device-side counters / status of one call (one 64-int block, zeroed per call).
enum CounterSlot {
    kCtrItems = 0,      // work items of surface cells (front of the item array)
    kCtrItemsBack = 14, // work items of interior cells (stored from the back of the item array)
    kCtrQueue = 1,      // work queue head of the integration kernel
    kCtrOverflow = 2,   // atoms whose neighbour list did not fit kNbCap
    kCtrMaxCand = 3,    // max candidate count among overflow atoms
    kCtrBadInput = 4,   // non-finite coordinate or radius seen
    kCtrQueue2 = 5,     // queue head of the overflow kernel
    kCtrStalled = 6,    // a warp gave up waiting for a tile (internal error, reported to the host)
    // 7..15: post-mortem of a stall; kCtrCertified shares slot 15 (only meaningful when nothing stalled)
    kCtrCertified = 15, // atoms proved completely buried by the coverage certificate (area 0 without integration)
    kCtrMarginal = 16,  // atoms with at least one marginal Lee-Richards slice redone in fp64
    kCtrCount = 20
};

// Plain data (no default member initialisers): api.cu zero-fills it and uses the bytes as the graph cache key.
struct Workspace {
    // inputs (device)
    const double *xyz;    // 3n AoS
    const double *radii;  // n, without probe
    int n;
    int n_struct;
    double probe;
    // per-structure
    int *offsets;                 // n_struct+1
    unsigned long long *bounds;   // 7 per structure, order-preserving encoding of doubles
    GridDesc *grid;
    // per-atom / per-cell
    int total_cells_cap;
    int *cell_of;      // n     global cell id of each atom (caller order)
    int *cell_start;   // total_cells_cap + 1  histogram -> exclusive scan
    int *cell_fill;    // total_cells_cap
    int *slot_atom;    // n     atom index by slot (unordered inside a cell)
    double4 *atoms;    // n     sorted {x,y,z,R=r+probe}
    int *perm;         // n     sorted position -> caller index
    Item *items;       // n
    int *scan_tmp;     // block sums for the scan
    int *counters;     // kCtrCount ints
    int *overflow;     // n     sorted positions of overflow atoms
    // output sharding of ONE replicated problem: only work items that touch sorted positions [shard_begin, shard_end)
    // are emitted by k_reorder, so a shard's queue holds its own items only (0, n = everything)
    int shard_begin;
    int shard_end;
};

// Control block of the split pipeline (zeroed per call, device memory)
struct TodoCtl {
    unsigned long long pool_head;   // bump allocator of the task-record pool (bytes)
    int n_todo;                     // task records written by k_integrate
    int task_head;                  // queue head of k_slices: task = (record, chunk of slices)
    int n_redo;                     // records with marginal slices (fp64 redo tasks, the first tasks of k_slices)
    int reserved;
    int n_inline;                   // atoms integrated inside k_integrate because the pool was full
    int pad;
};

struct IntegrateArgs {
    int alg;              // 0 LR, 1 SR
    int resolution;       // slices or points
    int precision;        // 0 fp32, 1 fp64
    int shard_begin;      // only items with first in [shard_begin, shard_end) are integrated
    int shard_end;
    int sorted_output;    // 1: out[sorted pos], 0: out[perm[pos]]
    double *out;          // n doubles
    int *nn_out;          // optional neighbour counts (caller order), or nullptr
    const float4 *points_f;   // SR: unit test points, float
    const double *points_d;   // SR: unit test points, 3*resolution doubles (bit-identical to the reference's)
    int grid_ctas;
    const float4 *cert_points;  // non-null: use the buried-atom certificate; kCertPairs unit vectors (the set is these and their negatives)
    // Fused collective: besides `out`, every area is also stored at the same index of these buffers, which may live on
    // OTHER GPUs (peer memory over NVLink: cudaDeviceEnablePeerAccess in one process, CUDA IPC between processes) — the
    // all-gather of the per-atom areas happens store by store from the integration epilogue, 8 B per atom and peer.
    int n_peer_out;
    double *peer_out[kMaxPeers];
    // owner_slice > 0: the result array is PARTITIONED by caller index over the buffers above — area of caller index i goes
    // to peer_out[i / owner_slice][i] and nowhere else (multi-GPU C entry point: every GPU ends up holding one contiguous
    // slice of the result, which it downloads over its own PCIe link)
    int owner_slice;
    int peer_skip_zero;   // the mirrors are zeroed by their owners before the call: an area of exactly 0 is not stored remotely
    // Split pipeline of the fp32 Lee-Richards path (null pool = fused: everything inside k_integrate): k_integrate only
    // gathers, certifies and PREPARES; atoms with exposed surface are written to the pool as task records and integrated by
    // k_slices (chunks of slices from a queue), marginal slices by k_redo.
    unsigned char *todo_pool;
    unsigned long long todo_cap;         // bytes
    unsigned long long *todo_list;       // [0, n): offsets of the task records; [todo_redo_base, +n): records with marginal slices
    int todo_redo_base;
    TodoCtl *todo_ctl;
};

// cells.cu
int launch_cell_build(const Workspace &ws, cudaStream_t stream);   // returns number of launches
// integrate.cu
int launch_integrate(const Workspace &ws, const IntegrateArgs &args, cudaStream_t stream);
size_t todo_pool_bytes_per_atom(int resolution);   // worst case of one task record
int launch_overflow(const Workspace &ws, const IntegrateArgs &args, int n_overflow, int list_cap,
                    void *scratch, cudaStream_t stream);
size_t overflow_scratch_bytes(int n_warps, int list_cap, int precision);
int overflow_warps(int n_overflow);
int integrate_grid_ctas(int alg, int precision, int device);
int launch_unpermute(const int *perm, const double *sorted, double *out, int n, cudaStream_t stream);

// ---- small device helpers --------------------------------------------------------------------
// order-preserving map double -> uint64 so min/max can use integer atomics
__host__ __device__ inline unsigned long long encode_ordered(double v)
{
    unsigned long long b;
#ifdef __CUDA_ARCH__
    b = (unsigned long long)__double_as_longlong(v);
#else
    union { double d; unsigned long long u; } c; c.d = v; b = c.u;
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double decode_ordered(unsigned long long k)
{
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    union { double d; unsigned long long u; } c; c.u = b; return c.d;
#endif
}

__host__ __device__ __forceinline__ int cell_cap_base(int atom_begin, int sid)
{
    return kCellsPerAtomCap * atom_begin + kCellsSlack * sid;
}

// structure owning caller-order atom i (offsets[sid] <= i < offsets[sid+1])
__device__ __forceinline__ int find_structure(const int *offsets, int n_struct, int i)
{
    int lo = 0, hi = n_struct;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offsets[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Cell coordinates of a point.  Division (not multiplication by a reciprocal) so that the cell
// of an atom is a monotone function of its coordinate with exactly the edge as period.
__device__ __forceinline__ void cell_coords(const GridDesc &g, double x, double y, double z, int c[3])
{
    int ix = (int)((x - g.lo[0]) / g.edge), iy = (int)((y - g.lo[1]) / g.edge), iz = (int)((z - g.lo[2]) / g.edge);
    c[0] = min(max(ix, 0), g.dim[0] - 1);
    c[1] = min(max(iy, 0), g.dim[1] - 1);
    c[2] = min(max(iz, 0), g.dim[2] - 1);
}

}  // namespace fsb200
