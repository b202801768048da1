// freesasa_b200/csrc/cells.cu — device-side uniform-grid cell list (counting sort by cell id).
//
// Replaces the serial, realloc-per-atom cell list of the reference (src/nb.c:43-72 bounds,
// :133-175 binning, :86-130 forward-cell table) with six small bandwidth-bound kernels whose
// launch geometry depends on the atom count only, so the host never waits for a device value.
// The result is DETERMINISTIC: inside a cell atoms are ordered by their caller index, so the sorted
// layout — and with it every floating-point summation order downstream — is identical from run to
// run and from GPU to GPU (the reference is bit-stable across thread counts,
// tests/test_freesasa.c:404-429).
#include "engine.cuh"

namespace fsb200 {

namespace {

constexpr int kBoundsChunk = 1024;  // atoms per block in k_bounds
constexpr int kScanChunk = 2048;    // ints per block in the scan (256 threads x 8)

__device__ __forceinline__ double warp_min(double v)
{
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- bounds ----------------------------------------------------------------------------------
// bounds[7*sid + {0,1,2}] = min x,y,z ; {3,4,5} = max x,y,z ; {6} = max R   (ordered encoding)
__global__ void __launch_bounds__(256) k_bounds(Workspace ws)
{
    const int b0 = blockIdx.x * kBoundsChunk;
    const int b1 = min(b0 + kBoundsChunk, ws.n);
    if (b0 >= b1) return;
    const int sid0 = ws.n_struct == 1 ? 0 : find_structure(ws.offsets, ws.n_struct, b0);
    const int sid1 = ws.n_struct == 1 ? 0 : find_structure(ws.offsets, ws.n_struct, b1 - 1);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf}, mr = 0.0;
    bool bad = false;

    for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
        const double x = ws.xyz[3 * i], y = ws.xyz[3 * i + 1], z = ws.xyz[3 * i + 2];
        const double R = ws.radii[i] + ws.probe;
        if (!(isfinite(x) && isfinite(y) && isfinite(z) && isfinite(R))) { bad = true; continue; }
        if (sid0 == sid1) {
            mn[0] = fmin(mn[0], x); mx[0] = fmax(mx[0], x);
            mn[1] = fmin(mn[1], y); mx[1] = fmax(mx[1], y);
            mn[2] = fmin(mn[2], z); mx[2] = fmax(mx[2], z);
            mr = fmax(mr, R);
        } else {  // chunk straddles a structure boundary (at most n_struct-1 blocks): per-atom atomics
            unsigned long long *b = ws.bounds + 7 * find_structure(ws.offsets, ws.n_struct, i);
            atomicMin(b + 0, encode_ordered(x)); atomicMax(b + 3, encode_ordered(x));
            atomicMin(b + 1, encode_ordered(y)); atomicMax(b + 4, encode_ordered(y));
            atomicMin(b + 2, encode_ordered(z)); atomicMax(b + 5, encode_ordered(z));
            atomicMax(b + 6, encode_ordered(R));
        }
    }
    if (bad) atomicExch(ws.counters + kCtrBadInput, 1);
    if (sid0 != sid1) return;

    __shared__ double red[7][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[7] = {warp_min(mn[0]), warp_min(mn[1]), warp_min(mn[2]),
                   warp_max(mx[0]), warp_max(mx[1]), warp_max(mx[2]), warp_max(mr)};
    if (lane == 0)
        for (int k = 0; k < 7; ++k) red[k][warp] = v[k];
    __syncthreads();
    if (threadIdx.x < 7) {
        const int k = threadIdx.x;
        double r = red[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = k < 3 ? fmin(r, red[k][w]) : fmax(r, red[k][w]);
        unsigned long long *b = ws.bounds + 7 * sid0 + k;
        if (k < 3) atomicMin(b, encode_ordered(r)); else atomicMax(b, encode_ordered(r));
    }
}

__global__ void k_init_bounds(unsigned long long *bounds, int n_struct)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 7 * n_struct) return;
    const int k = i % 7;
    bounds[i] = k < 3 ? ~0ull : 0ull;  // min slots start at +max, max slots at -max
}

// ---- grid ------------------------------------------------------------------------------------
// Cell edge = 2*max(R) as in the reference (src/nb.c:543), padded by 1e-9 so that two atoms closer
// than 2*max(R) along an axis can never be two cells apart after rounding.  Any edge >= 2*max(R)
// yields the same neighbour SET; if the bounding box would need more than 2*n+64 cells (sparse or
// elongated input) the edge grows until it fits.
__global__ void k_grid(Workspace ws)
{
    const int sid = blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= ws.n_struct) return;
    const int a0 = ws.n_struct == 1 ? 0 : ws.offsets[sid];
    const int a1 = ws.n_struct == 1 ? ws.n : ws.offsets[sid + 1];
    GridDesc g;
    g.atom_begin = a0;
    g.atom_end = a1;
    g.cell_base = cell_cap_base(a0, sid);
    const unsigned long long *b = ws.bounds + 7 * sid;
    double lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = decode_ordered(b[k]); hi[k] = decode_ordered(b[3 + k]); }
    const double rmax = decode_ordered(b[6]);
    if (a1 <= a0 || !(hi[0] >= lo[0])) {  // empty structure or only bad input
        for (int k = 0; k < 3; ++k) { g.lo[k] = 0; g.dim[k] = 1; }
        g.edge = 1.0;
        ws.grid[sid] = g;
        return;
    }
    double edge = 2.0 * rmax * (1.0 + 1e-9);
    if (!(edge > 0.0)) edge = 1.0;  // all radii zero: nobody can have a neighbour
    const double cap = (double)kCellsPerAtomCap * (a1 - a0) + kCellsSlack;
    // Every coordinate is finite (k_bounds checked), but the EXTENT hi - lo can still overflow to +inf (x = +-9e307):
    // then inf/edge = inf, inf/inf = NaN and `cells <= cap` would be false forever.  Such input is reported as bad input;
    // the loop is bounded in any case (1.26^3100 overflows a double, so a finite extent is always settled long before).
    bool settled = false;
    if (isfinite(hi[0] - lo[0]) && isfinite(hi[1] - lo[1]) && isfinite(hi[2] - lo[2])) {
        for (int iter = 0; iter < 4096 && !settled; ++iter) {
            double cells = 1.0;
            for (int k = 0; k < 3; ++k) {
                const double ext = floor((hi[k] - lo[k]) / edge) + 1.0;
                cells *= ext;
                g.dim[k] = ext < 2.0e9 ? (int)ext : 2000000000;
            }
            settled = cells <= cap;
            if (!settled) edge *= 1.26;
        }
    }
    if (!settled) {
        atomicExch(ws.counters + kCtrBadInput, 1);
        for (int k = 0; k < 3; ++k) { g.lo[k] = 0; g.dim[k] = 1; }
        g.edge = 1.0;
        ws.grid[sid] = g;
        return;
    }
    for (int k = 0; k < 3; ++k) g.lo[k] = lo[k];
    g.edge = edge;
    ws.grid[sid] = g;
}

__device__ __forceinline__ int local_cell(const GridDesc &g, double x, double y, double z)
{
    int c[3];
    cell_coords(g, x, y, z, c);
    return c[0] + g.dim[0] * (c[1] + g.dim[1] * c[2]);
}

// ---- count -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count(Workspace ws)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.n) return;
    const int sid = ws.n_struct == 1 ? 0 : find_structure(ws.offsets, ws.n_struct, i);
    const GridDesc &g = ws.grid[sid];
    double x = ws.xyz[3 * i], y = ws.xyz[3 * i + 1], z = ws.xyz[3 * i + 2];
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) { x = g.lo[0]; y = g.lo[1]; z = g.lo[2]; }  // flagged by k_bounds
    const int cell = g.cell_base + local_cell(g, x, y, z);
    ws.cell_of[i] = cell;
    atomicAdd(ws.cell_start + cell, 1);
}

// ---- exclusive scan (three small kernels) -------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int base = 0, sum = 0;
    for (int w = 0; w < 8; ++w) {
        if (w < warp) base += warp_sums[w];
        sum += warp_sums[w];
    }
    __syncthreads();
    *total = sum;
    return base + incl - v;
}

__global__ void __launch_bounds__(256) k_scan_sums(const int *data, int n, int *block_sums)
{
    const int base = blockIdx.x * kScanChunk + threadIdx.x * 8;
    int s = 0;
    for (int k = 0; k < 8; ++k)
        if (base + k < n) s += data[base + k];
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_top(int *block_sums, int n_blocks)
{
    int carry = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

// data[i] <- exclusive prefix; data[n] <- grand total (written by the last block)
__global__ void __launch_bounds__(256) k_scan_apply(int *data, int n, const int *block_sums)
{
    const int base = blockIdx.x * kScanChunk + threadIdx.x * 8;
    int v[8], s = 0;
    for (int k = 0; k < 8; ++k) {
        v[k] = base + k < n ? data[base + k] : 0;
        s += v[k];
    }
    int total;
    int run = block_sums[blockIdx.x] + block_exclusive_scan(s, &total);
    for (int k = 0; k < 8; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) data[n] = run;
}

// ---- scatter + deterministic reorder --------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(Workspace ws)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.n) return;
    const int cell = ws.cell_of[i];
    const int slot = atomicAdd(ws.cell_fill + cell, 1);
    ws.slot_atom[ws.cell_start[cell] + slot] = i;
}

// Each slot finds its atom's rank among the atoms of its cell (by caller index) and writes the
// packed record there; the first atom of every kItemAtoms-chunk publishes a work item.
__global__ void __launch_bounds__(256) k_reorder(Workspace ws)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ws.n) return;
    const int i = ws.slot_atom[p];
    const int cell = ws.cell_of[i];
    const int begin = ws.cell_start[cell], end = ws.cell_start[cell + 1];
    int rank = 0;
    for (int q = begin; q < end; ++q) rank += ws.slot_atom[q] < i;
    const int pos = begin + rank;
    const double R = ws.radii[i] + ws.probe;
    ws.atoms[pos] = make_double4(ws.xyz[3 * i], ws.xyz[3 * i + 1], ws.xyz[3 * i + 2], R);
    ws.perm[pos] = i;
    const int chunk_end = min(pos - rank % kItemAtoms + kItemAtoms, end);   // end of this atom's kItemAtoms-chunk
    if (rank % kItemAtoms == 0 && pos < ws.shard_end && chunk_end > ws.shard_begin) {   // other shards' items are never queued
        const int sid = ws.n_struct == 1 ? 0 : find_structure(ws.offsets, ws.n_struct, i);
        Item it;
        it.sid = sid;
        it.cell = cell - ws.grid[sid].cell_base;
        it.first = pos;
        it.count = min(kItemAtoms, end - pos);
        // Expensive work first: atoms of cells that touch empty space are the ones with exposed surface (they
        // cannot be settled by the buried-atom certificate and cost ~10x more), so their items go to the FRONT
        // of the queue and interior cells to the back; the kernel's tail then consists of cheap items.
        const GridDesc &g = ws.grid[sid];
        const int cx = it.cell % g.dim[0], cy = (it.cell / g.dim[0]) % g.dim[1], cz = it.cell / (g.dim[0] * g.dim[1]);
        bool surface = false;
        for (int dz = -1; dz <= 1 && !surface; ++dz)
            for (int dy = -1; dy <= 1 && !surface; ++dy)
                for (int dx = -1; dx <= 1 && !surface; ++dx) {
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    if (x < 0 || y < 0 || z < 0 || x >= g.dim[0] || y >= g.dim[1] || z >= g.dim[2]) { surface = true; break; }
                    const int c = g.cell_base + x + g.dim[0] * (y + g.dim[1] * z);
                    surface = ws.cell_start[c + 1] == ws.cell_start[c];
                }
        if (surface) ws.items[atomicAdd(ws.counters + kCtrItems, 1)] = it;
        else ws.items[ws.n - 1 - atomicAdd(ws.counters + kCtrItemsBack, 1)] = it;
    }
}

__global__ void k_unpermute(const int *perm, const double *sorted, double *out, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[perm[p]] = sorted[p];
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace

int launch_cell_build(const Workspace &ws, cudaStream_t stream)
{
    int launches = 0;
    const int n = ws.n, cells = ws.total_cells_cap;
    cudaMemsetAsync(ws.counters, 0, sizeof(int) * kCtrCount, stream);
    cudaMemsetAsync(ws.cell_start, 0, sizeof(int) * ((size_t)cells + 1), stream);
    cudaMemsetAsync(ws.cell_fill, 0, sizeof(int) * (size_t)cells, stream);
    k_init_bounds<<<div_up(7 * ws.n_struct, 256), 256, 0, stream>>>(ws.bounds, ws.n_struct); ++launches;
    k_bounds<<<div_up(n, kBoundsChunk), 256, 0, stream>>>(ws); ++launches;
    k_grid<<<div_up(ws.n_struct, 128), 128, 0, stream>>>(ws); ++launches;
    k_count<<<div_up(n, 256), 256, 0, stream>>>(ws); ++launches;
    const int scan_blocks = div_up(cells, kScanChunk);
    k_scan_sums<<<scan_blocks, 256, 0, stream>>>(ws.cell_start, cells, ws.scan_tmp); ++launches;
    k_scan_top<<<1, 256, 0, stream>>>(ws.scan_tmp, scan_blocks); ++launches;
    k_scan_apply<<<scan_blocks, 256, 0, stream>>>(ws.cell_start, cells, ws.scan_tmp); ++launches;
    k_scatter<<<div_up(n, 256), 256, 0, stream>>>(ws); ++launches;
    k_reorder<<<div_up(n, 256), 256, 0, stream>>>(ws); ++launches;
    return launches;
}

int launch_unpermute(const int *perm, const double *sorted, double *out, int n, cudaStream_t stream)
{
    k_unpermute<<<div_up(n, 256), 256, 0, stream>>>(perm, sorted, out, n);
    return 1;
}

}  // namespace fsb200
