// freesasa_b200/csrc/integrate.cu — the two surface integrators for sm_100a.
//
// k_integrate<ALG,T> is a persistent kernel (details at the kernel): every CTA keeps a ring of two shared-memory tiles.  A
// tile holds the 27-cell neighbourhood of one work item (<= 16 atoms of one grid cell): 9 contiguous runs of the
// cell-sorted atom array (x is the fastest cell coordinate), staged with TMA bulk copies (cp.async.bulk + mbarrier, SASS
// UBLKCP).  The CTA's 14 warps are independent: each claims ONE atom at a time with a compare-and-swap, and the warp that
// completes the last neighbour gather of a tile recycles it from the global work queue.
//
// Per atom, one warp:
//   1. gather      filters the staged candidates with the reference's exact fp64 contact test
//                  dx*dx+dy*dy+dz*dz < (Ri+Rj)^2 (src/nb.c:483-491) and ballot-compacts the hits into its
//                  private shared-memory list in the atom-local frame (differences formed in fp64, then
//                  rounded).  The neighbour SET equals the reference's, without its ~10 % duplicate entries
//                  (forward-cell rule, src/nb.c:103-110) and without any global adjacency array.
//   2. certificate tries to PROVE that the atom has no exposed surface (certify_buried): 9 atoms in 10 of a large
//                  structure end here with area exactly 0, as in the reference.
//   3. Lee & Richards (src/sasa_lr.c:270-408), lanes = neighbours.  fp32 production path = SPLIT PIPELINE: this kernel
//      z-sorts and prepares the records (lr_prepare_sorted), marks the slices in which some pair of circles is within
//      rounding distance of a tangency (lr_mark_marginal, closed form) and writes a task record; k_slices runs the slice
//      loop (lr_slices_chunk: records in registers, one REDUX for burial / arcs, sector-mask full-cover early-out, exact
//      sort-free merge of what is left) on chunks of slices from a queue and the marginal slices in fp64 (redo_task);
//      k_finish adds the chunks in order.  Other paths, all inside k_integrate:
//        lr_atom_fastk<3>   the same slice loop, chunk after chunk (pool full, or FSB200_PIPELINE=fused)
//        lr_atom_fast       fp32, 97..kNbCap neighbours: all arcs of a slice merged pairwise
//        lr_atom<T>         generic (fp64 mode, the fp64 redo, the global-memory overflow kernel)
//      All use the cancellation-free half-angle form tan^2(alpha/2) = (a+b-d)(d+b-a)/((d+a-b)(a+b+d))
//      instead of acos of a quotient (src/sasa_lr.c:335) and hoist beta = atan2(dy,dx)+pi out of the slice
//      loop (the reference recomputes it per slice, :337).
//   4. Shrake & Rupley (src/sasa_sr.c:276-338), lanes = test points (sr_atom): point u of atom i is hidden by
//      neighbour a iff u.D_a >= t_a, D_a = x_a - x_i, t_a = (Ri^2+|D_a|^2-Ra^2)/(2Ri) — algebraically the
//      reference's |Ri u + x_i - x_a|^2 <= Ra^2.  3 FMA + compare per test in fp32; tests inside a rounding
//      band are re-decided by replaying the reference's exact fp64 expression, so every inside/outside
//      decision equals the reference's.
#include "engine.cuh"

namespace fsb200 {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kStallPolls = 1 << 26;   // motionless polls of the ring before a warp reports a stall (>= 4 s)

// ---- PTX wrappers: mbarrier + TMA bulk copy ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  Returns false instead of spinning forever (a stalled ring is
// reported to the host as an error, it must never hang the GPU).
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_addr(bar);
    for (int spin = 0; spin < (1 << 28); ++spin) {         // bounded by attempts, not by a clock (see the stall detector)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
// one non-blocking probe of the phase with the given parity
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return done != 0;
}
// global -> shared bulk copy (TMA, 1-D); bytes must be a multiple of 16, both addresses 16-B aligned
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// the same mask from the lane index with two ALU instructions: under register pressure the compiler re-reads the special
// register at every use (S2R, ~20 cycles; ncu round 2: 12 M of the slice kernel's 228 M warp instructions)
__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// ---- description of a staged fill (Slot::w below), one word per lane -------------------------------------
// payload word indices: runs r = 0..8 of the 27-cell neighbourhood, then the scalars
constexpr int kWBegin = 0;    // + r: first sorted position of run r
constexpr int kWCount = 9;    // + r: atoms in run r
constexpr int kWOff = 18;     // + r: tile index of run r's first atom
constexpr int kWAtoms = 27;   // atoms in the item (0 for a dead slot)
constexpr int kWFirst = 28;   // first sorted position of the item
constexpr int kWTotal = 29;   // candidates in the neighbourhood
constexpr int kWStaged = 30;  // 1: neighbourhood is in the slot's tile; 0: too large, read global memory
constexpr int kWSelf = 31;    // tile index of sorted position p is w[kWSelf] + p

// ---- per-warp record types ------------------------------------------------------------------------
template <typename T> struct alignas(4 * sizeof(T)) Rec4 { T a, b, c, d; };   // raw {dx,dy,dz,R}; LR {dz,R,dxy,beta}; SR {dx,dy,dz,t}
template <typename T> struct alignas(2 * sizeof(T)) Arc { T st, en; };

template <typename T> struct Consts;
template <> struct Consts<float> {
    static __device__ __forceinline__ float two_pi() { return 6.283185307179586f; }
    static __device__ __forceinline__ float pi() { return 3.141592653589793f; }
};
template <> struct Consts<double> {
    static __device__ __forceinline__ double two_pi() { return 6.283185307179586; }
    static __device__ __forceinline__ double pi() { return 3.141592653589793; }
};

// arcs one slice can produce: every neighbour (+4 sentinels) in the general paths, at most
// 96 + 32 synthetic + 4 sentinels in the <= 96-neighbour path
__host__ __device__ constexpr int arc_cap(int cap) { return cap + 4 > 132 ? cap + 4 : 132; }
constexpr int kCertList = 64;   // neighbours the buried-atom certificate works with (16 B each; fits the L&R arc array)

template <int ALG, typename T> struct WarpLayout {
    // bytes of shared (or scratch) memory one warp needs for a neighbour list of `cap` entries:
    //   L&R: records + arc_cap(cap) arcs of the current slice + their exact starts (the certificate's short list
    //        lives in the arcs array before the slices begin);  S&R: records + candidate indices + that list
    static __host__ __device__ constexpr size_t bytes(int cap)
    {
        return ALG == 0 ? (size_t)cap * sizeof(Rec4<T>) + (size_t)arc_cap(cap) * (sizeof(Arc<T>) + sizeof(T))
                        : (size_t)cap * (sizeof(Rec4<T>) + sizeof(int)) + kCertList * sizeof(float4);
    }
};

struct Self {
    double x, y, z, R;
};

// ---- step 1: neighbour gather ------------------------------------------------------------------
// Appends every candidate of base[0..cnt) that touches `self` to recs, in the atom-local frame
// (differences formed in fp64, then rounded to T):
//   ALG 0 (L&R): raw {dx, dy, dz, Rj}, turned into {dz, Rj, dxy, beta} by lr_prepare()
//   ALG 1 (S&R): {dx, dy, dz, t_j}, t_j = (Ri^2 + |D|^2 - Rj^2) / (2 Ri) evaluated in fp64,
//                plus the candidate's index (index_base + c) for the exact re-check
// skip = index (within this run) of the atom itself, or -1.  Returns the new neighbour count;
// entries beyond `cap` are counted but not stored.
template <int ALG, typename T>
__device__ __forceinline__ int gather_run(const double4 *base, int cnt, int skip, int index_base, const Self &s,
                                          Rec4<T> *recs, int *cidx, int nn, int cap, int lane)
{
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        const int c = c0 + lane;
        bool hit = false;
        double dx = 0, dy = 0, dz = 0, Rj = 0, d2 = 0;
        if (c < cnt && c != skip) {
            const double4 q = base[c];
            Rj = q.w;
            // the reference's comparison, term for term and without FMA contraction (src/nb.c:483-491)
            const double cut = __dadd_rn(s.R, Rj);
            const double cut2 = __dmul_rn(cut, cut);
            dx = __dsub_rn(q.x, s.x);
            dy = __dsub_rn(q.y, s.y);
            dz = __dsub_rn(q.z, s.z);
            d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            hit = d2 < cut2;
        }
        const unsigned m = __ballot_sync(kFull, hit);
        if (hit) {
            const int slot = nn + __popc(m & lanemask_lt(lane));
            if (slot < cap) {
                Rec4<T> r;
                r.a = (T)dx; r.b = (T)dy; r.c = (T)dz;
                if (ALG == 0) {
                    r.d = (T)Rj;
                } else {
                    r.d = s.R > 0.0 ? (T)((s.R * s.R + d2 - Rj * Rj) / (2.0 * s.R)) : (T)0;
                }
                if (ALG == 1 || FSB200_EXACT_SLICES) cidx[slot] = index_base + c;   // S&R: exact re-check; L&R: fp64 redo
                recs[slot] = r;
            }
        }
        nn += __popc(m);
    }
    return nn;
}

// ---- step 2: Lee & Richards ----------------------------------------------------------------------
// raw {dx,dy,dz,R} -> {dz, R, dxy, beta}
template <typename T>
__device__ __forceinline__ void lr_prepare(Rec4<T> *recs, int nn, int lane)
{
    for (int j = lane; j < nn; j += 32) {
        const Rec4<T> r = recs[j];
        Rec4<T> o;
        o.a = r.c;
        o.b = r.d;
        o.c = sqrt(r.a * r.a + r.b * r.b);                 // src/nb.c:440
        o.d = atan2(r.b, r.a) + Consts<T>::pi();           // src/sasa_lr.c:337, hoisted
        recs[j] = o;
    }
    __syncwarp();
}

// `only`: when non-zero in any lane, just the slices whose bit is set are evaluated (lane w holds the bits of slices
// 32 w .. 32 w + 31); used to redo marginal slices of the fp32 path in fp64
template <typename T>
__device__ __forceinline__ double lr_atom(const Rec4<T> *recs, Arc<T> *arcs, int nn, double Ri_d, int ns, int lane,
                                          bool masked = false, unsigned only = 0u)
{
    const T Ri = (T)Ri_d;
    const T two_pi = Consts<T>::two_pi();
    const double delta = 2.0 * Ri_d / ns;                  // src/sasa_lr.c:304
    const unsigned lt = lanemask_lt();
    double acc = 0.0;                                      // per-lane share of the exposed angle, all slices

    for (int s = 0; s < ns; ++s) {
        if (masked && !((__shfl_sync(kFull, only, (s >> 5) & 31) >> (s & 31)) & 1u)) continue;
        const T zr = (T)(-Ri_d + (s + 0.5) * delta);       // slice centre relative to the atom centre (:305-307)
        const T az = fabs(zr);
        const T a2 = (Ri - az) * (Ri + az);                // Ri'^2 (:309), factored to avoid cancellation
        if (!(a2 > (T)0)) continue;                        // :310-312
        const T a = sqrt(a2);
        int narc = 0;
        bool buried = false;
        for (int base = 0; base < nn; base += 32) {
            const int j = base + lane;
            bool has = false, bur = false;
            T st = 0, en = 0;
            if (j < nn) {
                const Rec4<T> r = recs[j];                 // {dz, R, dxy, beta}
                const T dj = fabs(r.a - zr);               // :317
                if (dj < r.b) {                            // :320
                    const T b = sqrt((r.b - dj) * (r.b + dj));  // Rj' (:321-322)
                    const T d = r.c;
                    const T f1 = (a + b) - d;              // > 0  <=> circles touch      (:324)
                    if (f1 > (T)0) {
                        const T f3 = (d + a) - b;          // < 0  <=> circle i inside j  (:327)
                        if (f3 < (T)0) {
                            bur = true;
                        } else {
                            const T f2 = (d + b) - a;      // < 0  <=> circle j inside i  (:331)
                            if (!(f2 < (T)0)) {
                                // alpha = acos((a^2+d^2-b^2)/(2ad)) (:335) in half-angle form
                                const T alpha = (T)2 * atan2(sqrt(f1 * f2), sqrt(f3 * ((a + b) + d)));
                                st = r.d - alpha;          // :338-341, kept as [st, st+2alpha) with st in [0,2pi)
                                if (st < (T)0) st += two_pi;
                                en = st + (T)2 * alpha;
                                has = true;
                            }
                        }
                    }
                }
            }
            if (__any_sync(kFull, bur)) { buried = true; break; }
            const unsigned m = __ballot_sync(kFull, has);
            if (has) {
                Arc<T> arc;
                arc.st = st; arc.en = en;
                arcs[narc + __popc(m & lt)] = arc;
            }
            narc += __popc(m);
        }
        if (buried) continue;                              // :359
        if (narc == 0) {                                   // :395
            if (lane == 0) acc += (double)two_pi;
            continue;
        }
        __syncwarp();
        // sort-free union of arcs on the circle (equivalent to :367-408)
        T max_en = 0;
        for (int k = lane; k < narc; k += 32) {
            const Arc<T> me = arcs[k];
            T P = 0, mx = 0;
            for (int m = 0; m < narc; ++m) {
                const Arc<T> o = arcs[m];                  // warp-uniform address: shared-memory broadcast
                mx = fmax(mx, o.en);
                if (o.st < me.st || (o.st == me.st && m < k)) P = fmax(P, o.en);
            }
            const T W = fmax(mx - two_pi, (T)0);           // part of [0, .) covered by arcs that wrap past 2pi
            acc += (double)fmax(me.st - fmax(W, P), (T)0);
            max_en = mx;
        }
        if (lane == 0) acc += (double)fmax(two_pi - max_en, (T)0);
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    return delta * Ri_d * acc;                             // :360: delta * R_i * exposed angle
}


// ---- step 2b: the fp32 production path of Lee & Richards -------------------------------------------
// Same mathematics as lr_atom<float>, with the two instruction hogs of that version (ncu, round 1:
// atan2f + IEEE sqrtf = 40 % of all issued instructions, the tie-breaking all-pairs merge 33 %)
// replaced:
//   * alpha = 2 atan(sqrt(q)) with q = min(N,D)/max(N,D) in [0,1], N = f1 f2, D = f3 (a+b+d), via
//     MUFU rcp/sqrt and a degree-7 minimax polynomial in q (|err| < 1.5e-7 rad), reflected for N > D;
//   * every arc gets a UNIQUE integer sort key = (bits(start) & ~0xff) | arc index (7 index bits in the
//     <= 96-neighbour path, which never holds more than 112 arcs), so "arc m comes before arc k" is one integer
//     compare; two starts within 256 (128) ulp may swap order, which changes the union only by that sliver
//     (<= 1e-4 rad) and only in the ~1e-3 of slices where it happens.
//     The exact start (kept in a side array) is still what the exposed gap is measured from.
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// atan(sqrt(q)) for q in [0,1]
__device__ __forceinline__ float atan_sqrt01(float q)
{
    float p = -4.054468951e-03f;
    p = fmaf(p, q, 2.186259337e-02f);
    p = fmaf(p, q, -5.591178698e-02f);
    p = fmaf(p, q, 9.642156725e-02f);
    p = fmaf(p, q, -1.390861325e-01f);
    p = fmaf(p, q, 1.994656231e-01f);
    p = fmaf(p, q, -3.332986048e-01f);
    p = fmaf(p, q, 9.999993355e-01f);
    return fast_sqrt(q) * p;
}

struct alignas(8) KeyArc {
    int key;    // bits of the arc start: starts are >= 0, so the integers order exactly like the floats
    float en;   // start + 2 alpha (may exceed one turn: the arc wraps)
};

// Exposed length contributed by this lane's arcs: sum_k max(0, start_k - max(W, P_k)) with
// P_k = max{ en_m : (start_m, m) < (start_k, k) lexicographically } — the order of the reference's sorted sweep
// (src/sasa_lr.c:367-408) without sorting, ties between equal starts broken by the arc index exactly as in lr_atom<T>.
// (Round 1 replaced the low key bits by the arc index to save the tie test; that let two starts within 128 ulp swap order,
// an error of up to that sliver per slice — 2e-4 A^2 at n_slices = 5.  The tie test costs one compare per FOUR arcs.)
// arcs[0..narc) hold (start bits, end); sentinels are appended so the loop can read four arcs at a time.  All lanes must
// call it (it synchronises the warp).
__device__ __forceinline__ float merge_keyed(KeyArc *arcs, int narc, float W, int lane)
{
    if (lane < 4) {
        KeyArc pad;
        pad.key = 0x7f800000;
        pad.en = 0.f;
        arcs[narc + lane] = pad;
    }
    __syncwarp();
    const float4 *quad = reinterpret_cast<const float4 *>(arcs);
    const int n_quads = (narc + 3) >> 2;
    float sum = 0.f;
    for (int k = lane; k < narc; k += 32) {
        const int my_key = arcs[k].key;
        const int kq = k >> 2;
        float P = W;
        for (int g = 0; g < n_quads; ++g) {
            const float4 lo = quad[2 * g], hi = quad[2 * g + 1];   // warp-uniform addresses: broadcast
            const int thr = my_key + (g < kq ? 1 : 0);             // arcs of earlier quads precede me on equal starts too
            if (__float_as_int(lo.x) < thr) P = fmaxf(P, lo.y);
            if (__float_as_int(lo.z) < thr) P = fmaxf(P, lo.w);
            if (__float_as_int(hi.x) < thr) P = fmaxf(P, hi.y);
            if (__float_as_int(hi.z) < thr) P = fmaxf(P, hi.w);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {                              // my own quad: equal starts with a smaller index
            const KeyArc o = arcs[4 * kq + j];
            if (4 * kq + j < k && o.key == my_key) P = fmaxf(P, o.en);
        }
        sum += fmaxf(__int_as_float(my_key) - P, 0.f);
    }
    return sum;
}

// General fp32 path (any neighbour count): all arcs of a slice are compacted to shared memory and merged
// pairwise with unique integer keys.
__device__ __forceinline__ double lr_atom_fast(const Rec4<float> *recs, KeyArc *arcs, int nn, double Ri_d, int ns, int lane)
{
    const float Ri = (float)Ri_d;
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    const double delta = 2.0 * Ri_d / ns;
    const unsigned lt = lanemask_lt();
    double acc = 0.0;

    for (int s = 0; s < ns; ++s) {
        const float zr = (float)(-Ri_d + (s + 0.5) * delta);
        const float az = fabsf(zr);
        const float a2 = (Ri - az) * (Ri + az);
        if (!(a2 > 0.f)) continue;
        const float a = fast_sqrt(a2);
        int narc = 0;
        bool buried = false;
        float my_max = 0.f;
        __syncwarp();                                      // a buried slice may have left arcs of other lanes in the array
        for (int base = 0; base < nn; base += 32) {
            const int j = base + lane;
            const bool valid = j < nn;
            const Rec4<float> r = recs[valid ? j : 0];     // {dz, R, dxy, beta}
            const float dj = fabsf(r.a - zr);
            const float b2 = (r.b - dj) * (r.b + dj);      // Rj'^2; > 0  <=>  dj < Rj
            const float b = fast_sqrt(fmaxf(b2, 0.f));
            const float d = r.c;
            const float ab = a + b;
            const float f1 = ab - d;                       // > 0  <=> circles touch      (src/sasa_lr.c:324)
            const float f3 = (d + a) - b;                  // < 0  <=> circle i inside j  (:327)
            const float f2 = (d + b) - a;                  // < 0  <=> circle j inside i  (:331)
            const bool touch = valid && b2 > 0.f && f1 > 0.f;
            const bool bur = touch && f3 < 0.f;
            const bool has = touch && !(f3 < 0.f) && !(f2 < 0.f);
            if (__any_sync(kFull, bur)) { buried = true; break; }
            const float N = f1 * f2, D = f3 * (ab + d);
            const float hi = fmaxf(N, D), lo = fminf(N, D);
            const float q = hi > 0.f ? fminf(lo * fast_rcp(hi), 1.f) : 0.f;
            const float u2 = 2.f * atan_sqrt01(q);
            const float alpha = N <= D ? u2 : pi - u2;
            float st = r.d - alpha;
            st += st < 0.f ? two_pi : 0.f;
            const float en = fmaf(2.f, alpha, st);
            const unsigned m = __ballot_sync(kFull, has);
            if (has) {
                const int slot = narc + __popc(m & lt);
                KeyArc arc;
                arc.key = __float_as_int(st);
                arc.en = en;
                arcs[slot] = arc;
                my_max = fmaxf(my_max, en);
            }
            narc += __popc(m);
        }
        if (buried) continue;
        if (narc == 0) {
            if (lane == 0) acc += (double)two_pi;
            continue;
        }
        my_max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(my_max)));  // non-negative floats order like uints
        acc += (double)merge_keyed(arcs, narc, fmaxf(my_max - two_pi, 0.f), lane);
        if (lane == 0) acc += (double)fmaxf(two_pi - my_max, 0.f);
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    return delta * Ri_d * acc;
}

// ---- step 2c: the common case, at most 32 K neighbours, K records per lane (instantiated with K = 3) ----
// The (z-sorted) records live in REGISTERS, K per lane, for all slices of the atom.  Per slice:
//   1. each group of 32 records evaluates its circle-circle configurations; a group whose neighbours are
//      all out of z-range is skipped (the records are z-sorted, so low slices skip the upper groups);
//   2. one vote: some circle swallows the slice circle -> the slice is buried (src/sasa_lr.c:327);
//   3. angles are measured in SECTORS (1/32 of the circle, so pi is exactly 16): every arc marks the
//      sectors it covers completely in a 32-bit mask, one REDUX.OR gives the sectors covered by a single
//      arc.  If that is all 32, the slice exposes nothing and we are done — on a dense globule this settles
//      ~90 % of the slices that survive step 2;
//   4. otherwise only arcs with an end in an uncovered sector can matter.  They, plus one synthetic arc per
//      run of covered sectors, are compacted (typically 2-4 arcs) and merged exactly as in lr_atom_fast.
// Exactness: a sector counts as covered only if ONE arc contains it entirely (integer sector borders are
// exact in fp32), so replacing the arcs inside covered sectors by the sectors themselves does not change
// the union.
// beta = atan2(dy, dx) + pi in sectors (src/sasa_lr.c:337, hoisted out of the slice loop).  Out of line: atan2f is ~40
// instructions and would be inlined once per record group.
__device__ __noinline__ float lr_beta_sectors(float dy, float dx)
{
    return (atan2f(dy, dx) + 3.141592653589793f) * 5.092958178940651f;   // sectors per radian
}

template <int K>
__device__ __forceinline__ void lr_prepare_sorted(Rec4<float> *recs, int nn, int lane)
{
    Rec4<float> mine[K];
    float z[K];
    int rank[K];
#pragma unroll
    for (int h = 0; h < K; ++h) {
        const int j = lane + 32 * h;
        mine[h] = recs[j < nn ? j : 0];                    // raw {dx, dy, dz, R}
        z[h] = j < nn ? mine[h].c : 3.0e38f;
        rank[h] = 0;
    }
    // rank of each of my records among all dz (ties by index): ONE pass over the list, every value broadcast once
#pragma unroll 1
    for (int k = 0; k < nn; ++k) {
        const float dzk = recs[k].c;                       // warp-uniform address: broadcast
#pragma unroll
        for (int h = 0; h < K; ++h) rank[h] += (dzk < z[h] || (dzk == z[h] && k < lane + 32 * h)) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < K; ++h)
        if (lane + 32 * h < nn) {
            Rec4<float> o;
            o.a = mine[h].c;
            o.b = mine[h].d;
            o.c = sqrtf(mine[h].a * mine[h].a + mine[h].b * mine[h].b);   // src/nb.c:440
            o.d = lr_beta_sectors(mine[h].b, mine[h].a);
            recs[rank[h]] = o;
        }
    __syncwarp();
}

struct HalfArc {
    float st, en;     // sectors; st in [0,32), en in [st, st+32]
    bool has, bur;
};

__device__ __forceinline__ HalfArc half_eval(const Rec4<float> &r, bool valid, float zr, float a)
{
    HalfArc h;
    h.st = 0.f; h.en = 0.f; h.has = false; h.bur = false;
    const float dj = fabsf(r.a - zr);
    const float b2 = (r.b - dj) * (r.b + dj);
    const bool act = valid && b2 > 0.f;
    if (__any_sync(kFull, act)) {
        const float b = fast_sqrt(fmaxf(b2, 0.f));
        const float d = r.c;
        const float ab = a + b;
        const float f1 = ab - d, f3 = (d + a) - b, f2 = (d + b) - a;
        const bool touch = act && f1 > 0.f;
        h.bur = touch && f3 < 0.f;
        h.has = touch && !(f3 < 0.f) && !(f2 < 0.f);
        // N, D kept in st/en until the burial vote is over
        h.st = f1 * f2;
        h.en = f3 * (ab + d);
    }
    return h;
}

// ---- marginal slices, found BEFORE the slice loop ------------------------------------------------------------------
// fp32 can neither decide nor measure a pair of slice circles that is within rounding distance of a tangency: with
// N = (a+b-d)(d+b-a) and D = (d+a-b)(a+b+d) (the three comparisons of src/sasa_lr.c:324-333 are the signs of their factors)
// the half-angle is alpha = 2 atan sqrt(N/D), a SQUARE ROOT of the gap, which amplifies the ~1e-7 relative rounding of a,
// b, d by 1/sqrt(q), q = min(|N|,|D|)/max(|N|,|D|).  Such slices are left out of the fp32 loop and redone in fp64
// (lr_redo_exact).  Round 2 first found them inside the loop (q < q_min for every pair of every slice: 4 instructions per
// 32 pairs plus the vote — 7 % of the kernel, most of it instruction-cache pressure in the hottest loop); but WHERE a
// pair is marginal is known in closed form: the circles of atoms i and j in the plane z are tangent exactly where that
// plane touches the intersection circle of the two SPHERES, i.e. at its lowest and highest point
//     z± = t dz/|D| ± rho dxy/|D|,   t = (|D|^2 + Ri^2 - Rj^2) / (2|D|),   rho^2 = Ri^2 - t^2
// (centre line D = (dxy, dz), circle centre at distance t from atom i, radius rho).  Near z±, N (or D) crosses zero with
// slope N' = 2 dz - 2 z dxy/a (D' = -2 dz - 2 z dxy/a), so q < q_min  <=>  |z - z±| < q_min |other| / |slope|.  One pass
// over the neighbours (two heights each, ~150 instructions per 32 neighbours, once per atom instead of once per
// slice) marks the slice planes inside those bands — with a factor 2 of safety, a floor for the fp32 records' own
// rounding, and the whole interval between two heights that nearly coincide.  tests/tools/marginal_model.py checks the
// construction against the brute-force q-test (it marks a superset, about twice as many slices).
// Everything is fp32 (MUFU rsqrt / sqrt / rcp): the bands carry a slack for the rounding of z± itself, which grows like
// 1/rho when the spheres barely intersect.
// recs: prepared records {dz, R, dxy, beta}; side[0..31]: the bits, zeroed by the caller.
__device__ __noinline__ void lr_mark_marginal(const Rec4<float> *recs, int *side, int nn, float Ri, int ns, float q_min, int lane)
{
    const float delta = 2.f * Ri / (float)ns, inv_delta = (float)ns / (2.f * Ri);
    for (int j = lane; j < nn; j += 32) {
        const Rec4<float> r = recs[j];
        const float dz = r.a, Rj = r.b, d = r.c;
        const float D3sq = fmaf(d, d, dz * dz);
        if (!(D3sq > 0.f)) continue;
        const float inv_D3 = rsqrtf(D3sq);
        const float t = 0.5f * (D3sq + (Ri - Rj) * (Ri + Rj)) * inv_D3;
        const float rho2 = (Ri - t) * (Ri + t);
        if (!(rho2 > 0.f)) continue;                       // one sphere inside the other: their surfaces do not meet
        const float rho = fast_sqrt(rho2);
        const float zc = t * dz * inv_D3, ext = rho * d * inv_D3;
        const float slack = 4.0e-6f * (1.f + fast_rcp(fmaxf(rho, 1.0e-3f)));
        if (2.f * ext < 1.0e-3f) {                         // the two heights nearly coincide: everything in between is marginal
            const float w = 1.0e-4f + slack;
            const int s0 = max((int)floorf((zc - ext - w + Ri) * inv_delta - 0.5f), 0);
            const int s1 = min((int)ceilf((zc + ext + w + Ri) * inv_delta - 0.5f), ns - 1);
            for (int s = s0; s <= s1; ++s) atomicOr(&side[s >> 5], 1 << (s & 31));
        }
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            const float z = k ? zc + ext : zc - ext;
            const float a = fast_sqrt(fmaxf((Ri - z) * (Ri + z), 1.0e-12f));
            const float zz = z - dz;
            const float b = fast_sqrt(fmaxf((Rj - zz) * (Rj + zz), 0.f));
            const float Nv = fabsf((a + b - d) * (d + b - a)), Dv = fabsf((d + a - b) * (a + b + d));
            const float zda = 2.f * z * d * fast_rcp(a);
            const float slope = Nv < Dv ? fabsf(2.f * dz - zda) : fabsf(-2.f * dz - zda);
            const float eps = fminf(fmaxf(2.f * q_min * fmaxf(Nv, Dv) * fast_rcp(fmaxf(slope, 1.0e-9f)), 2.0e-6f) + 2.0e-6f + slack, delta);
            const int sc = (int)rintf((z + Ri) * inv_delta - 0.5f);
            for (int s = max(sc - 1, 0); s <= min(sc + 1, ns - 1); ++s) {
                const float zs = fmaf((float)s + 0.5f, delta, -Ri);
                if (fabsf(zs - z) < eps) atomicOr(&side[s >> 5], 1 << (s & 31));
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned half_finish(HalfArc &h, float beta_s)
{
    // alpha in sectors: (32/2pi) * 2 atan(sqrt(q)); pi = 16 sectors
    const float N = h.st, D = h.en;
    const float hi = fmaxf(N, D), lo = fminf(N, D);
    const float q = hi > 0.f ? fminf(lo * fast_rcp(hi), 1.f) : 0.f;
    const float u2 = 10.185916357881302f * atan_sqrt01(q);
    const float alpha = N <= D ? u2 : 16.f - u2;
    float st = beta_s - alpha;
    st += st < 0.f ? 32.f : 0.f;
    st = st >= 32.f ? st - 32.f : st;
    const float en = fmaf(2.f, alpha, st);
    h.st = st;
    h.en = en;
    // sectors completely inside [st, en)
    const int js = (int)st;
    const int jf = js + ((float)js < st ? 1 : 0);
    const int cnt = (int)en - jf;
    unsigned mask = cnt >= 32 ? 0xffffffffu : __funnelshift_l((1u << (cnt & 31)) - 1u, (1u << (cnt & 31)) - 1u, jf & 31);
    return (h.has && cnt > 0) ? mask : 0u;
}

// The slice loop of the fp32 production path over ONE CHUNK of slices [s_begin, s_end).  An atom's slices are always
// summed chunk by chunk (kernel k_slices works on chunks from a task queue, the fused path walks them in order), so that
// the per-atom total has the same bits whichever way it was computed.
//   r[K], v[K]   this lane's K prepared records {dz, R, dxy, beta in sectors} (z-sorted) and their validity
//   marginal     bits of the slices that are left to the fp64 redo (word s / 32), or nullptr
// Returns the exposed angle of the chunk in sectors (the same value in every lane).
constexpr int kLrK = 3;   // at most 32 K = 96 neighbours in registers

__host__ __device__ inline int lr_chunk_slices(int ns) { return ns <= 64 * FSB200_CHUNK ? FSB200_CHUNK : (ns + 63) / 64; }   // at most 64 chunks per atom
__host__ __device__ inline int lr_n_chunks(int ns) { return (ns + lr_chunk_slices(ns) - 1) / lr_chunk_slices(ns); }

// area = delta Ri (exposed angle) (src/sasa_lr.c:360) from the fp32 path's angle in sectors (2 pi / 32 radians each) plus the
// fp64 redo's contribution.  ONE definition with explicit roundings (no FMA contraction), used by the split pipeline
// (k_finish) and by the fused path alike, so that an atom has the same bits whichever way it was computed.
__device__ __forceinline__ double lr_area(double Ri, int ns, double sectors, double extra)
{
    const double delta = 2.0 * Ri / ns;
    return __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(delta, Ri), sectors), 0.19634954084936207), extra);
}

template <int K>
__device__ __forceinline__ double lr_slices_chunk(const Rec4<float> (&r)[K], const bool (&v)[K], KeyArc *arcs,
                                                  const unsigned *marginal, int nn, double Ri_d, int ns, int s_begin, int s_end,
                                                  int lane)
{
    const float Ri = (float)Ri_d;
    const double delta = 2.0 * Ri_d / ns;
    const unsigned lt = lanemask_lt(lane);
    double acc = 0.0;                                      // exposed angle in sectors

    for (int s = s_begin; s < s_end; ++s) {
#if FSB200_EXACT_SLICES
        if (marginal && ((marginal[s >> 5] >> (s & 31)) & 1u)) continue;   // redone in fp64 (one uniform load)
#endif
        // slice centre relative to the atom centre (src/sasa_lr.c:305-307), formed in fp64 and rounded ONCE: a two-float
        // fp32 version (one ulp instead of half an ulp of z) was measured at 3.4e-4 A^2 on the 1M-atom shell (2e-4 bar)
        const float zr = (float)(-Ri_d + (s + 0.5) * delta);
        const float az = fabsf(zr);
        const float a2 = (Ri - az) * (Ri + az);
        if (!(a2 > 0.f)) continue;
        const float a = fast_sqrt(a2);
        HalfArc h[K];
        // ONE warp reduction settles everything the warp has to agree on before any angle is computed: bit 0 = some
        // circle swallows the slice circle (src/sasa_lr.c:327), bit 1+i = group i has an arc.
        bool bur = false;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            h[i].st = 0.f; h[i].en = 0.f; h[i].has = false; h[i].bur = false;
            if (i == 0 || nn > 32 * i) h[i] = half_eval(r[i], v[i], zr, a);
            bur = bur || h[i].bur;
        }
#if FSB200_FUSED_VOTES
        unsigned flags = bur ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < K; ++i) flags |= h[i].has ? (2u << i) : 0u;
        const unsigned agreed = __reduce_or_sync(kFull, flags);
#else
        unsigned agreed = __any_sync(kFull, bur) ? 1u : 0u;
        if (!agreed) {
#pragma unroll
            for (int i = 0; i < K; ++i) agreed |= __any_sync(kFull, h[i].has) ? (2u << i) : 0u;
        }
#endif
        if (agreed & 1u) continue;                                         // buried slice
        if ((agreed >> 1) == 0u) {                                         // free circle
            if (lane == 0) acc += 32.0;
            continue;
        }
        bool any[K];
#pragma unroll
        for (int i = 0; i < K; ++i) any[i] = (agreed >> (1 + i)) & 1u;
        unsigned mask = 0u;
#pragma unroll
        for (int i = 0; i < K; ++i)
            if (any[i]) mask |= half_finish(h[i], r[i].d);
        const unsigned full = __reduce_or_sync(kFull, mask);
        if (full == 0xffffffffu) continue;                                 // every sector inside some arc
        // arcs with an end in an uncovered sector, plus one synthetic arc per run of covered sectors
        int narc = 0;
        float my_max = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const bool rel = h[i].has && (!((full >> ((int)h[i].st & 31)) & 1u) || !((full >> ((int)h[i].en & 31)) & 1u));
            const unsigned b = __ballot_sync(kFull, rel);
            if (rel) {
                const int slot = narc + __popc(b & lt);
                KeyArc arc;
                arc.key = __float_as_int(h[i].st);
                arc.en = h[i].en;
                arcs[slot] = arc;
                my_max = fmaxf(my_max, h[i].en);
            }
            narc += __popc(b);
        }
        {
            const bool syn = ((full >> lane) & 1u) && !((full >> ((lane + 31) & 31)) & 1u);
            const unsigned b = __ballot_sync(kFull, syn);
            if (syn) {
                const unsigned rot = __funnelshift_r(full, full, lane);    // covered run starts at bit 0
                const int ones = __ffs(~rot) - 1;                          // full != all ones here
                const int slot = narc + __popc(b & lt);
                const float st = (float)lane, en = (float)(lane + ones);
                KeyArc arc;
                arc.key = __float_as_int(st);
                arc.en = en;
                arcs[slot] = arc;
                my_max = fmaxf(my_max, en);
            }
            narc += __popc(b);
        }
        my_max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(my_max)));
        acc += (double)merge_keyed(arcs, narc, fmaxf(my_max - 32.f, 0.f), lane);
        if (lane == 0) acc += (double)fmaxf(32.f - my_max, 0.f);
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    return acc;
}

// all chunks of one atom in order (the fused path); recs: the warp's prepared records in shared memory
template <int K>
__device__ __forceinline__ double lr_atom_fastk(const Rec4<float> *recs, KeyArc *arcs, const unsigned *marginal, int nn, double Ri_d,
                                                int ns, int lane)
{
    Rec4<float> r[K];                                      // this lane's K records, for all slices of the atom
    bool v[K];
#pragma unroll
    for (int h = 0; h < K; ++h) {
        v[h] = lane + 32 * h < nn;
        r[h] = recs[v[h] ? lane + 32 * h : 0];
    }
    const int S = lr_chunk_slices(ns);
    double sum = 0.0;
    for (int s0 = 0; s0 < ns; s0 += S) sum += lr_slices_chunk<K>(r, v, arcs, marginal, nn, Ri_d, ns, s0, min(s0 + S, ns), lane);
    return sum;                                            // exposed angle in SECTORS: lr_area() turns it into an area
}

// ---- step 3: Shrake & Rupley ---------------------------------------------------------------------
// fp32 decision band of this atom: the dot product and t each carry at most ~4e-7 * (|D|+|t|) of
// rounding; every test closer than 10x that to the threshold is re-decided exactly in fp64.
template <typename T>
__device__ __forceinline__ float sr_band(const Rec4<T> *recs, int nn, int lane)
{
    float scale = 0.f;
    for (int j = lane; j < nn; j += 32) {
        const Rec4<T> r = recs[j];
        const float dx = (float)r.a, dy = (float)r.b, dz = (float)r.c;
        scale = fmaxf(scale, sqrtf(dx * dx + dy * dy + dz * dz) + fabsf((float)r.d));
    }
    for (int o = 16; o; o >>= 1) scale = fmaxf(scale, __shfl_xor_sync(kFull, scale, o));
    return 4e-6f * scale + 1e-30f;
}

// the reference's own expression for "test point q of self is inside neighbour cand"
// (src/sasa_sr.c:297-299 scale+translate via src/coord.c:306-342, then :312-317), no FMA contraction
__device__ __forceinline__ bool sr_exact_hidden(const double *pd, int q, const Self &s, const double4 cand)
{
    const double px = __dadd_rn(__dmul_rn(pd[3 * q], s.R), s.x);
    const double py = __dadd_rn(__dmul_rn(pd[3 * q + 1], s.R), s.y);
    const double pz = __dadd_rn(__dmul_rn(pd[3 * q + 2], s.R), s.z);
    const double dx = __dsub_rn(px, cand.x), dy = __dsub_rn(py, cand.y), dz = __dsub_rn(pz, cand.z);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double r2 = __dmul_rn(cand.w, cand.w);           // src/sasa_sr.c:146
    return !(d2 > r2);
}

template <typename T>
__device__ __forceinline__ double sr_atom(Rec4<T> *recs, int *cidx, const double4 *cand_base, int nn,
                                          const Self &s, int npts, const float4 *pf, const double *pd, int lane)
{
    int exposed = 0;
    if (sizeof(T) == 8) {  // fp64 mode: the reference's expression for every pair
        for (int q0 = 0; q0 < npts; q0 += 32) {
            const int q = q0 + lane;
            bool bur = q >= npts;
            for (int k = 0; k < nn; ++k) {
                if (!bur) bur = sr_exact_hidden(pd, q, s, cand_base[cidx[k]]);
                if (__all_sync(kFull, bur)) break;
            }
            exposed += __popc(__ballot_sync(kFull, !bur));
        }
    } else {
        // fp32 production path.  The test points arrive PATCH-ORDERED from the host (32 consecutive
        // points form a compact patch of the sphere; the count does not depend on the order), so a
        // round of 32 lanes is usually hidden by one or two neighbours and leaves the loop early.
        // Neighbours are tested four at a time with one vote per block; the block that finished a
        // round is where the next (adjacent) patch starts — the warp-wide version of the reference's
        // "begin with the neighbour that hid the previous point" (src/sasa_sr.c:305-328).
        const float band = sr_band<T>(recs, nn, lane);
        const int nn4 = (nn + 3) & ~3;
        if (lane < nn4 - nn) {                             // pad with neighbours that hide nothing
            Rec4<T> pad;
            pad.a = 0; pad.b = 0; pad.c = 0; pad.d = (T)3.0e38f;
            recs[nn + lane] = pad;
            cidx[nn + lane] = 0;
        }
        __syncwarp();
        int k_start = 0;
        for (int q0 = 0; q0 < npts; q0 += 32) {
            const int q = q0 + lane;
            const bool active = q < npts;
            const float4 u = active ? pf[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            bool bur = !active;
            int kb = k_start;
            for (int done = 0; done < nn4; done += 4) {
                const Rec4<T> r0 = recs[kb], r1 = recs[kb + 1], r2 = recs[kb + 2], r3 = recs[kb + 3];   // broadcast
                const float d0 = fmaf(u.x, (float)r0.a, fmaf(u.y, (float)r0.b, u.z * (float)r0.c)) - (float)r0.d;
                const float d1 = fmaf(u.x, (float)r1.a, fmaf(u.y, (float)r1.b, u.z * (float)r1.c)) - (float)r1.d;
                const float d2 = fmaf(u.x, (float)r2.a, fmaf(u.y, (float)r2.b, u.z * (float)r2.c)) - (float)r2.d;
                const float d3 = fmaf(u.x, (float)r3.a, fmaf(u.y, (float)r3.b, u.z * (float)r3.c)) - (float)r3.d;
                const float dmax = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
                if (dmax > band) bur = true;
                // anything within the rounding band of its threshold is re-decided exactly (rare)
                const bool amb = !bur && dmax >= -band;
                if (__any_sync(kFull, amb)) {
                    if (amb) {
                        if (fabsf(d0) <= band) bur = sr_exact_hidden(pd, q, s, cand_base[cidx[kb]]);
                        if (!bur && fabsf(d1) <= band) bur = sr_exact_hidden(pd, q, s, cand_base[cidx[kb + 1]]);
                        if (!bur && fabsf(d2) <= band) bur = sr_exact_hidden(pd, q, s, cand_base[cidx[kb + 2]]);
                        if (!bur && fabsf(d3) <= band) bur = sr_exact_hidden(pd, q, s, cand_base[cidx[kb + 3]]);
                    }
                }
                if (__all_sync(kFull, bur)) {
                    k_start = kb;
                    break;
                }
                kb += 4;
                if (kb >= nn4) kb = 0;
            }
            exposed += __popc(__ballot_sync(kFull, !bur));
        }
    }
    // src/sasa_sr.c:337, same expression order
    return (4.0 * 3.14159265358979323846 * s.R * s.R * exposed) / npts;
}

// ---- buried-atom certificate ---------------------------------------------------------------------------
// Most atoms of a large structure have NO exposed surface (92 % of the 100k-atom benchmark globule), and
// for them every slice ends "buried" or "fully covered" and every test point is hidden.  This routine proves
// that outcome for a whole atom at once, at ~5 % of the cost of integrating it:
//   * kCertPoints = 128 probe directions u_k — 64 antipodal pairs, Lloyd-relaxed under that symmetry (cert_dirs.inc,
//     tests/golden/make_cert_directions.py) — cover the unit sphere with patches of angular radius rho = 12.5 deg:
//     measured covering radius 12.17 deg (tests/test_certificate.py measures it from the library's own table and asserts
//     it against kCertCos), i.e. a margin of 0.3 deg + the fp32 safety term of the threshold below;
//   * neighbour a hides the cap of half-angle theta_a around D_a, cos(theta_a) = t_a/|D_a|,
//     t_a = (Ri^2+|D_a|^2-Ra^2)/(2Ri); it hides the WHOLE patch k iff angle(u_k, D_a) <= theta_a - rho, i.e.
//     u_k.D_a >= |D_a| cos(theta_a - rho) = t_a cos(rho) + sqrt(|D_a|^2 - t_a^2) sin(rho)  (=: t'_a, plus an
//     fp32 safety margin);
//   * if every patch is hidden entirely by a single neighbour, every point of the sphere is strictly inside
//     some neighbour sphere, so every Lee-Richards slice circle is covered with positive overlap and every
//     Shrake-Rupley point is hidden: the reference's area is exactly 0 (its arc sweep returns 0 + 2pi - 2pi,
//     src/sasa_lr.c:389-408; its point count is 0) and so is ours.
// The test is conservative — a certified atom is always truly buried (never the other way round) — so it
// changes no result, only the time (tests: certificate on/off give bit-identical arrays).
constexpr float kCertCos = 0.97629601f;   // cos(12.5 deg)
constexpr float kCertSin = 0.21643961f;   // sin(12.5 deg)

// lanes = directions.  The probe set is 64 antipodal pairs (cert_dirs.inc: Lloyd-relaxed, covering radius 12.17 deg);
// a lane owns two pairs, i.e. four directions, and ONE dot product decides a pair: u.D >= t' hides the patch of u,
// -(u.D) >= t' hides the patch of -u.  Neighbours whose cap is wider than a patch are compacted into a short list in
// shared memory (at most kCertList = 64), wide caps first; the loop over the list costs one broadcast load per cap and
// ends as soon as every direction has found a cap that hides its whole patch.
//
// History (profiles/): v1 lanes = directions over ALL neighbours; v2 lanes = neighbours, 128 directions from constant
// memory with one vote each (1700 warp instructions per atom); this version ~800.
template <typename T, bool HAS_T>   // HAS_T: recs hold {dx,dy,dz,t} (S&R); otherwise raw {dx,dy,dz,Ra} (L&R)
__device__ __forceinline__ bool certify_buried(const Rec4<T> *recs, float4 *list, int nn, float Ri, int lane,
                                               const float4 *__restrict__ dirs)
{
    const unsigned lt = lanemask_lt();
    int n_useful = 0, n_wide = 0;
    bool inside = false;
    // Short list = neighbours whose cap is wider than a patch, WIDE caps (half-angle > ~50 deg: the bonded and second-
    // shell neighbours, which hide most of the sphere between them) compacted from the front of the list, the others
    // from the back, so the coverage loop meets the wide ones first.  If more than kCertList qualify (dense packing,
    // explicit hydrogens), the bar is raised to caps at least 20, 35, 50 degrees wider, until they fit: the widest caps
    // are the ones that hide whole patches anyway.
    constexpr float kWide = 0.65f;   // cos(49.5 deg); measured optimum on the 100k globule (0.5: 0.508 ms, 0.65: 0.467, 0.85: 0.484)
    const float inv_2Ri = 0.5f / Ri;
    // (The loops of this routine are deliberately NOT unrolled and use the approximate MUFU square root: it runs once per
    //  atom, in warps that are all at different places of it, so its footprint in the instruction caches matters more than
    //  its instruction count — ncu, round 2: half of the stall samples inside the unrolled version were `no_instruction`.
    //  Every approximation is covered by the safety margin of the threshold below, 1e-5 d + 1e-6, ~80 ulp.)
#pragma unroll 1
    for (int attempt = 0; attempt < 4; ++attempt) {
        const float bar = attempt == 0 ? kCertCos : attempt == 1 ? 0.84339145f : attempt == 2 ? 0.67559021f : 0.46174861f;   // cos(12.5, 32.5, 47.5, 62.5 deg)
        int n_back = 0;
        n_useful = 0;
        if (attempt > 0) __syncwarp();                     // the list is rewritten by other lanes than in the last attempt
#pragma unroll 1
        for (int base = 0; base < nn; base += 32) {
            const int j = base + lane;
            float4 e = make_float4(0.f, 0.f, 0.f, 3.0e38f);
            bool useful = false, wide = true;
            if (j < nn) {
                const Rec4<T> r = recs[j];
                e.x = (float)r.a; e.y = (float)r.b; e.z = (float)r.c;
                const float d2 = e.x * e.x + e.y * e.y + e.z * e.z;
                const float d = fast_sqrt(d2);
                const float t = HAS_T ? (float)r.d : (Ri * Ri + d2 - (float)r.d * (float)r.d) * inv_2Ri;
                if (t < -d - 1e-4f * (d + Ri)) inside = true;  // sphere i lies strictly inside sphere a (coincident equal
                                                               // spheres, t = d = 0, do NOT count: the reference decides
                                                               // their points one rounding at a time)
                else if (t < d * bar) {                        // cap wide enough for this attempt
                    e.w = t * kCertCos + fast_sqrt(fmaxf(d2 - t * t, 0.f)) * kCertSin + 1e-5f * d + 1e-6f;
                    useful = true;
                    wide = attempt > 0 || t < d * kWide;
                }
            }
            const unsigned mf = __ballot_sync(kFull, useful && wide), mb = __ballot_sync(kFull, useful && !wide);
            n_useful += __popc(mf);
            n_back += __popc(mb);
            if (useful) {   // (the bounds keep front and back apart even in an attempt that overflows and is discarded)
                const int slot = wide ? n_useful - __popc(mf) + __popc(mf & lt) : kCertList - 1 - (n_back - __popc(mb) + __popc(mb & lt));
                if (wide ? slot < kCertList - n_back : slot >= n_useful) list[slot] = e;
            }
        }
        n_wide = n_useful;
        n_useful += n_back;
        if (n_useful <= kCertList) break;
    }
    if (__any_sync(kFull, inside)) return true;
    if (n_useful == 0 || n_useful > kCertList) return false;   // nothing to work with / too many for the short list
    __syncwarp();
    const float4 u0 = __ldg(dirs + lane), u1 = __ldg(dirs + lane + 32);   // two antipodal pairs per lane, L1-resident table
    unsigned hidden = 0u;                                                 // bit 0..3: patch of +u0, -u0, +u1, -u1 hidden
    const float4 none = make_float4(0.f, 0.f, 0.f, 3.0e38f);
    // Phase 1 — lanes = directions, loop over the WIDE caps (front of the list): a dozen caps hide most of the sphere.
#pragma unroll 1
    for (int j = 0; j < n_wide; j += 2) {
        const float4 e = list[j];
        const float4 f = j + 1 < n_wide ? list[j + 1] : none;
        const float e0 = fmaf(u0.x, e.x, fmaf(u0.y, e.y, u0.z * e.z)), e1 = fmaf(u1.x, e.x, fmaf(u1.y, e.y, u1.z * e.z));
        const float f0 = fmaf(u0.x, f.x, fmaf(u0.y, f.y, u0.z * f.z)), f1 = fmaf(u1.x, f.x, fmaf(u1.y, f.y, u1.z * f.z));
        hidden |= ((e0 >= e.w || f0 >= f.w) ? 1u : 0u) | ((-e0 >= e.w || -f0 >= f.w) ? 2u : 0u) |
                  ((e1 >= e.w || f1 >= f.w) ? 4u : 0u) | ((-e1 >= e.w || -f1 >= f.w) ? 8u : 0u);
    }
    // Phase 2 — roles swapped for the few directions still open: lanes = the remaining (narrower) caps, two per lane
    // in registers, one direction at a time (uniform table load), one vote each.  The first direction nobody hides ends
    // the attempt, which is also what makes surface atoms cheap to reject.
    const int n_back = n_useful - n_wide;
    const float4 c0 = lane < n_back ? list[kCertList - 1 - lane] : none;
    const float4 c1 = lane + 32 < n_back ? list[kCertList - 1 - (lane + 32)] : none;
    __syncwarp();                                              // the list's memory is reused by the integrators
    bool ok = true;
#pragma unroll 1
    for (int q = 0; q < 4 && ok; ++q) {
        unsigned open = __ballot_sync(kFull, !((hidden >> q) & 1u));
        const float sign = (q & 1) ? -1.f : 1.f;
        while (open) {
            const int l = __ffs(open) - 1;
            open &= open - 1;
            const float4 u = __ldg(dirs + l + (q >= 2 ? 32 : 0));
            const float s0 = sign * fmaf(u.x, c0.x, fmaf(u.y, c0.y, u.z * c0.z));
            const float s1 = sign * fmaf(u.x, c1.x, fmaf(u.y, c1.y, u.z * c1.z));
            if (!__any_sync(kFull, s0 >= c0.w || s1 >= c1.w)) {
                ok = false;
                break;
            }
        }
    }
    return ok;
}

// ---- one atom after its neighbours have been gathered ------------------------------------------------
template <int ALG, typename T> struct WarpMem {
    Rec4<T> *recs;
    Arc<T> *arcs;      // L&R: arc_cap(cap) arcs of the current slice
    T *starts;         // L&R fp32 fast path: exact arc starts
    int *cidx;         // candidate positions: S&R after the records; L&R in the LAST cap ints of the warp's region (over the
                       // tail of `starts`, dead until the slices begin; the certificate's list sits at the front of `arcs`)
    float4 *cert_list; // short list of the buried-atom certificate (kCertList entries)
    __device__ __forceinline__ WarpMem(unsigned char *mem, int cap)
    {
        recs = reinterpret_cast<Rec4<T> *>(mem);
        arcs = reinterpret_cast<Arc<T> *>(mem + (size_t)cap * sizeof(Rec4<T>));
        starts = reinterpret_cast<T *>(mem + (size_t)cap * sizeof(Rec4<T>) + (size_t)arc_cap(cap) * sizeof(Arc<T>));
        cidx = ALG == 0 ? reinterpret_cast<int *>(mem + WarpLayout<ALG, T>::bytes(cap) - (size_t)cap * sizeof(int))
                        : reinterpret_cast<int *>(mem + (size_t)cap * sizeof(Rec4<T>));
        cert_list = ALG == 0 ? reinterpret_cast<float4 *>(arcs)
                             : reinterpret_cast<float4 *>(mem + (size_t)cap * (sizeof(Rec4<T>) + sizeof(int)));
    }
};

// Slices the fp32 path marked as marginal, redone with the generic integrator in fp64: the records are rebuilt from the
// fp64 atoms in global memory (differences in the atom-local frame, as the gather forms them) INTO the warp's own
// shared-memory region, whose fp32 contents are dead by now.  The gather left each neighbour's candidate index in `cidx`
// (the tail of the warp's region, untouched by the slice loop): a tile index when the neighbourhood was staged — turned into
// a sorted position with the fill's run table, which the warp still holds in registers (`pw`, one word per lane) although
// the tile itself has long been recycled — or the sorted position itself.  Out of line and called after the slice loop, so
// that its register needs do not touch the hot loop.  Needs 48 nn bytes (records + arcs): nn <= 94 of the <= 96 this path serves.
constexpr int kRedoMaxNeighbours = 94;
constexpr int kSideWords = 96;   // cidx[96..127]: marginal-slice bits, cidx[128..159]: the fill's run table (kNbCap = 160 ints in all)
__device__ __noinline__ double lr_redo_exact(const double4 *atoms, unsigned char *warp_mem, const int *cidx, int nn,
                                             double sx, double sy, double sz, double Ri, int ns, int lane)
{
    Rec4<double> *recs = reinterpret_cast<Rec4<double> *>(warp_mem);
    Arc<double> *arcs = reinterpret_cast<Arc<double> *>(warp_mem + (size_t)kRedoMaxNeighbours * sizeof(Rec4<double>));
    const unsigned marginal = (unsigned)cidx[kSideWords + lane];
    const int pw = cidx[kSideWords + 32 + lane];
    const bool staged = __shfl_sync(kFull, pw, kWStaged) != 0;
    int pos[3];
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        const bool valid = lane + 32 * h < nn;
        const int c = valid ? cidx[lane + 32 * h] : 0;
        int r = 0;
#pragma unroll
        for (int q = 1; q < 9; ++q) r += c >= __shfl_sync(kFull, pw, kWOff + q) ? 1 : 0;   // runs are laid out in order
        const int p = __shfl_sync(kFull, pw, kWBegin + r) + (c - __shfl_sync(kFull, pw, kWOff + r));
        pos[h] = valid ? (staged ? p : c) : -1;
    }
    __syncwarp();                                          // the indices sit where the fp64 arcs go
#pragma unroll
    for (int h = 0; h < 3; ++h)
        if (pos[h] >= 0) {
            const double4 q = atoms[pos[h]];
            const double dx = q.x - sx, dy = q.y - sy, dz = q.z - sz;
            Rec4<double> o;
            o.a = dz;
            o.b = q.w;
            o.c = sqrt(dx * dx + dy * dy);
            o.d = atan2(dy, dx) + 3.141592653589793;
            recs[lane + 32 * h] = o;
        }
    __syncwarp();
    return lr_atom<double>(recs, arcs, nn, Ri, ns, lane, true, marginal);
}

// ---- split pipeline: task records ------------------------------------------------------------------------------------
// One record per atom that has to be integrated (fp32 Lee-Richards, <= 96 neighbours), written by k_integrate:
//   TodoHeader | marginal bits (ceil(ns / 32) words, padded to 16 B) | partial sums (one double per chunk) |
//   prepared records {dz, R, dxy, beta} (16 B each, z-sorted) | sorted positions of the neighbours (for the fp64 redo)
struct alignas(16) TodoHeader {
    int pos, nn, has_marginal, pad0;
    double Ri, x, y, z;
    double extra;       // exposed angle (radians) of the marginal slices, from the fp64 redo task; 0 if there are none
    int pad[2];
};
static_assert(sizeof(TodoHeader) == 64, "TodoHeader layout");

struct TodoLayout {
    unsigned marginal, partial, recs, gpos, bytes;   // byte offsets inside a record
};
__host__ __device__ inline TodoLayout todo_layout(int ns, int nn)
{
    TodoLayout L;
    L.marginal = (unsigned)sizeof(TodoHeader);
    L.partial = L.marginal + (((unsigned)(ns + 31) / 32 * 4 + 15) & ~15u);
    L.recs = L.partial + (((unsigned)lr_n_chunks(ns) * 8 + 15) & ~15u);
    L.gpos = L.recs + (unsigned)nn * 16;
    L.bytes = (L.gpos + (unsigned)nn * 4 + 15) & ~15u;
    return L;
}

// the epilogue of every path: one area to the caller's array (and its mirrors / owner, see IntegrateArgs)
__device__ __forceinline__ void write_area(const Workspace &ws, const IntegrateArgs &args, int pos, double area)
{
    const int i = ws.perm[pos];
    const int idx = args.sorted_output ? pos : i;
    if (args.owner_slice > 0) {
        args.peer_out[idx / args.owner_slice][idx] = area;                     // to the GPU that owns this part of the result
    } else {
        args.out[idx] = area;
        if (area != 0.0 || !args.peer_skip_zero)
            for (int q = 0; q < args.n_peer_out; ++q) args.peer_out[q][idx] = area;   // the all-gather, store by store (NVLink)
    }
}

// k_integrate, atom not certified: hand it over to k_slices.  recs: prepared records in the warp's shared memory; side[0..31]:
// marginal bits; cidx: candidate indices of the gather (tile indices when staged, see lr_redo_exact); pw: the fill's run table.
// Returns false when the pool is full (the caller then integrates the atom itself).
__device__ __forceinline__ bool lr_emit_task(const IntegrateArgs &args, const Rec4<float> *recs, const int *side, const int *cidx,
                                             int pw, int nn, const Self &s, int pos, int lane)
{
    const int ns = args.resolution;
    const TodoLayout L = todo_layout(ns, nn);
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(&args.todo_ctl->pool_head, (unsigned long long)L.bytes);
    off = __shfl_sync(kFull, off, 0);
    if (off + L.bytes > args.todo_cap) return false;
    unsigned char *rec = args.todo_pool + off;
    const int words = (ns + 31) / 32;
    const unsigned mbits = lane < words && ns <= 1024 ? (unsigned)side[lane] : 0u;
    const bool any_marginal = __any_sync(kFull, mbits != 0u);
    if (lane == 0) {
        TodoHeader h;
        h.pos = pos; h.nn = nn; h.has_marginal = any_marginal ? 1 : 0; h.pad0 = 0;
        h.Ri = s.R; h.x = s.x; h.y = s.y; h.z = s.z;
        h.extra = 0.0;
        h.pad[0] = h.pad[1] = 0;
        *reinterpret_cast<TodoHeader *>(rec) = h;
    }
    if (lane < ((words + 3) & ~3)) reinterpret_cast<unsigned *>(rec + L.marginal)[lane] = mbits;
    const bool staged = __shfl_sync(kFull, pw, kWStaged) != 0;
    for (int j0 = 0; j0 < nn; j0 += 32) {
        const int j = j0 + lane;
        const int c = j < nn ? cidx[j] : 0;
        int r = 0;
#pragma unroll
        for (int q = 1; q < 9; ++q) r += c >= __shfl_sync(kFull, pw, kWOff + q) ? 1 : 0;   // runs are laid out in order
        const int p = __shfl_sync(kFull, pw, kWBegin + r) + (c - __shfl_sync(kFull, pw, kWOff + r));
        if (j < nn) {
            reinterpret_cast<int *>(rec + L.gpos)[j] = staged ? p : c;
            reinterpret_cast<Rec4<float> *>(rec + L.recs)[j] = recs[j];
        }
    }
    __threadfence();                                       // the record is complete before it becomes visible in the list
    __syncwarp();
    if (lane == 0) {
        args.todo_list[atomicAdd(&args.todo_ctl->n_todo, 1)] = off;
        if (any_marginal) args.todo_list[args.todo_redo_base + atomicAdd(&args.todo_ctl->n_redo, 1)] = off;
    }
    return true;
}

template <int ALG, typename T, bool FAST>
__device__ __forceinline__ bool finish_atom(const Workspace &ws, const IntegrateArgs &args, const WarpMem<ALG, T> &wm,
                                            const double4 *cand_base, const Self &s, int nn, int cap, int pos,
                                            bool allow_overflow, int lane, int pw = 0)
{
    __syncwarp();
    if (nn > cap) {
        if (allow_overflow && lane == 0) {
            ws.overflow[atomicAdd(ws.counters + kCtrOverflow, 1)] = pos;
            atomicMax(ws.counters + kCtrMaxCand, nn);
        }
        return false;
    }
    double area = 0.0;
    bool certified = false;
    if (s.R > 0.0) {
        if constexpr (FAST && sizeof(T) == 4) {
            if (args.cert_points != nullptr && nn > 0)
                certified = certify_buried<T, ALG == 1>(wm.recs, wm.cert_list, nn, (float)s.R, lane, args.cert_points);
        }
        if (certified) {
            // area stays 0: proved completely buried
        } else if (ALG == 0) {
            if constexpr (FAST && sizeof(T) == 4) {
                Rec4<float> *recs = reinterpret_cast<Rec4<float> *>(wm.recs);
                KeyArc *arcs = reinterpret_cast<KeyArc *>(wm.arcs);
                if (nn <= 96) {                            // ONE instantiation (K = 3) for all of them: the kernel is
                    const bool can_redo = FSB200_EXACT_SLICES && nn <= kRedoMaxNeighbours && args.resolution <= 1024;
                    int *side = wm.cidx + kSideWords;
                    side[lane] = 0;
                    side[32 + lane] = pw;
                    lr_prepare_sorted<3>(recs, nn, lane);           // instruction-cache sensitive (ncu: no_instruction stalls
                    if (can_redo) {                                 // with K = 1, 2, 3 side by side)
                        // thick slices (low resolution) need a wider band: an angular error e costs delta Ri e of area.  Floor
                        // 3e-6 (1M atoms, PDB-rounded, n = 100: 4.5e-4 -> 7e-5 A^2), growing with (delta Ri)^2 to 2.4e-5 at n = 5
                        const float dR = (float)(2.0 * s.R * s.R / args.resolution);
                        lr_mark_marginal(recs, side, nn, (float)s.R, args.resolution, fmaxf(FSB200_NEAR_FLOOR, FSB200_NEAR_SCALE * dR * dR), lane);
                    }
                    if (args.todo_pool != nullptr && lr_emit_task(args, recs, side, wm.cidx, pw, nn, s, pos, lane)) {
                        if (lane == 0 && args.nn_out) args.nn_out[ws.perm[pos]] = nn;
                        return false;                               // k_slices (and k_redo) finish this atom
                    }
                    if (lane == 0 && args.todo_pool != nullptr) atomicAdd(&args.todo_ctl->n_inline, 1);
                    const double sectors = lr_atom_fastk<3>(recs, arcs, can_redo ? reinterpret_cast<const unsigned *>(side) : nullptr,
                                                            nn, s.R, args.resolution, lane);
                    double extra = 0.0;
                    if (can_redo && __any_sync(kFull, side[lane] != 0)) {
                        if (lane == 0) atomicAdd(ws.counters + kCtrMarginal, 1);
                        extra = lr_redo_exact(ws.atoms, reinterpret_cast<unsigned char *>(wm.recs), wm.cidx, nn, s.x, s.y, s.z,
                                              s.R, args.resolution, lane);
                    }
                    area = lr_area(s.R, args.resolution, sectors, extra);
                } else {
                    lr_prepare<float>(recs, nn, lane);
                    area = lr_atom_fast(recs, arcs, nn, s.R, args.resolution, lane);
                }
            } else {
                lr_prepare<T>(wm.recs, nn, lane);
                area = lr_atom<T>(wm.recs, wm.arcs, nn, s.R, args.resolution, lane);
            }
        } else {
            area = sr_atom<T>(wm.recs, wm.cidx, cand_base, nn, s, args.resolution, args.points_f, args.points_d, lane);
        }
    }
    if (lane == 0) {
        write_area(ws, args, pos, area);
        if (args.nn_out) args.nn_out[ws.perm[pos]] = nn | (certified ? (1 << 30) : 0);
    }
    return certified;
}

__device__ __forceinline__ Self load_self(const double4 me)
{
    Self s;
    s.x = me.x; s.y = me.y; s.z = me.z; s.R = me.w;
    return s;
}

// 9 runs of the sorted atom array covering the 27 cells around `cell` (local id) of structure g
__device__ __forceinline__ void cell_run(const Workspace &ws, const GridDesc &g, int cx, int cy, int cz, int r,
                                         int *begin, int *count)
{
    const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
    *begin = 0;
    *count = 0;
    if (y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) return;
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    const int row = g.cell_base + g.dim[0] * (y + g.dim[1] * z);
    const int b = ws.cell_start[row + x0], e = ws.cell_start[row + x1 + 1];
    *begin = b;
    *count = e - b;
}

// ---- the persistent kernel: a ring of TMA-staged tiles, warps claim atoms one by one -------------------
// Each CTA owns kRingSlots tile buffers.  A "fill" is one work item (<= kItemAtoms atoms of one cell) plus
// its 27-cell neighbourhood, staged into the slot fill % kRingSlots by nine cp.async.bulk copies that
// complete on the slot's mbarrier.  Fills are numbered 0,1,2,... per CTA and `cur` is the fill currently
// open for claiming.  A warp
//   1. reads cur, waits for that fill's mbarrier phase, claims one atom with a compare-and-swap on the
//      slot's packed (fill number, next atom) word — a warp can therefore never claim from a slot that has
//      moved on, however long it was away;
//   2. gathers the atom's neighbours from the tile into its private list.  That is the LAST use of the tile:
//      the warp that completes the last gather of a fill immediately refills the slot with fill + kRingSlots
//      from the global queue, and only then
//   3. integrates the atom (the long part: slices or test points) — nobody waits for it.
// History (profiles/): per-item __syncthreads cost 40 % of all warp samples because buried atoms finish long
// before surface atoms; a first ring that refilled a slot only after every warp had LEFT it still left
// 18 % of the samples (and 25 % of the issued instructions) in the mbarrier wait.
struct Slot {
    int claim;        // (fill & 0xffff) << 16 | n_atoms << 8 | next unclaimed atom: ONE word, so that "is there an
                      // atom left in fill f" and the claim itself are decided on the same atomic snapshot
    int gathered;     // atoms of this fill whose neighbour gather is complete
    int dead;         // the global queue was empty when this slot was last refilled: no more fills here
    int pad;
    int w[32];        // the fill's description, ONE WORD PER LANE: a warp writes / reads it with a single warp-wide
                      // atomic and looks fields up with shuffles (kW* below)
};
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Memory-model hygiene of the ring.  EVERY word of a slot that another warp may read or write concurrently — cur, claim,
// gathered, dead and the 32 payload words — is accessed only through block-scope atomic read-modify-write operations with
// acquire / release semantics: a reader that observes a published claim word (acquire) is guaranteed to see the payload
// written before it (release); the last gather's acq_rel increment of `gathered` orders every reader of a fill before the
// refill that overwrites it; and compute-sanitizer's racecheck sees atomics on both sides of every pair.  Round 1 used plain
// stores + __threadfence_block() against volatile polls: it worked on sm_100 but was a formal data race (62 racecheck
// reports).  FSB200_RING_ATOMICS=0 rebuilds that protocol for A/B timing only.
#if FSB200_RING_ATOMICS
__device__ __forceinline__ int ld_acquire(int *p)
{
    int v;
    asm volatile("atom.acquire.cta.shared::cta.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(smem_addr(p)) : "memory");
    return v;
}
__device__ __forceinline__ int st_release(int *p, int v)   // returns the previous value (callers ignore it)
{
    int old;
    asm volatile("atom.release.cta.shared::cta.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(smem_addr(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ int cas_acq_rel(int *p, int expected, int desired)
{
    int old;
    asm volatile("atom.acq_rel.cta.shared::cta.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(smem_addr(p)), "r"(expected), "r"(desired) : "memory");
    return old;
}
__device__ __forceinline__ int add_acq_rel(int *p, int v)
{
    int old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_addr(p)), "r"(v) : "memory");
    return old;
}
#else
__device__ __forceinline__ int ld_acquire(int *p) { return *reinterpret_cast<volatile int *>(p); }
__device__ __forceinline__ int st_release(int *p, int v) { __threadfence_block(); *reinterpret_cast<volatile int *>(p) = v; return 0; }
__device__ __forceinline__ int cas_acq_rel(int *p, int expected, int desired) { return atomicCAS(p, expected, desired); }
__device__ __forceinline__ int add_acq_rel(int *p, int v) { __threadfence_block(); return atomicAdd(p, v); }
#endif
__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }   // global counters only

// executed by one full warp: stage fill number `fill` into slot sl
__device__ __noinline__ void fill_slot(const Workspace &ws, const IntegrateArgs &args, int n_items, Slot *sl,
                                          double4 *tile, uint64_t *bar, int fill, int lane)
{
    const int n_front = ws.counters[kCtrItems];
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(ws.counters + kCtrQueue, 1);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= n_items) {
            st_release(&sl->w[lane], 0);
            __syncwarp();
            if (lane == 0) {
                st_release(&sl->gathered, 0);
                st_release(&sl->dead, 1);
                st_release(&sl->claim, (fill & 0xffff) << 16);  // n_atoms = 0, next = 0
                mbar_arrive(bar);
            }
            return;
        }
        const Item it = ws.items[idx < n_front ? idx : ws.n - 1 - (idx - n_front)];   // surface cells first, then interior
        if (!(it.first < args.shard_end && it.first + it.count > args.shard_begin)) continue;  // another shard's
        const GridDesc &g = ws.grid[it.sid];
        const int cx = it.cell % g.dim[0], cy = (it.cell / g.dim[0]) % g.dim[1], cz = it.cell / (g.dim[0] * g.dim[1]);
        int b = 0, n = 0;
        if (lane < 9) cell_run(ws, g, cx, cy, cz, lane, &b, &n);
        int incl = n;                                      // inclusive prefix sum over lanes 0..8
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const int off = incl - n;
        const int total = __shfl_sync(kFull, incl, 8);
        const int off4 = __shfl_sync(kFull, off, 4), b4 = __shfl_sync(kFull, b, 4);
        const bool staged = total <= kTileCap;
        {   // the fill's description, one word per lane, written with ONE warp-wide atomic
            const int r = lane % 9;
            const int rb = __shfl_sync(kFull, b, r), rn = __shfl_sync(kFull, n, r), ro = __shfl_sync(kFull, off, r);
            const int word = lane < kWCount ? rb : lane < kWOff ? rn : lane < kWAtoms ? ro
                           : lane == kWAtoms ? it.count : lane == kWFirst ? it.first : lane == kWTotal ? total
                           : lane == kWStaged ? (staged ? 1 : 0) : off4 - b4;
            st_release(&sl->w[lane], word);
        }
        __syncwarp();                                      // all 32 words are written
        if (lane == 0) {
            st_release(&sl->gathered, 0);
            st_release(&sl->claim, ((fill & 0xffff) << 16) | (it.count << 8));   // publishes the payload above
        }
        __syncwarp();
        if (staged) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the old tile vs. async writes
                mbar_expect_tx(bar, (uint32_t)total * (uint32_t)sizeof(double4));
            }
            __syncwarp();
            if (lane < 9 && n > 0) tma_load_1d(tile + off, ws.atoms + b, (uint32_t)n * (uint32_t)sizeof(double4), bar);
        } else if (lane == 0) {
            mbar_arrive(bar);
        }
        return;
    }
}

template <int ALG, typename T>
__global__ void __launch_bounds__(kCtaThreads, 2) k_integrate(Workspace ws, IntegrateArgs args)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[kRingSlots];
    __shared__ Slot slots[kRingSlots];
    __shared__ int cur;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double4 *tiles = reinterpret_cast<double4 *>(smem);
    unsigned char *warp_mem = smem + (size_t)kRingSlots * kTileCap * sizeof(double4) + (size_t)warp * WarpLayout<ALG, T>::bytes(kNbCap);
    const WarpMem<ALG, T> wm(warp_mem, kNbCap);
    const int n_items = ws.counters[kCtrItems] + ws.counters[kCtrItemsBack];
    if (ws.counters[kCtrBadInput]) return;                 // the call fails anyway (uniform exit: nothing is armed yet)
    if (tid == 0) {
        cur = 0;
        for (int s = 0; s < kRingSlots; ++s) {
            slots[s].dead = 0;
            slots[s].gathered = 0;
            slots[s].claim = ((s - kRingSlots) & 0xffff) << 16;   // "one fill older than the first": not ready yet
            mbar_init(&full[s], 1);
        }
    }
    __syncthreads();
    if (warp == 0)
        for (int s = 0; s < kRingSlots; ++s) fill_slot(ws, args, n_items, &slots[s], tiles + (size_t)s * kTileCap, &full[s], s, lane);

    int idle_polls = 0, last_seen = -1;
    int n_certified = 0;
    for (;;) {
        // ---- claim one atom (lane 0 negotiates, the warp follows) ----------------------------------------
        // The slot's claim word names the fill it hosts.  Only when that is the fill `cur` points at do we
        // look at the mbarrier (whose parity is only meaningful for the phase of the hosted fill); a slot
        // that already hosts a LATER fill means fill `cur` is history (its last gather recycled the slot
        // before anybody noticed it was exhausted), an EARLIER one means the refill is still to come.
        int code = 0, f = 0, a = 0;                        // code 0: look again, 1: atom claimed, 2: all work done, 3: stalled
        if (lane == 0) {
            f = ld_acquire(&cur);
            Slot *sl = &slots[f % kRingSlots];
            if (ld_acquire(&sl->dead)) {
                bool all = true;
                for (int t = 0; t < kRingSlots; ++t) all = all && ld_acquire(&slots[t].dead) != 0;
                if (all) code = 2;
                else cas_acq_rel(&cur, f, f + 1);
            } else {
                const int w = ld_acquire(&sl->claim);
                const unsigned gen = (unsigned)w >> 16;
                if (gen == (unsigned)(f & 0xffff)) {
                    if (mbar_test(&full[f % kRingSlots], (uint32_t)(f / kRingSlots) & 1u)) {
                        a = w & 0xff;
                        // (a fill's atom count travels in the claim word: reading it from the slot instead let a
                        //  warp claim a non-existent atom of an exhausted one-atom fill whose slot was just being
                        //  recycled — caught by the self-check below on 1024-structure batches)
                        if (a >= ((w >> 8) & 0xff)) cas_acq_rel(&cur, f, f + 1);           // fill exhausted: open the next one
                        else if (cas_acq_rel(&sl->claim, w, w + 1) == w) code = 1;
                    }
                } else if (((gen - (unsigned)f) & 0xffffu) < 0x8000u) {
                    cas_acq_rel(&cur, f, f + 1);
                }
            }
            if (code == 0) {                               // nothing to do right now: back off, but never hang
                // Stall detector: counts POLLS during which the ring did not move (cur and the polled claim word
                // unchanged), not elapsed time — a context that is time-sliced out, single-stepped or run under
                // compute-sanitizer polls slowly but is not stalled.  2^26 polls of >= 64 ns each is >= 4 s of
                // continuous, motionless polling.
                const int seen = f * 31 + ld_acquire(&slots[f % kRingSlots].claim);
                if (seen != last_seen) {
                    last_seen = seen;
                    idle_polls = 0;
                } else if (++idle_polls > kStallPolls) {
                    if (atomicExch(ws.counters + kCtrStalled, 1) == 0) {   // first reporter leaves a post-mortem
                        int *dbg = ws.counters + 7;
                        dbg[0] = f;
                        dbg[1] = ld_acquire(&slots[0].claim); dbg[2] = ld_acquire(&slots[0].gathered);
                        dbg[3] = ld_acquire(&slots[0].w[kWAtoms]) * 2 + ld_acquire(&slots[0].dead);
                        dbg[4] = ld_acquire(&slots[1].claim); dbg[5] = ld_acquire(&slots[1].gathered);
                        dbg[6] = ld_acquire(&slots[1].w[kWAtoms]) * 2 + ld_acquire(&slots[1].dead);
                        dbg[7] = (int)mbar_test(&full[0], 0) + 2 * (int)mbar_test(&full[0], 1) + 4 * (int)mbar_test(&full[1], 0) + 8 * (int)mbar_test(&full[1], 1);
                        dbg[8] = ld_volatile(ws.counters + kCtrQueue);
                    }
                    code = 3;
                }
                __nanosleep(64);
            } else {
                idle_polls = 0;
            }
        }
        code = __shfl_sync(kFull, code, 0);
        if (code >= 2) break;
        if (code == 0) continue;
        f = __shfl_sync(kFull, f, 0);
        a = __shfl_sync(kFull, a, 0);
        const int s = f % kRingSlots;
        Slot *sl = &slots[s];
        double4 *tile = tiles + (size_t)s * kTileCap;
        mbar_wait(&full[s], (uint32_t)(f / kRingSlots) & 1u);     // every lane observes the completed phase (returns at once)
        const int pw = ld_acquire(&sl->w[lane]);                  // the fill's description: one word per lane, fields by shuffle
        const int n_atoms = __shfl_sync(kFull, pw, kWAtoms);
        const int pos = __shfl_sync(kFull, pw, kWFirst) + a;
        const bool mine = pos >= args.shard_begin && pos < args.shard_end;

        // ---- gather: the only use of the tile ----------------------------------------------------------------
        Self me;
        int nn = 0;
        if (mine) {
            const int begin4 = __shfl_sync(kFull, pw, kWBegin + 4);
            if (__shfl_sync(kFull, pw, kWStaged)) {
                const int self_idx = __shfl_sync(kFull, pw, kWSelf) + pos;
                me = load_self(tile[self_idx]);
                if (ALG == 0) {
                    nn = gather_run<ALG, T>(tile, __shfl_sync(kFull, pw, kWTotal), self_idx, 0, me, wm.recs, wm.cidx, 0, kNbCap, lane);
                } else {  // S&R keeps GLOBAL candidate positions (its exact re-check outlives the tile)
                    for (int r = 0; r < 9; ++r) {
                        const int rb = __shfl_sync(kFull, pw, kWBegin + r);
                        nn = gather_run<ALG, T>(tile + __shfl_sync(kFull, pw, kWOff + r), __shfl_sync(kFull, pw, kWCount + r),
                                                r == 4 ? pos - begin4 : -1, rb, me, wm.recs, wm.cidx, nn, kNbCap, lane);
                    }
                }
            } else {  // oversized neighbourhood: read the candidates straight from global memory
                me = load_self(ws.atoms[pos]);
                for (int r = 0; r < 9; ++r) {
                    const int rb = __shfl_sync(kFull, pw, kWBegin + r);
                    nn = gather_run<ALG, T>(ws.atoms + rb, __shfl_sync(kFull, pw, kWCount + r), r == 4 ? pos - begin4 : -1, rb, me,
                                            wm.recs, wm.cidx, nn, kNbCap, lane);
                }
            }
        }
        __syncwarp();
        int g = 0;
        if (lane == 0) {
            const int w_now = ld_acquire(&sl->claim);
            g = add_acq_rel(&sl->gathered, 1) + 1;         // release: my reads of the tile are done; acquire: the last one sees all
            // protocol self-check: the slot must still host my fill, and the count can never pass n_atoms
            if (((unsigned)w_now >> 16) != (unsigned)(f & 0xffff) || g > n_atoms) {
                if (atomicExch(ws.counters + kCtrStalled, 2) == 0) {
                    int *dbg = ws.counters + 7;
                    dbg[0] = f; dbg[1] = w_now; dbg[2] = g; dbg[3] = n_atoms * 2; dbg[4] = a; dbg[5] = ld_acquire(&sl->w[kWAtoms]); dbg[6] = pos;
                    dbg[7] = -1; dbg[8] = ld_acquire(&cur);
                }
            }
        }
        g = __shfl_sync(kFull, g, 0);
        if (g == n_atoms) fill_slot(ws, args, n_items, sl, tile, &full[s], f + kRingSlots, lane);  // last gather: recycle the slot

        // ---- integrate (no shared state besides the warp's own lists) -----------------------------------------
        if (mine) n_certified += finish_atom<ALG, T, true>(ws, args, wm, ws.atoms, me, nn, kNbCap, pos, true, lane, pw) ? 1 : 0;
        __syncwarp();
    }
    if (lane == 0 && n_certified) atomicAdd(ws.counters + kCtrCertified, n_certified);
}

// Atoms whose neighbour list exceeded kNbCap: one warp per atom, lists in global scratch.
template <int ALG, typename T>
__global__ void __launch_bounds__(128) k_overflow(Workspace ws, IntegrateArgs args, int n_overflow, int list_cap,
                                                  unsigned char *scratch)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned char *warp_mem = scratch + (size_t)gwarp * WarpLayout<ALG, T>::bytes(list_cap);
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(ws.counters + kCtrQueue2, 1);
        k = __shfl_sync(kFull, k, 0);
        if (k >= n_overflow) break;
        const int pos = ws.overflow[k];
        const int i = ws.perm[pos];
        const int sid = ws.n_struct == 1 ? 0 : find_structure(ws.offsets, ws.n_struct, i);
        const GridDesc &g = ws.grid[sid];
        const double4 me = ws.atoms[pos];
        int c[3];
        cell_coords(g, me.x, me.y, me.z, c);
        const Self s = load_self(me);
        const WarpMem<ALG, T> wm(warp_mem, list_cap);
        int nn = 0;
        for (int r = 0; r < 9; ++r) {
            int b, n;
            cell_run(ws, g, c[0], c[1], c[2], r, &b, &n);
            nn = gather_run<ALG, T>(ws.atoms + b, n, r == 4 ? pos - b : -1, b, s, wm.recs, wm.cidx, nn, list_cap, lane);
        }
        finish_atom<ALG, T, false>(ws, args, wm, ws.atoms, s, nn, list_cap, pos, false, lane);
        __syncwarp();
    }
}

// ---- split pipeline, second kernel: the slice loop and nothing else ---------------------------------------------------
// Persistent; a warp takes tasks from one queue.  The first n_redo tasks are the fp64 redos of the atoms with marginal
// slices (long, rare: started first so that they overlap with everything else); every other task is a (task record, chunk
// of slices) pair: the warp loads its K records straight into registers (coalesced 16-B loads from the pool, L2-resident),
// runs the chunk and stores the partial sum.  Nothing is synchronised between tasks; k_finish adds the partials of an atom
// IN CHUNK ORDER after this kernel (bit-reproducible).  Why a kernel of its own: the hot loop of the fused kernel shared
// the SM's instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB) with gather, certificate, ring and sort code of other
// warps; ncu (round 2): 29 % of all warp stall samples were `no_instruction`, 52 % inside the certificate, and every 2 KB
// of added per-atom code cost 10 % of kernel time.  Here every warp of the SM runs the same ~6 KB, and the unit of work is
// a few microseconds, so the tail of the launch is short without any ordering tricks.
constexpr int kSliceWarps = 8;

__device__ __noinline__ void redo_task(const Workspace &ws, const IntegrateArgs &args, unsigned char *rec, unsigned char *mem, int lane)
{
    Rec4<double> *recs = reinterpret_cast<Rec4<double> *>(mem);
    Arc<double> *arcs = reinterpret_cast<Arc<double> *>(mem + (size_t)kRedoMaxNeighbours * sizeof(Rec4<double>));
    TodoHeader *h = reinterpret_cast<TodoHeader *>(rec);
    const int ns = args.resolution, nn = h->nn;
    const TodoLayout L = todo_layout(ns, nn);
    const int *gpos = reinterpret_cast<const int *>(rec + L.gpos);
    const unsigned marginal = lane < (ns + 31) / 32 ? reinterpret_cast<const unsigned *>(rec + L.marginal)[lane] : 0u;
    for (int j = lane; j < nn; j += 32) {                  // fp64 records in the atom-local frame, as the gather forms them
        const double4 q = ws.atoms[gpos[j]];
        const double dx = q.x - h->x, dy = q.y - h->y, dz = q.z - h->z;
        Rec4<double> o;
        o.a = dz;
        o.b = q.w;
        o.c = sqrt(dx * dx + dy * dy);
        o.d = atan2(dy, dx) + 3.141592653589793;
        recs[j] = o;
    }
    __syncwarp();
    const double extra = lr_atom<double>(recs, arcs, nn, h->Ri, ns, lane, true, marginal);
    if (lane == 0) {
        h->extra = extra;
        atomicAdd(ws.counters + kCtrMarginal, 1);
    }
    __syncwarp();
}

// a task of k_slices: this lane's K records straight from the pool into registers, then the chunk
template <int K>
__device__ __forceinline__ double chunk_from_pool(const Rec4<float> *recs, KeyArc *arcs, const unsigned *marginal, int nn, double Ri,
                                                  int ns, int s0, int s1, int lane)
{
    Rec4<float> r[K];
    bool v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        v[k] = lane + 32 * k < nn;
        r[k] = recs[v[k] ? lane + 32 * k : 0];
    }
    return lr_slices_chunk<K>(r, v, arcs, marginal, nn, Ri, ns, s0, s1, lane);
}

__global__ void __launch_bounds__(kSliceWarps * 32, FSB200_SLICE_CTAS) k_slices(Workspace ws, IntegrateArgs args)
{
    __shared__ __align__(16) unsigned char mem_all[kSliceWarps][kRedoMaxNeighbours * (sizeof(Rec4<double>) + sizeof(Arc<double>))];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KeyArc *arcs = reinterpret_cast<KeyArc *>(mem_all[warp]);
    TodoCtl *ctl = args.todo_ctl;
    const int ns = args.resolution, S = lr_chunk_slices(ns), n_chunks = lr_n_chunks(ns);
    const int n_redo = ctl->n_redo;                                      // k_integrate has finished: final
    const long long n_tasks = (long long)n_redo + (long long)ctl->n_todo * n_chunks;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&ctl->task_head, 1);
        t = __shfl_sync(kFull, t, 0);
        if (t >= n_tasks) break;
        if (t < n_redo) {
            redo_task(ws, args, args.todo_pool + args.todo_list[args.todo_redo_base + t], mem_all[warp], lane);
            continue;
        }
        t -= n_redo;
        const int a = t / n_chunks, c = t - a * n_chunks;
        unsigned char *rec = args.todo_pool + args.todo_list[a];
        const TodoHeader *h = reinterpret_cast<const TodoHeader *>(rec);
        const int nn = h->nn;
        const double Ri = h->Ri;
        const TodoLayout L = todo_layout(ns, nn);
        const Rec4<float> *recs = reinterpret_cast<const Rec4<float> *>(rec + L.recs);
        const unsigned *marginal = h->has_marginal ? reinterpret_cast<const unsigned *>(rec + L.marginal) : nullptr;
        // One instantiation of the slice loop per number of record groups: in this kernel (unlike the fused one, where a
        // second instantiation cost more in instruction-cache misses than it saved) specialisation pays: C2 0.498 ms with
        // K = 3 only, 0.474 with K = 2 added for the atoms with <= 64 neighbours (most atoms with exposed surface), 0.466
        // with K = 1 as well.
        double part;
        const int s0 = c * S, s1 = min(c * S + S, ns);
#if FSB200_SLICES_K >= 3
        if (nn <= 32) part = chunk_from_pool<1>(recs, arcs, marginal, nn, Ri, ns, s0, s1, lane);
        else
#endif
#if FSB200_SLICES_K >= 2
        if (nn <= 64) part = chunk_from_pool<2>(recs, arcs, marginal, nn, Ri, ns, s0, s1, lane);
        else
#endif
            part = chunk_from_pool<kLrK>(recs, arcs, marginal, nn, Ri, ns, s0, s1, lane);
        if (lane == 0) reinterpret_cast<double *>(rec + L.partial)[c] = part;
        __syncwarp();
    }
}

// ---- split pipeline, last kernel: one thread per task record adds the chunks in order and writes the area ---------------
__global__ void __launch_bounds__(256) k_finish(Workspace ws, IntegrateArgs args)
{
    const int n_todo = args.todo_ctl->n_todo, ns = args.resolution, n_chunks = lr_n_chunks(ns);
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n_todo; a += gridDim.x * blockDim.x) {
        const unsigned char *rec = args.todo_pool + args.todo_list[a];
        const TodoHeader *h = reinterpret_cast<const TodoHeader *>(rec);
        const double *partial = reinterpret_cast<const double *>(rec + todo_layout(ns, h->nn).partial);
        double sum = 0.0;
        for (int k = 0; k < n_chunks; ++k) sum += partial[k];
        write_area(ws, args, h->pos, lr_area(h->Ri, ns, sum, h->extra));   // the fp64 redo returns delta * Ri * angle itself
    }
}

template <int ALG, typename T> size_t cta_smem_bytes()
{
    return (size_t)kRingSlots * kTileCap * sizeof(double4) + (size_t)kWarpsPerCta * WarpLayout<ALG, T>::bytes(kNbCap);
}

template <int ALG, typename T> int configure_and_occupancy(int device)
{
    const size_t smem = cta_smem_bytes<ALG, T>();
    cudaFuncSetAttribute(k_integrate<ALG, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_integrate<ALG, T>, kCtaThreads, smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (per_sm < 1) per_sm = 1;
    return per_sm * sms;
}

}  // namespace


int integrate_grid_ctas(int alg, int precision, int device)
{
    if (alg == 0) return precision == 0 ? configure_and_occupancy<0, float>(device) : configure_and_occupancy<0, double>(device);
    return precision == 0 ? configure_and_occupancy<1, float>(device) : configure_and_occupancy<1, double>(device);
}

size_t todo_pool_bytes_per_atom(int resolution) { return todo_layout(resolution, 96).bytes; }

int launch_integrate(const Workspace &ws, const IntegrateArgs &args, cudaStream_t stream)
{
    const int grid = args.grid_ctas;
    if (args.alg == 0 && args.precision == 0 && args.todo_pool != nullptr) {
        // split pipeline: prepare -> slices -> marginal slices; all three grids are persistent, sized for the device, and
        // read their queue lengths from device memory, so nothing comes back to the host in between
        static int slices_ctas = 0, sms = 0;
        if (slices_ctas == 0) {
            int dev = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_slices, kSliceWarps * 32, 0);
            slices_ctas = (per_sm < 1 ? 1 : per_sm) * sms;
        }
        cudaMemsetAsync(args.todo_ctl, 0, sizeof(TodoCtl), stream);
        k_integrate<0, float><<<grid, kCtaThreads, cta_smem_bytes<0, float>(), stream>>>(ws, args);
        k_slices<<<slices_ctas, kSliceWarps * 32, 0, stream>>>(ws, args);
        k_finish<<<sms, 256, 0, stream>>>(ws, args);
        return 3;
    }
    if (args.alg == 0) {
        if (args.precision == 0)
            k_integrate<0, float><<<grid, kCtaThreads, cta_smem_bytes<0, float>(), stream>>>(ws, args);
        else
            k_integrate<0, double><<<grid, kCtaThreads, cta_smem_bytes<0, double>(), stream>>>(ws, args);
    } else {
        if (args.precision == 0)
            k_integrate<1, float><<<grid, kCtaThreads, cta_smem_bytes<1, float>(), stream>>>(ws, args);
        else
            k_integrate<1, double><<<grid, kCtaThreads, cta_smem_bytes<1, double>(), stream>>>(ws, args);
    }
    return 1;
}

int overflow_warps(int n_overflow)
{
    const int want = n_overflow < 1 ? 1 : n_overflow;
    return want < 1024 ? ((want + 3) / 4) * 4 : 1024;  // 4 warps per CTA
}

size_t overflow_scratch_bytes(int n_warps, int list_cap, int precision)
{
    // sized for the larger of the two algorithms so one buffer serves both
    const size_t a = precision == 0 ? WarpLayout<0, float>::bytes(list_cap) : WarpLayout<0, double>::bytes(list_cap);
    const size_t b = precision == 0 ? WarpLayout<1, float>::bytes(list_cap) : WarpLayout<1, double>::bytes(list_cap);
    return (size_t)n_warps * (a > b ? a : b);
}

int launch_overflow(const Workspace &ws, const IntegrateArgs &args, int n_overflow, int list_cap, void *scratch,
                    cudaStream_t stream)
{
    const int warps = overflow_warps(n_overflow);
    const int grid = warps / 4;
    unsigned char *sc = static_cast<unsigned char *>(scratch);
    if (args.alg == 0) {
        if (args.precision == 0) k_overflow<0, float><<<grid, 128, 0, stream>>>(ws, args, n_overflow, list_cap, sc);
        else k_overflow<0, double><<<grid, 128, 0, stream>>>(ws, args, n_overflow, list_cap, sc);
    } else {
        if (args.precision == 0) k_overflow<1, float><<<grid, 128, 0, stream>>>(ws, args, n_overflow, list_cap, sc);
        else k_overflow<1, double><<<grid, 128, 0, stream>>>(ws, args, n_overflow, list_cap, sc);
    }
    return 1;
}

}  // namespace fsb200
