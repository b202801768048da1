/* include/fsb200.h — C ABI of libfsb200.so, the B200-native SASA engine.
 *
 * This is the drop-in boundary for the FreeSASA hot path.  The reference has no plugin API; its
 * seam is the dispatch inside freesasa_calc() (reference src/freesasa.c:97-107) to
 *
 *     int freesasa_lee_richards (double *sasa, const coord_t *c, const double *radii,
 *                                const freesasa_parameters *param);   // src/freesasa_internal.h:100-103
 *     int freesasa_shrake_rupley(double *sasa, const coord_t *c, const double *radii,
 *                                const freesasa_parameters *param);   // src/freesasa_internal.h:74-77
 *
 * which in turn own the neighbour list (src/nb.c:524 freesasa_nb_new).  fsb200_lr()/fsb200_sr()
 * replace the bodies of those two functions (see INTEGRATION.md for the few-line shim that the
 * reference's src/sasa_lr.c / src/sasa_sr.c become); everything else in the reference — parsing,
 * classifiers, result tree, output — keeps calling freesasa_calc() unchanged.
 *
 * Conventions (all entry points):
 *   - plain C types only; `xyz` is the reference's AoS layout x1,y1,z1,x2,... (src/coord.h:26-38),
 *     `radii` are van der Waals radii WITHOUT the probe (the engine adds it, as
 *     src/sasa_lr.c:135-138 / src/sasa_sr.c:143-147 do), `sasa` receives one area per atom in Å^2;
 *   - return value is FSB200_SUCCESS (0) or FSB200_FAIL (-1), the same numeric values as
 *     FREESASA_SUCCESS / FREESASA_FAIL (src/freesasa.h:151-155); the engine never prints — the
 *     caller fetches the message with fsb200_last_error() and reports it through its own
 *     fail_msg() (src/freesasa_internal.h:29-32);
 *   - there is NO CPU fallback: without a usable sm_100 device every compute call fails loudly;
 *   - re-entrant: host-pointer calls use a per-thread context; no state survives a call other
 *     than cached device scratch.
 */
#ifndef FSB200_H
#define FSB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FSB200_SUCCESS 0
#define FSB200_FAIL (-1)

/* same numeric values as enum freesasa_algorithm (src/freesasa.h:89-92) */
#define FSB200_LEE_RICHARDS 0
#define FSB200_SHRAKE_RUPLEY 1

/* arithmetic of the integration kernels */
#define FSB200_FP32 0 /* default: fp32 in the atom-local frame, fp64 neighbour test + fp64 tie re-check */
#define FSB200_FP64 1 /* everything in fp64 (validation / maximum fidelity); also selected for every new context by the
                       * environment variable FSB200_PRECISION=fp64, which is how a caller of the drop-in entry points
                       * (no precision argument) asks for it */

typedef struct fsb200_ctx fsb200_ctx;

/* What the last call on a context did (diagnostics, benchmark bookkeeping). */
typedef struct fsb200_stats {
    int n_atoms;             /* atoms in the call */
    int n_structures;        /* independent structures in the call */
    int n_items;             /* work items (cell chunks) integrated */
    int n_overflow;          /* atoms that took the large-neighbourhood path */
    int max_neighbours;      /* largest neighbour count seen by the overflow path (0 if unused) */
    int n_certified;         /* atoms proved completely buried (area exactly 0) without being integrated */
    int kernel_launches;     /* kernels launched by this call */
    float device_ms;         /* device time of the call, CUDA events on the call's stream */
    float integrate_ms;      /* device time of the integration kernel alone */
    float host_stage_ms;     /* host-pointer calls: wall time spent staging + enqueueing the upload */
    float host_total_ms;     /* host-pointer calls: wall time of the whole call */
} fsb200_stats;

/* ---- library / device -------------------------------------------------------------------- */
int fsb200_available(void);           /* 1 if at least one compute-capability-10.x device is usable */
int fsb200_device_count(void);
const char *fsb200_last_error(void);  /* thread-local, never NULL */
const char *fsb200_version(void);
unsigned long long fsb200_launch_count(void); /* kernels launched by this process so far */

/* ---- drop-in entry points: host buffers in, host buffer out -------------------------------- */
/* Replaces the body of freesasa_lee_richards() (src/sasa_lr.c:156-216) after its own parameter
 * validation; n_slices = param->lee_richards_n_slices, probe = param->probe_radius. */
int fsb200_lr(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_slices);
/* Replaces the body of freesasa_shrake_rupley() (src/sasa_sr.c:168-224). */
int fsb200_sr(double *sasa, const double *xyz, const double *radii, int n, double probe, int n_points);
/* Many independent structures in one device pass (what the CLI's serial loop over structures,
 * src/main.cc:334-362, would hand over).  sasa[k] receives n_atoms[k] doubles. */
int fsb200_calc_batch(int alg, int n_struct, const int *n_atoms, const double *const *xyz,
                      const double *const *radii, double *const *sasa, double probe, int resolution);

/* ---- explicit contexts (one per device / stream user) ----------------------------------------- */
fsb200_ctx *fsb200_ctx_create(int device); /* NULL on failure */
void fsb200_ctx_destroy(fsb200_ctx *ctx);
int fsb200_ctx_set_precision(fsb200_ctx *ctx, int precision);
/* The buried-atom certificate (fp32 mode) proves "no exposed surface" for most interior atoms and skips their
 * integration; it never changes a result.  On by default; off is for ablation and tests. */
int fsb200_ctx_set_certificate(fsb200_ctx *ctx, int on);
int fsb200_ctx_calc(fsb200_ctx *ctx, int alg, double *sasa, const double *xyz, const double *radii,
                    int n, double probe, int resolution);
int fsb200_ctx_calc_batch(fsb200_ctx *ctx, int alg, int n_struct, const int *n_atoms,
                          const double *const *xyz, const double *const *radii,
                          double *const *sasa, double probe, int resolution);
int fsb200_ctx_stats(const fsb200_ctx *ctx, fsb200_stats *out);

/* ---- device-resident entry point ------------------------------------------------------------- */
/* Inputs and output already in device memory of ctx's device:
 *   d_xyz    3*n_total doubles, structures concatenated
 *   d_radii  n_total doubles
 *   offsets  host array of n_struct+1 atom offsets (offsets[0]=0, offsets[n_struct]=n_total);
 *            NULL means one structure
 *   shard_index/shard_count  this caller integrates only its share (contiguous range of the
 *            cell-sorted atom order) of ONE replicated problem; 0/1 = everything
 *   d_sasa   n_total doubles; with shard_count==1 written in the caller's atom order.  With
 *            shard_count>1 the owned range [fsb200_shard_begin, fsb200_shard_end) of the SORTED
 *            order is written to d_sasa[sorted position]; after gathering all shards, call
 *            fsb200_ctx_unpermute() to obtain the caller's order.
 *   stream   a cudaStream_t (as void*), NULL = the context's own stream
 * The call enqueues all work on the stream and then synchronises it (status words are read back).
 */
int fsb200_ctx_calc_device(fsb200_ctx *ctx, int alg, const double *d_xyz, const double *d_radii,
                           int n_total, int n_struct, const int *offsets, double probe,
                           int resolution, int shard_index, int shard_count, double *d_sasa,
                           void *stream);
int fsb200_shard_begin(int n_total, int shard_index, int shard_count);
int fsb200_shard_end(int n_total, int shard_index, int shard_count);
/* d_out[perm[p]] = d_sorted[p] using the permutation of the last calc_device call on ctx. */
int fsb200_ctx_unpermute(fsb200_ctx *ctx, const double *d_sorted, double *d_out, int n_total,
                         void *stream);

/* ---- test hook ------------------------------------------------------------------------------ */
/* Per-atom neighbour counts |{j != i : |x_i-x_j|^2 < (R_i+R_j)^2}| as the engine's cell list sees
 * them (parity check for the src/nb.c row of the scope table).  Bit 30 of a count is set when the atom was
 * settled by the buried-atom certificate. */
int fsb200_ctx_neighbour_counts(fsb200_ctx *ctx, int *counts, const double *xyz, const double *radii,
                                int n, double probe);
/* The n Shrake-Rupley unit test points exactly as the engine hands them to the device (3n doubles): the
 * reference's golden spiral (src/sasa_sr.c:56-90), bit-identical values, in the engine's patch order.
 * Pure host code: works without a GPU. */
int fsb200_test_points(int n_points, double *out);
/* The 128 probe directions of the buried-atom certificate (384 doubles: 64 unit vectors, then their negatives) exactly as
 * the device uses them, so that a test can measure their covering radius.  Pure host code. */
int fsb200_cert_directions(double *out);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
