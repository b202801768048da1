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
 *   - re-entrant: host-pointer calls borrow a context from a pool; no state survives a call other
 *     than cached device scratch and pinned staging (bounded: see fsb200_trim());
 *   - environment: FSB200_PRECISION=fp64 (all-fp64 kernels), FREESASA_B200_DEVICE=<k> (device of the context-free entry
 *     points when the caller has not chosen one), FREESASA_B200_VERBOSE=1 (host layer prints one timing line per call).
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
    int n_marginal;          /* L&R fp32: atoms with at least one near-tangent slice redone in fp64 */
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

/* ---- several GPUs behind one call (single process, one host thread per device, no NCCL) ------------ */
/* What SURVEY.md section 8(b) calls fsb200_batch(..., n_devices) / fsb200_lr_multi(..., n_devices): the multi-GPU form
 * of the entry points above for a C caller of freesasa_calc_structure() (src/freesasa.c:144-153) or of the CLI's loop
 * over structures (src/main.cc:334-362).  n_devices <= 0 means every visible sm_100 device.
 *   n_struct == 1  one structure: inputs replicated (each device uploads 1/N over its own PCIe link, the rest is an
 *                  all-gather of the inputs over NVLink peer copies), outputs partitioned (each device integrates its
 *                  share of the cell-sorted atoms and stores every area, over NVLink, straight into the result slice of the
 *                  device that owns that part of the caller's array; each device downloads its slice over its own PCIe
 *                  link).  At most 8 devices.  Whole-structure SASA: every atom sees all its neighbours (NOT the per-chain quantity of
 *                  --separate-chains, src/structure.c:955-1081).  Results are bit-identical to the one-device call.
 *   n_struct  > 1  independent structures dealt to the devices by longest-processing-time on their atom counts. */
#define FSB200_MAX_DEVICES 16
int fsb200_calc_multi(int alg, int n_struct, const int *n_atoms, const double *const *xyz,
                      const double *const *radii, double *const *sasa, double probe, int resolution,
                      int n_devices);
int fsb200_lr_multi(double *sasa, const double *xyz, const double *radii, int n, double probe,
                    int n_slices, int n_devices);
int fsb200_sr_multi(double *sasa, const double *xyz, const double *radii, int n, double probe,
                    int n_points, int n_devices);
/* Host-side timing of the last fsb200_calc_multi() of the process (benchmark bookkeeping). */
typedef struct fsb200_multi_stats {
    int n_devices, n_atoms, n_structures, n_certified;
    float upload_ms;   /* one structure: wall time until the input all-gather is enqueued on device 0 */
    float compute_ms;  /* until every device has finished its share */
    float download_ms; /* device 0 -> host */
    float total_ms;    /* the whole call */
    float integrate_ms[FSB200_MAX_DEVICES]; /* one structure: integration kernel per device (CUDA events) */
    float device_ms[FSB200_MAX_DEVICES];    /* one structure: cell list + kernel per device; batch: wall time of the device's share */
} fsb200_multi_stats;
int fsb200_get_multi_stats(fsb200_multi_stats *out);
/* Host memory policy of the context-free entry points: idle pooled contexts keep their device scratch, at most 4 per
 * device, and give back pinned staging above 256 MiB when a call returns.  fsb200_trim() destroys every idle pooled
 * context (device scratch, pinned staging, stream) and returns how many there were. */
int fsb200_trim(void);

/* Statistics of the last context-free call (fsb200_lr / _sr / _calc_batch) made by the calling thread. */
int fsb200_last_stats(fsb200_stats *out);

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
/* The same call in two halves, so that a collective (or anything else) can be queued on the stream BEHIND the
 * integration kernel and BEFORE the one host synchronisation of the call:
 *     fsb200_ctx_calc_device_async(...)   validates and enqueues everything, returns without waiting
 *     ... ncclAllGather(d_sasa ...) on the same stream, result download, peer signalling ...
 *     fsb200_ctx_finish(ctx)              waits, reads the status words; FSB200_SUCCESS, FSB200_FAIL, or
 *                                         FSB200_SECOND_PASS: atoms with more than 160 neighbours (never seen in
 *                                         proteins) were completed by a second kernel AFTER the work queued in between,
 *                                         so that work (the collective) has to be issued again.
 * No other call may use the context between the two halves. */
#define FSB200_SECOND_PASS 1
int fsb200_ctx_calc_device_async(fsb200_ctx *ctx, int alg, const double *d_xyz, const double *d_radii,
                                 int n_total, int n_struct, const int *offsets, double probe,
                                 int resolution, int shard_index, int shard_count, double *d_sasa,
                                 void *stream);
int fsb200_ctx_finish(fsb200_ctx *ctx);
/* Fused all-gather: every area computed by later device-resident calls on ctx is ALSO stored, at the same index, into
 * these n_peers buffers (n_total doubles each), which normally live on OTHER GPUs — peer memory of the same process
 * (cudaDeviceEnablePeerAccess) or buffers opened from another process with fsb200_ipc_open().  With peers set, a sharded
 * call writes the caller's atom order (not the sorted order), so after all shards have finished every buffer holds the
 * complete result and no unpermute / NCCL all-gather is needed.  n_peers = 0 switches it off. */
int fsb200_ctx_set_peer_outputs(fsb200_ctx *ctx, int n_peers, double *const *d_peer_sasa);
/* With zero skipping on, an area of exactly 0 (nine atoms in ten of a large structure: everything the buried-atom certificate
 * settles) is NOT stored into the peer buffers: their owners zero them before the step (cudaMemsetAsync ahead of their
 * barrier signal), which turns 8 B per atom and peer into 8 B per EXPOSED atom and peer. */
int fsb200_ctx_set_peer_zero_skipping(fsb200_ctx *ctx, int on);
/* CUDA IPC plumbing for the one-process-per-GPU case: allocate a device buffer that other processes can map, export its
 * 64-byte handle, map a peer's buffer, unmap it. */
int fsb200_ipc_alloc(int device, unsigned long long bytes, void **d_ptr, unsigned char handle[64]);
int fsb200_ipc_open(int device, const unsigned char handle[64], void **d_ptr);
int fsb200_ipc_close(int device, void *d_ptr);
int fsb200_ipc_free(int device, void *d_ptr);
/* Barrier between the GPUs of world ranks WITHOUT NCCL and without the host: enqueues a one-warp kernel on the stream that
 * stores a new epoch into slot `rank` of every rank's flag array (d_flags[r]: `world` ints on rank r's GPU, zeroed, peer-
 * mapped) and waits until all slots of its own array carry it.  Queued behind fsb200_ctx_calc_device_async() it completes
 * when every rank's integration kernel — and with it every peer store into this rank's buffer — is done.  Bounded wait:
 * fsb200_ctx_peer_barrier_status() (after a stream synchronisation) reports a peer that never arrived. world <= 9. */
int fsb200_ctx_peer_barrier(fsb200_ctx *ctx, int rank, int world, int *const *d_flags, void *stream);
int fsb200_ctx_peer_barrier_status(fsb200_ctx *ctx);
/* Identifies the last pipeline run on ctx (stamps the permutation fsb200_ctx_unpermute() would use). */
unsigned long long fsb200_ctx_generation(const fsb200_ctx *ctx);
int fsb200_shard_begin(int n_total, int shard_index, int shard_count);
int fsb200_shard_end(int n_total, int shard_index, int shard_count);
/* d_out[perm[p]] = d_sorted[p] using the permutation of the last calc_device call on ctx.  Fails if any other call ran
 * on the context since (it would have replaced the permutation). */
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
