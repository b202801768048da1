/* freesasa_b200/csrc/host_shim.c — C host layer: the reference's hot-path entry points on top of
 * the fsb200 C ABI (include/fsb200.h).  See include/freesasa_b200_host.h.
 *
 * What stays identical to the reference (so existing callers and tests keep working):
 *   - freesasa_calc() allocates the result with malloc, dispatches on parameters->alg, frees the
 *     result and returns NULL when the engine reports FREESASA_FAIL, sums `total` serially in atom
 *     order and copies the parameters (src/freesasa.c:76-120);
 *   - freesasa_lee_richards()/freesasa_shrake_rupley() validate like src/sasa_lr.c:169-193 and
 *     src/sasa_sr.c:173-200: more than 16 threads -> FAIL, resolution <= 0 -> FAIL, zero atoms ->
 *     WARN, more threads than atoms -> warning and carry on;
 *   - messages go to the error stream selected with freesasa_set_err_out() and obey the verbosity.
 * What changes: the numeric work is one call into the GPU engine; n_threads is validated but does
 * not drive anything.  There is no CPU fallback: if the engine fails, the message says why.
 */
#include "host_internal.h"
#include "fsb200.h"

#include <assert.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_THREADS 16 /* MAX_LR_THREADS / MAX_SR_THREADS, src/sasa_lr.c:17, src/sasa_sr.c:16 */

const int FREESASA_DEF_NUMBER_THREADS = 2; /* the threaded default, src/freesasa.c:31-36 */
const freesasa_parameters freesasa_default_parameters = {
    FREESASA_DEF_ALGORITHM, FREESASA_DEF_PROBE_RADIUS, FREESASA_DEF_SR_N, FREESASA_DEF_LR_N, 2};

static freesasa_verbosity verbosity = FREESASA_V_NORMAL;
static FILE *errlog = NULL;
static const char *prog = "freesasa";

int freesasa_set_verbosity(freesasa_verbosity v)
{
    if (v == FREESASA_V_NORMAL || v == FREESASA_V_NOWARNINGS || v == FREESASA_V_SILENT || v == FREESASA_V_DEBUG) {
        verbosity = v;
        return FREESASA_SUCCESS;
    }
    return FREESASA_WARN;
}
freesasa_verbosity freesasa_get_verbosity(void) { return verbosity; }
void freesasa_set_err_out(FILE *fp)
{
    assert(fp);
    errlog = fp;
}
FILE *freesasa_get_err_out(void) { return errlog; }

/* Messages of worker threads (parallel PDB reading) are captured per worker and replayed in input order, so that what
 * reaches the error stream is exactly what a serial run prints. */
__thread struct fsb_capture *fsb_capture_current = NULL;

void fsb_capture_flush(struct fsb_capture *c)
{
    if (c->len) {
        FILE *fp = errlog ? errlog : stderr;
        fwrite(c->text, 1, c->len, fp);
        fflush(fp);
    }
    free(c->text);
    c->text = NULL;
    c->len = c->cap = 0;
}

int fsb_report(int code, const char *where, int line, const char *fmt, ...)
{
    va_list ap;
    FILE *fp = errlog ? errlog : stderr;
    char buf[1024];
    int n;
    if (verbosity == FREESASA_V_SILENT) return code;
    if (code == FREESASA_WARN && verbosity == FREESASA_V_NOWARNINGS) return code;
    if (where)
        n = snprintf(buf, sizeof buf, "%s:%s:%d: %s: ", prog, where, line, code == FREESASA_WARN ? "warning" : "error");
    else
        n = snprintf(buf, sizeof buf, "%s: %s: ", prog, code == FREESASA_WARN ? "warning" : "error");
    va_start(ap, fmt);
    n += vsnprintf(buf + n, sizeof buf - (size_t)n, fmt, ap);
    va_end(ap);
    if (n > (int)sizeof buf - 2) n = (int)sizeof buf - 2;
    buf[n++] = '\n';
    if (fsb_capture_current) {
        struct fsb_capture *c = fsb_capture_current;
        if (c->len + (size_t)n > c->cap) {
            const size_t cap = 2 * c->cap + (size_t)n + 256;
            char *p = realloc(c->text, cap);
            if (!p) return code;
            c->text = p;
            c->cap = cap;
        }
        memcpy(c->text + c->len, buf, (size_t)n);
        c->len += (size_t)n;
        return code;
    }
    fwrite(buf, 1, (size_t)n, fp);
    fflush(fp);
    return code;
}

/* shared validation of src/sasa_lr.c:169-193 / src/sasa_sr.c:173-200; returns 1 if the caller should
 * return *rc immediately */
static int validate(const char *alg, const char *func, int n_atoms, int n_threads, int resolution, int *rc)
{
    *rc = FREESASA_SUCCESS;
    if (n_threads > MAX_THREADS) {
        *rc = FAIL_MSG("%s does not support more than %d threads", alg, MAX_THREADS);
        return 1;
    }
    if (resolution <= 0) {
        *rc = FAIL_MSG("%d is an invalid resolution in %s, must be > 0", resolution, alg);
        return 1;
    }
    if (n_atoms == 0) {
        *rc = WARN_MSG("in %s(): empty coordinates", func);
        return 1;
    }
    if (n_threads > n_atoms)
        WARN_MSG("no sense in having more threads than atoms, only using %d threads", n_atoms);
    return 0;
}

/* FREESASA_B200_DEVICES=<n|all>: the drop-in entry points spread ONE call over n GPUs (fsb200_calc_multi: a structure of a
 * million atoms is replicated and its outputs partitioned; a batch is dealt structure by structure).  Default: one GPU. */
static int multi_devices(void)
{
    static int cached = -2;
    if (cached == -2) {
        const char *e = getenv("FREESASA_B200_DEVICES");
        if (!e || !*e) cached = 1;
        else if (strcmp(e, "all") == 0 || strcmp(e, "ALL") == 0) cached = 0;
        else cached = atoi(e) > 0 ? atoi(e) : 1;
    }
    return cached;
}

/* FREESASA_B200_VERBOSE=1: one line per engine call on the error stream (freesasa_set_err_out, src/util.c:131-141) with
 * what the device did — the engine itself never prints. */
static void verbose_line(const char *what)
{
    static int on = -1;
    fsb200_stats s;
    FILE *fp;
    if (on < 0) {
        const char *e = getenv("FREESASA_B200_VERBOSE");
        on = e && *e && strcmp(e, "0") != 0;
    }
    if (!on || verbosity == FREESASA_V_SILENT) return;
    fp = errlog ? errlog : stderr;
    if (multi_devices() != 1) {
        fsb200_multi_stats m;
        if (fsb200_get_multi_stats(&m) == FSB200_SUCCESS && m.n_devices > 1) {
            fprintf(fp, "%s: B200 engine: %s: %d atoms in %d structure(s) on %d GPUs: upload %.3f ms, compute %.3f ms, download %.3f ms, "
                        "total %.3f ms\n", prog, what, m.n_atoms, m.n_structures, m.n_devices, m.upload_ms, m.compute_ms, m.download_ms,
                    m.total_ms);
            fflush(fp);
            return;
        }
    }
    if (fsb200_last_stats(&s) != FSB200_SUCCESS) return;
    fprintf(fp, "%s: B200 engine: %s: %d atoms in %d structure(s): device %.3f ms (integration kernel %.3f ms, %d launches), "
                "%d atoms proved buried, %d work items, %d large neighbourhoods; host: staging %.3f ms, whole call %.3f ms\n",
            prog, what, s.n_atoms, s.n_structures, s.device_ms, s.integrate_ms, s.kernel_launches, s.n_certified, s.n_items,
            s.n_overflow, s.host_stage_ms, s.host_total_ms);
    fflush(fp);
}

int freesasa_lee_richards(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param)
{
    int rc;
    assert(sasa);
    assert(c);
    assert(radii);
    if (param == NULL) param = &freesasa_default_parameters;
    if (validate("L&R", __func__, c->n, param->n_threads, param->lee_richards_n_slices, &rc)) return rc;
    if ((multi_devices() == 1 ? fsb200_lr(sasa, c->xyz, radii, c->n, param->probe_radius, param->lee_richards_n_slices)
                              : fsb200_lr_multi(sasa, c->xyz, radii, c->n, param->probe_radius, param->lee_richards_n_slices,
                                                multi_devices())) != FSB200_SUCCESS)
        return FAIL_MSG("B200 engine: %s", fsb200_last_error());
    verbose_line("Lee & Richards");
    return FREESASA_SUCCESS;
}

int freesasa_shrake_rupley(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param)
{
    int rc;
    assert(sasa);
    assert(c);
    assert(radii);
    if (param == NULL) param = &freesasa_default_parameters;
    if (validate("S&R", __func__, c->n, param->n_threads, param->shrake_rupley_n_points, &rc)) return rc;
    if ((multi_devices() == 1 ? fsb200_sr(sasa, c->xyz, radii, c->n, param->probe_radius, param->shrake_rupley_n_points)
                              : fsb200_sr_multi(sasa, c->xyz, radii, c->n, param->probe_radius, param->shrake_rupley_n_points,
                                                multi_devices())) != FSB200_SUCCESS)
        return FAIL_MSG("B200 engine: %s", fsb200_last_error());
    verbose_line("Shrake & Rupley");
    return FREESASA_SUCCESS;
}

static freesasa_result *result_new(int n)
{
    freesasa_result *r = malloc(sizeof *r);
    if (r == NULL) {
        FAIL_MSG("Out of memory");
        return NULL;
    }
    r->sasa = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (r->sasa == NULL) {
        FAIL_MSG("Out of memory");
        free(r);
        return NULL;
    }
    r->n_atoms = n;
    r->total = 0;
    return r;
}

void freesasa_result_free(freesasa_result *r)
{
    if (r) {
        free(r->sasa);
        free(r);
    }
}

static void finish_result(freesasa_result *r, const freesasa_parameters *p)
{
    int i;
    r->total = 0; /* serial sum in atom order, as src/freesasa.c:113-116 */
    for (i = 0; i < r->n_atoms; ++i) r->total += r->sasa[i];
    r->parameters = *p;
}

freesasa_result *freesasa_calc(const coord_t *c, const double *radii, const freesasa_parameters *parameters)
{
    freesasa_result *result;
    int ret = FREESASA_SUCCESS;
    assert(c);
    assert(radii);
    result = result_new(c->n);
    if (result == NULL) {
        FAIL_MSG("%s", "");
        return NULL;
    }
    if (parameters == NULL) parameters = &freesasa_default_parameters;
    switch (parameters->alg) {
    case FREESASA_SHRAKE_RUPLEY:
        ret = freesasa_shrake_rupley(result->sasa, c, radii, parameters);
        break;
    case FREESASA_LEE_RICHARDS:
        ret = freesasa_lee_richards(result->sasa, c, radii, parameters);
        break;
    default:
        assert(0);
        break;
    }
    if (ret == FREESASA_FAIL) {
        freesasa_result_free(result);
        return NULL;
    }
    finish_result(result, parameters);
    return result;
}

freesasa_result *freesasa_calc_coord(const double *xyz, const double *radii, int n, const freesasa_parameters *parameters)
{
    coord_t view; /* linked view of the caller's array: zero-copy, read-only (src/coord.c:72-88) */
    freesasa_result *result;
    assert(xyz);
    assert(radii);
    assert(n > 0);
    view.n = n;
    view.is_linked = 1;
    view.xyz = (double *)xyz;
    result = freesasa_calc(&view, radii, parameters);
    if (result == NULL) FAIL_MSG("%s", "");
    return result;
}

int freesasa_calc_coord_batch(int n_struct, const double *const *xyz, const double *const *radii, const int *n_atoms,
                              const freesasa_parameters *parameters, freesasa_result **results)
{
    int k, rc, resolution;
    double **sasa;
    if (parameters == NULL) parameters = &freesasa_default_parameters;
    if (n_struct <= 0 || !xyz || !radii || !n_atoms || !results) return FAIL_MSG("invalid batch arguments");
    resolution = parameters->alg == FREESASA_LEE_RICHARDS ? parameters->lee_richards_n_slices
                                                          : parameters->shrake_rupley_n_points;
    for (k = 0; k < n_struct; ++k) {
        results[k] = NULL;
        if (n_atoms[k] <= 0) return FAIL_MSG("structure %d has no atoms", k);
        if (validate(parameters->alg == FREESASA_LEE_RICHARDS ? "L&R" : "S&R", __func__, n_atoms[k],
                     k == 0 ? parameters->n_threads : 1, resolution, &rc))
            return rc == FREESASA_WARN ? FREESASA_FAIL : rc;
    }
    sasa = malloc(sizeof(double *) * (size_t)n_struct);
    if (!sasa) return FAIL_MSG("Out of memory");
    for (k = 0; k < n_struct; ++k) {
        results[k] = result_new(n_atoms[k]);
        if (!results[k]) goto cleanup;
        sasa[k] = results[k]->sasa;
    }
    if ((multi_devices() == 1 ? fsb200_calc_batch((int)parameters->alg, n_struct, n_atoms, xyz, radii, sasa,
                                                  parameters->probe_radius, resolution)
                              : fsb200_calc_multi((int)parameters->alg, n_struct, n_atoms, xyz, radii, sasa,
                                                  parameters->probe_radius, resolution, multi_devices())) != FSB200_SUCCESS) {
        FAIL_MSG("B200 engine: %s", fsb200_last_error());
        goto cleanup;
    }
    for (k = 0; k < n_struct; ++k) finish_result(results[k], parameters);
    free(sasa);
    verbose_line("batch");
    return FREESASA_SUCCESS;
cleanup:
    for (k = 0; k < n_struct; ++k) {
        freesasa_result_free(results[k]);
        results[k] = NULL;
    }
    free(sasa);
    return FREESASA_FAIL;
}
