/* include/freesasa_b200_host.h — the hot-path subset of the FreeSASA C API, B200-backed.
 *
 * libfreesasa_b200_host.so exports, under the reference's own names and with ABI-identical
 * structs, exactly the functions that sit between a FreeSASA caller and the numeric hot path:
 *
 *   freesasa_calc_coord()      reference src/freesasa.h:475-479, src/freesasa.c:122-142
 *   freesasa_calc()            reference src/freesasa_internal.h:119-122, src/freesasa.c:76-120
 *   freesasa_lee_richards()    reference src/freesasa_internal.h:100-103, src/sasa_lr.c:156-216
 *   freesasa_shrake_rupley()   reference src/freesasa_internal.h:74-77,   src/sasa_sr.c:168-224
 *   freesasa_result_free()     reference src/freesasa.h:530, src/freesasa.c:68-74
 *   freesasa_default_parameters, freesasa_set_verbosity/get_verbosity, freesasa_set_err_out
 *
 * Same argument meaning, same return codes, same parameter validation and the same error/warning
 * text conventions ("freesasa: error: ...").  A program written against freesasa.h that only uses
 * these calls can link this library instead of libfreesasa; a full FreeSASA build swaps its
 * src/sasa_lr.c + src/sasa_sr.c + src/nb.c for the shim in INTEGRATION.md and keeps everything else.
 *
 * Additive (not in the reference): freesasa_calc_coord_batch().
 */
#ifndef FREESASA_B200_HOST_H
#define FREESASA_B200_HOST_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/freesasa.h:89-92 */
enum freesasa_algorithm { FREESASA_LEE_RICHARDS, FREESASA_SHRAKE_RUPLEY };
/* reference src/freesasa.h:104-109 */
enum freesasa_verbosity { FREESASA_V_NORMAL, FREESASA_V_NOWARNINGS, FREESASA_V_SILENT, FREESASA_V_DEBUG };
/* reference src/freesasa.h:151-155 */
enum freesasa_error_codes { FREESASA_SUCCESS = 0, FREESASA_FAIL = -1, FREESASA_WARN = -2 };

#ifndef __cplusplus
typedef enum freesasa_algorithm freesasa_algorithm;
typedef enum freesasa_verbosity freesasa_verbosity;
#endif

#define FREESASA_DEF_ALGORITHM FREESASA_LEE_RICHARDS /* src/freesasa.h:115-118 */
#define FREESASA_DEF_PROBE_RADIUS 1.4
#define FREESASA_DEF_SR_N 100
#define FREESASA_DEF_LR_N 20

/* reference src/freesasa.h:232-238 — layout must not change */
struct freesasa_parameters {
    freesasa_algorithm alg;
    double probe_radius;
    int shrake_rupley_n_points;
    int lee_richards_n_slices;
    int n_threads; /* validated as in the reference (<= 16), otherwise unused: the GPU grid replaces the pthread split */
};
/* reference src/freesasa.h:267-272 */
struct freesasa_result {
    double total;
    double *sasa; /* malloc'd; freed by freesasa_result_free() */
    int n_atoms;
    struct freesasa_parameters parameters;
};
/* reference src/coord.h:26-38 */
typedef struct coord_t {
    int n;
    int is_linked;
    double *xyz;
} coord_t;

#ifndef __cplusplus
typedef struct freesasa_parameters freesasa_parameters;
typedef struct freesasa_result freesasa_result;
#endif

extern const freesasa_parameters freesasa_default_parameters;
extern const int FREESASA_DEF_NUMBER_THREADS;

freesasa_result *freesasa_calc_coord(const double *xyz, const double *radii, int n,
                                     const freesasa_parameters *parameters);
freesasa_result *freesasa_calc(const coord_t *c, const double *radii, const freesasa_parameters *parameters);
int freesasa_lee_richards(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param);
int freesasa_shrake_rupley(double *sasa, const coord_t *c, const double *radii, const freesasa_parameters *param);
void freesasa_result_free(freesasa_result *result);

int freesasa_set_verbosity(freesasa_verbosity v);
freesasa_verbosity freesasa_get_verbosity(void);
void freesasa_set_err_out(FILE *err);
FILE *freesasa_get_err_out(void);

/* Additive: n_struct independent structures in one device pass.  results[k] is NULL-initialised by
 * the callee and receives a result per structure; returns FREESASA_SUCCESS or FREESASA_FAIL (in
 * which case no result is left allocated). */
int freesasa_calc_coord_batch(int n_struct, const double *const *xyz, const double *const *radii,
                              const int *n_atoms, const freesasa_parameters *parameters,
                              freesasa_result **results);

#ifdef __cplusplus
}
#endif
#endif
