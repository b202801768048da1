/* examples/reference_patch/sasa_sr.c — what the reference's src/sasa_sr.c becomes in a B200 build (see sasa_lr.c next to
 * this file and INTEGRATION.md).  Validation as in src/sasa_sr.c:188-200. */
#include <assert.h>

#include "freesasa_internal.h"
#include "coord.h"

#include <fsb200.h>

#define MAX_SR_THREADS 16

int freesasa_shrake_rupley(double *sasa, const coord_t *xyz, const double *atom_radii, const freesasa_parameters *param)
{
    int n_atoms;
    assert(sasa);
    assert(xyz);
    assert(atom_radii);
    if (param == NULL) param = &freesasa_default_parameters;
    n_atoms = freesasa_coord_n(xyz);
    if (param->n_threads > MAX_SR_THREADS) return fail_msg("S&R does not support more than %d threads", MAX_SR_THREADS);
    if (param->shrake_rupley_n_points <= 0)
        return fail_msg("%d test points invalid resolution in S&R, must be > 0", param->shrake_rupley_n_points);
    if (n_atoms == 0) return freesasa_warn("in %s(): empty coordinates", __func__);
    if (param->n_threads > n_atoms)
        freesasa_warn("no sense in having more threads than atoms, only using %d threads", n_atoms);
    if (fsb200_sr(sasa, freesasa_coord_all(xyz), atom_radii, n_atoms, param->probe_radius, param->shrake_rupley_n_points) !=
        FSB200_SUCCESS)
        return fail_msg("B200 engine: %s", fsb200_last_error());
    return FREESASA_SUCCESS;
}
