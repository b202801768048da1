/* examples/reference_patch/sasa_lr.c — what the reference's src/sasa_lr.c becomes in a B200 build.
 *
 * Compiled against the REFERENCE's own headers (src/freesasa_internal.h, src/coord.h) in place of its src/sasa_lr.c; every
 * other file of the reference — structure.c, classifier.c, node.c, pdb.c, freesasa.c, ... — is used unmodified.  The
 * parameter validation keeps the meaning of src/sasa_lr.c:177-193; the numeric work is one call into libfsb200.so.
 * `make -C oracle patched` builds oracle/_ref/libfreesasa_patched.so this way (dev container only). */
#include <assert.h>

#include "freesasa_internal.h"
#include "coord.h"

#include <fsb200.h>

#define MAX_LR_THREADS 16

int freesasa_lee_richards(double *sasa, const coord_t *xyz, const double *atom_radii, const freesasa_parameters *param)
{
    int n_atoms;
    assert(sasa);
    assert(xyz);
    assert(atom_radii);
    if (param == NULL) param = &freesasa_default_parameters;
    n_atoms = freesasa_coord_n(xyz);
    if (param->n_threads > MAX_LR_THREADS) return fail_msg("L&R does not support more than %d threads", MAX_LR_THREADS);
    if (param->lee_richards_n_slices <= 0)
        return fail_msg("%d slices per atom invalid resolution in L&R, must be > 0", param->lee_richards_n_slices);
    if (n_atoms == 0) return freesasa_warn("in %s(): empty coordinates", __func__);
    if (param->n_threads > n_atoms)
        freesasa_warn("no sense in having more threads than atoms, only using %d threads", n_atoms);
    if (fsb200_lr(sasa, freesasa_coord_all(xyz), atom_radii, n_atoms, param->probe_radius, param->lee_richards_n_slices) !=
        FSB200_SUCCESS)
        return fail_msg("B200 engine: %s", fsb200_last_error());
    return FREESASA_SUCCESS;
}
