/* examples/example_coord.c — the reference's "minimal calculation" (its src/example.c and
 * tests/test_freesasa.c:138-153) written against the hot-path subset of freesasa.h that
 * libfreesasa_b200_host.so exports.  Nothing here is B200-specific: the same source compiles against the
 * reference's freesasa.h and libfreesasa.
 *
 *   gcc -Iinclude examples/example_coord.c -Lfreesasa_b200/csrc -lfreesasa_b200_host -lfsb200 \
 *       -Wl,-rpath,$PWD/freesasa_b200/csrc -lm -o example_coord
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "freesasa_b200_host.h"

int main(void)
{
    /* two intersecting spheres (tests/test_freesasa.c:79-101) and an isolated one */
    const double xyz[9] = {0, 0, 0, 2, 0, 0, 50, 0, 0};
    const double radii[3] = {1.0, 2.0, 1.5};
    freesasa_parameters p = freesasa_default_parameters;
    freesasa_result *r;
    int i;

    p.alg = FREESASA_LEE_RICHARDS;
    p.lee_richards_n_slices = 2000;
    r = freesasa_calc_coord(xyz, radii, 3, &p);
    if (r == NULL) {
        fprintf(stderr, "calculation failed\n");
        return EXIT_FAILURE;
    }
    printf("Lee & Richards, %d slices: total %.4f A2\n", r->parameters.lee_richards_n_slices, r->total);
    for (i = 0; i < r->n_atoms; ++i) printf("  atom %d: %.4f\n", i, r->sasa[i]);
    {
        const double R = radii[2] + p.probe_radius, pi = 3.14159265358979323846;
        if (fabs(r->sasa[2] - 4 * pi * R * R) > 1e-3) {
            fprintf(stderr, "isolated sphere is off\n");
            return EXIT_FAILURE;
        }
    }
    freesasa_result_free(r);

    p.alg = FREESASA_SHRAKE_RUPLEY;
    p.shrake_rupley_n_points = 5000;
    r = freesasa_calc_coord(xyz, radii, 3, &p);
    if (r == NULL) return EXIT_FAILURE;
    printf("Shrake & Rupley, %d points: total %.4f A2\n", r->parameters.shrake_rupley_n_points, r->total);
    freesasa_result_free(r);
    return EXIT_SUCCESS;
}
