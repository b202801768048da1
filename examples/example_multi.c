/* examples/example_multi.c — one structure on every GPU of the box through the C ABI (include/fsb200.h).
 *
 * What a C caller of freesasa_calc_structure() (reference src/freesasa.c:144-153) on a virus-capsid-sized structure would
 * do after the two-file patch of INTEGRATION.md: the same arrays, one more argument.  The inputs are replicated over
 * NVLink, the outputs partitioned; the answer is the whole-structure SASA, bit-identical to the one-GPU call.
 *
 *   gcc -Iinclude examples/example_multi.c -Lfreesasa_b200/csrc -lfsb200 -Wl,-rpath,$PWD/freesasa_b200/csrc -lm -o example_multi
 *   ./example_multi 1000000        # atoms (default 200000); needs >= 1 B200, uses all of them
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fsb200.h"

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 200000;
    const double spacing = 2.6, probe = 1.4;
    const int side = (int)ceil(cbrt((double)n));
    double *xyz = malloc(sizeof(double) * 3 * (size_t)n), *radii = malloc(sizeof(double) * (size_t)n);
    double *one = malloc(sizeof(double) * (size_t)n), *all = malloc(sizeof(double) * (size_t)n);
    fsb200_multi_stats st;
    double total = 0;
    int i;

    if (n <= 0 || !xyz || !radii || !one || !all) return EXIT_FAILURE;
    for (i = 0; i < n; ++i) { /* a jittered cubic lattice at protein heavy-atom density */
        unsigned h = (unsigned)i * 2654435761u;
        xyz[3 * i] = spacing * (i % side) + 0.4 * ((h & 255) / 255.0 - 0.5);
        xyz[3 * i + 1] = spacing * ((i / side) % side) + 0.4 * (((h >> 8) & 255) / 255.0 - 0.5);
        xyz[3 * i + 2] = spacing * (i / (side * side)) + 0.4 * (((h >> 16) & 255) / 255.0 - 0.5);
        radii[i] = 1.6 + 0.3 * ((h >> 24) / 255.0);
    }
    if (!fsb200_available()) {
        fprintf(stderr, "no B200 visible: %s\n", "the engine has no CPU path");
        return EXIT_FAILURE;
    }
    if (fsb200_lr_multi(one, xyz, radii, n, probe, 20, 1) != FSB200_SUCCESS ||
        fsb200_lr_multi(all, xyz, radii, n, probe, 20, 0 /* every visible GPU */) != FSB200_SUCCESS) {
        fprintf(stderr, "calculation failed: %s\n", fsb200_last_error());
        return EXIT_FAILURE;
    }
    fsb200_get_multi_stats(&st);
    for (i = 0; i < n; ++i) total += all[i];
    printf("%d atoms on %d GPU(s): total %.3f A2; call %.3f ms (upload %.3f, compute %.3f, download %.3f)\n", n, st.n_devices,
           total, st.total_ms, st.upload_ms, st.compute_ms, st.download_ms);
    if (memcmp(one, all, sizeof(double) * (size_t)n) != 0) {
        fprintf(stderr, "multi-GPU result differs from the one-GPU result\n");
        return EXIT_FAILURE;
    }
    printf("bit-identical to the one-GPU call\n");
    fsb200_trim();
    free(xyz); free(radii); free(one); free(all);
    return EXIT_SUCCESS;
}
