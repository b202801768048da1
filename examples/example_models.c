/* examples/example_models.c — every MODEL of a PDB file (an NMR ensemble) in ONE device pass: what the reference CLI's
 * loop over structures (src/main.cc:334-362, one freesasa_calc_tree per model) becomes with the additive batch calls.
 *
 *   gcc -Iinclude examples/example_models.c -Lfreesasa_b200/csrc -lfreesasa_b200_host -lfsb200 -o example_models
 *   ./example_models < ensemble.pdb
 */
#include <stdio.h>
#include <stdlib.h>

#include "freesasa.h"

int main(void)
{
    int n = 0, k;
    freesasa_structure **models = freesasa_structure_array(stdin, &n, &freesasa_default_classifier, FREESASA_SEPARATE_MODELS);
    freesasa_node **trees;

    if (models == NULL) return EXIT_FAILURE;
    trees = calloc((size_t)n, sizeof *trees);
    if (trees == NULL || freesasa_calc_tree_batch(n, models, &freesasa_default_parameters, NULL, trees) != FREESASA_SUCCESS)
        return EXIT_FAILURE;
    for (k = 0; k < n; ++k) {
        const freesasa_nodearea *area = freesasa_node_area(freesasa_node_children(freesasa_node_children(trees[k])));
        printf("MODEL %d : %d atoms, total %f A2, polar %f A2, apolar %f A2\n", freesasa_structure_model(models[k]),
               freesasa_structure_n(models[k]), area->total, area->polar, area->apolar);
        freesasa_node_free(trees[k]);
        freesasa_structure_free(models[k]);
    }
    free(trees);
    free(models);
    return EXIT_SUCCESS;
}
