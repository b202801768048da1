/* examples/example_structure.c — PDB on stdin -> class areas, per-chain areas and (with an argument) the PDB file with
 * radii and SASA written back, through the reference's own API names backed by the B200 engine.
 *
 *   gcc -Iinclude examples/example_structure.c -Lfreesasa_b200/csrc -lfreesasa_b200_host -lfsb200 -o example_structure
 *   ./example_structure [out.pdb] < input.pdb
 */
#include <stdio.h>
#include <stdlib.h>

#include "freesasa.h"

int main(int argc, char **argv)
{
    freesasa_structure *structure = freesasa_structure_from_pdb(stdin, &freesasa_default_classifier, 0);
    freesasa_node *tree, *chain;
    const freesasa_nodearea *area;

    if (structure == NULL) return EXIT_FAILURE;
    tree = freesasa_calc_tree(structure, &freesasa_default_parameters, "stdin");
    if (tree == NULL) {
        freesasa_structure_free(structure);
        return EXIT_FAILURE;
    }
    /* root -> result -> structure */
    area = freesasa_node_area(freesasa_node_children(freesasa_node_children(tree)));
    printf("atoms  : %d\n", freesasa_structure_n(structure));
    printf("Total  : %f A2\nApolar : %f A2\nPolar  : %f A2\n", area->total, area->apolar, area->polar);
    for (chain = freesasa_node_children(freesasa_node_children(freesasa_node_children(tree))); chain;
         chain = freesasa_node_next(chain))
        printf("CHAIN %s : %f A2 (%d residues)\n", freesasa_node_name(chain), freesasa_node_area(chain)->total,
               freesasa_node_chain_n_residues(chain));
    if (argc > 1) {
        FILE *out = fopen(argv[1], "w");
        if (out == NULL || freesasa_write_pdb(out, tree) != FREESASA_SUCCESS) return EXIT_FAILURE;
        fclose(out);
    }
    freesasa_node_free(tree);
    freesasa_structure_free(structure);
    return EXIT_SUCCESS;
}
