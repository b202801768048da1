/* oracle/ref_stubs.c — TEST INFRASTRUCTURE ONLY.
 *
 * The reference's result tree (src/node.c) refers to four entry points of its selection
 * mini-language (src/selection.c + flex/bison output).  That language is outside the hot path
 *  and is not compiled into oracle/_ref/libfreesasa_ref.so; these stubs only
 * satisfy the dynamic linker so the shared object loads with RTLD_NOW.  Nothing in tests/,
 * bench.py or smoke() ever reaches them; if something does, fail loudly.
 */
#include <stdio.h>
#include <stdlib.h>

static void unreachable(const char *name)
{
    fprintf(stderr, "oracle/_ref: %s() is not part of the hot-path reference build\n", name);
    abort();
}

void *freesasa_selection_clone(const void *s) { (void)s; unreachable("freesasa_selection_clone"); return NULL; }
void freesasa_selection_free(void *s) { (void)s; /* freeing NULL selections is legal */ }
const char *freesasa_selection_name(const void *s) { (void)s; unreachable("freesasa_selection_name"); return NULL; }
double freesasa_selection_area(const void *s) { (void)s; unreachable("freesasa_selection_area"); return 0; }
