/* oracle/oracle.h — CPU oracle for the FreeSASA hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, double-precision restatement of the reference algorithm behind freesasa_calc():
 *   neighbour search   reference src/nb.c:458-557
 *   Lee & Richards     reference src/sasa_lr.c:270-408
 *   Shrake & Rupley    reference src/sasa_sr.c:56-90,276-338
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (freesasa_b200/csrc/libfsb200.so) never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this restatement against the golden totals of
 * the reference's own unit tests (tests/test_freesasa.c:155-178,302-332,432-473), the arc-merge
 * known answers (src/sasa_lr.c:436-475), the 6-atom contact test (tests/test_nb.c:7-27), the
 * analytic two-sphere areas (tests/test_freesasa.c:27-43), committed per-atom fixtures generated
 * from the unmodified reference (tests/golden/), and — when oracle/_ref is built — bit-for-bit
 * against the reference itself on seeded random inputs.
 */
#ifndef FSB200_ORACLE_H
#define FSB200_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_OK 0
#define ORACLE_FAIL (-1)

/* Symmetric neighbour list in CSR form: atoms i and j are neighbours iff
 * |x_i - x_j|^2 < (R_i + R_j)^2 (strict), R already including the probe (src/nb.c:487-491).
 * *start (n+1 ints) and *list (start[n] ints, ascending within each row, no duplicates) are
 * malloc'd; free with oracle_free(). */
int oracle_neighbours(const double *xyz, const double *R, int n, int **start, int **list);

/* Per-atom SASA, Lee & Richards; radii WITHOUT probe (src/sasa_lr.c:135-138 adds it). */
int oracle_lee_richards(double *sasa, const double *xyz, const double *radii, int n,
                        double probe, int n_slices);

/* Per-atom SASA, Shrake & Rupley. */
int oracle_shrake_rupley(double *sasa, const double *xyz, const double *radii, int n,
                         double probe, int n_points);

/* Golden-spiral unit test points, 3*n_points doubles (src/sasa_sr.c:56-90). */
void oracle_test_points(int n_points, double *out);

/* Uncovered angle of the unit circle given n buried arcs (inf,sup) with 0<=inf<=sup<=2pi
 * (src/sasa_lr.c:389-408).  The array is reordered in place. */
double oracle_exposed_arc(double *arcs, int n);

/* 1 if j is in i's neighbour row (src/nb.c:559-573). */
int oracle_contact(const int *start, const int *list, int i, int j);

void oracle_free(void *p);

/* threads the oracle will use (OpenMP), for cpu_baseline reporting */
int oracle_max_threads(void);
void oracle_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
