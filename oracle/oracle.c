/* oracle/oracle.c — CPU oracle for the FreeSASA hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * See oracle.h for scope and how parity is pinned.  This is a restatement of WHAT the reference
 * computes (same arithmetic expressions, in double, so that results agree with the reference to
 * the last bit on non-degenerate inputs), not of HOW it is organised: the reference grows four
 * ragged arrays per atom with realloc and walks "forward" cell pairs (src/nb.c:86-115,409-522);
 * here the neighbour list is a duplicate-free CSR built by a counting sort over a uniform grid
 * and a 27-cell gather, and all per-atom loops are OpenMP-parallel.
 *
 * Compile with the reference's own floating-point environment: gcc -O2, no -march=native, so no
 * FMA contraction (matters for the Shrake-Rupley inside/outside decisions).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const double TWO_PI = 2 * M_PI;

void oracle_free(void *p) { free(p); }

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Uniform grid.  The reference uses cell edge 2*max(R) and bounds padded by half a cell        */
/* (src/nb.c:43-72,543); the neighbour SET only requires edge >= 2*max(R), so the edge is also  */
/* allowed to grow to keep the grid at most ~8 cells per atom (sparse inputs).                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double lo[3], edge;
    int dim[3];
    int n_cells;
    int *cell_start; /* n_cells+1 */
    int *order;      /* atoms sorted by cell (ascending atom index inside a cell) */
} grid_t;

static void grid_release(grid_t *g)
{
    free(g->cell_start);
    free(g->order);
    g->cell_start = g->order = NULL;
}

static int grid_cell_of(const grid_t *g, const double *p, int c[3])
{
    for (int a = 0; a < 3; ++a) {
        int k = (int)((p[a] - g->lo[a]) / g->edge);
        if (k < 0) k = 0;
        if (k >= g->dim[a]) k = g->dim[a] - 1;
        c[a] = k;
    }
    return c[0] + g->dim[0] * (c[1] + g->dim[1] * c[2]);
}

static int grid_build(grid_t *g, const double *xyz, const double *R, int n)
{
    double hi[3], rmax = 0;
    memset(g, 0, sizeof *g);
    for (int a = 0; a < 3; ++a) g->lo[a] = hi[a] = xyz[a];
    for (int i = 0; i < n; ++i) {
        for (int a = 0; a < 3; ++a) {
            double v = xyz[3 * i + a];
            if (v < g->lo[a]) g->lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
        if (R[i] > rmax) rmax = R[i];
    }
    g->edge = 2 * rmax;
    if (!(g->edge > 0)) g->edge = 1.0; /* all radii zero: nobody has neighbours */
    for (;;) {
        double cells = 1;
        for (int a = 0; a < 3; ++a) {
            double ext = (hi[a] - g->lo[a]) / g->edge;
            g->dim[a] = (int)ext + 1;
            cells *= g->dim[a];
        }
        if (cells <= 8.0 * n + 64) break;
        g->edge *= 1.26; /* ~ halves the cell count */
    }
    g->n_cells = g->dim[0] * g->dim[1] * g->dim[2];
    g->cell_start = calloc((size_t)g->n_cells + 1, sizeof(int));
    g->order = malloc(sizeof(int) * (size_t)n);
    int *cell = malloc(sizeof(int) * (size_t)n);
    if (!g->cell_start || !g->order || !cell) {
        free(cell);
        grid_release(g);
        return ORACLE_FAIL;
    }
    int c[3];
    for (int i = 0; i < n; ++i) {
        cell[i] = grid_cell_of(g, xyz + 3 * i, c);
        ++g->cell_start[cell[i] + 1];
    }
    for (int k = 0; k < g->n_cells; ++k) g->cell_start[k + 1] += g->cell_start[k];
    int *fill = malloc(sizeof(int) * (size_t)g->n_cells);
    if (!fill) {
        free(cell);
        grid_release(g);
        return ORACLE_FAIL;
    }
    memcpy(fill, g->cell_start, sizeof(int) * (size_t)g->n_cells);
    for (int i = 0; i < n; ++i) g->order[fill[cell[i]]++] = i;
    free(fill);
    free(cell);
    return ORACLE_OK;
}

/* Visit every atom j != i with |x_j-x_i|^2 < (R_i+R_j)^2.  The comparison is the reference's,
 * term for term (src/nb.c:483-491): cut2 = (ri+rj)*(ri+rj); dx*dx+dy*dy+dz*dz < cut2.
 * If out != NULL the hits are stored (unsorted); returns the number of hits. */
static int gather_row(const grid_t *g, const double *xyz, const double *R, int i, int *out)
{
    int c[3], count = 0;
    const double xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2], ri = R[i];
    grid_cell_of(g, xyz + 3 * i, c);
    for (int kz = c[2] - 1; kz <= c[2] + 1; ++kz) {
        if (kz < 0 || kz >= g->dim[2]) continue;
        for (int ky = c[1] - 1; ky <= c[1] + 1; ++ky) {
            if (ky < 0 || ky >= g->dim[1]) continue;
            for (int kx = c[0] - 1; kx <= c[0] + 1; ++kx) {
                if (kx < 0 || kx >= g->dim[0]) continue;
                int cell = kx + g->dim[0] * (ky + g->dim[1] * kz);
                for (int p = g->cell_start[cell]; p < g->cell_start[cell + 1]; ++p) {
                    int j = g->order[p];
                    if (j == i) continue;
                    double rj = R[j];
                    double cut2 = (ri + rj) * (ri + rj);
                    double dx = xyz[3 * j] - xi, dy = xyz[3 * j + 1] - yi, dz = xyz[3 * j + 2] - zi;
                    if (dx * dx + dy * dy + dz * dz < cut2) {
                        if (out) out[count] = j;
                        ++count;
                    }
                }
            }
        }
    }
    return count;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

int oracle_neighbours(const double *xyz, const double *R, int n, int **start_out, int **list_out)
{
    grid_t g;
    *start_out = NULL;
    *list_out = NULL;
    if (n <= 0) return ORACLE_FAIL;
    if (grid_build(&g, xyz, R, n)) return ORACLE_FAIL;
    int *start = malloc(sizeof(int) * ((size_t)n + 1));
    if (!start) {
        grid_release(&g);
        return ORACLE_FAIL;
    }
    start[0] = 0;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) start[i + 1] = gather_row(&g, xyz, R, i, NULL);
    for (int i = 0; i < n; ++i) start[i + 1] += start[i];
    int *list = malloc(sizeof(int) * (size_t)(start[n] > 0 ? start[n] : 1));
    if (!list) {
        free(start);
        grid_release(&g);
        return ORACLE_FAIL;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; ++i) {
        int m = gather_row(&g, xyz, R, i, list + start[i]);
        qsort(list + start[i], (size_t)m, sizeof(int), cmp_int);
    }
    grid_release(&g);
    *start_out = start;
    *list_out = list;
    return ORACLE_OK;
}

int oracle_contact(const int *start, const int *list, int i, int j)
{
    for (int p = start[i]; p < start[i + 1]; ++p)
        if (list[p] == j) return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Lee & Richards                                                                              */
/* ------------------------------------------------------------------------------------------ */

/* Sort (inf,sup) pairs by inf.  Any correct sort gives the same sum below: when two arcs share
 * an inf the later one adds nothing (its inf is <= the running sup). */
static void order_arcs(double *arcs, int n)
{
    for (int gap = n / 2; gap > 0; gap /= 2) { /* shell sort, pairs move together */
        for (int k = gap; k < n; ++k) {
            double lo = arcs[2 * k], hi = arcs[2 * k + 1];
            int m = k;
            while (m >= gap && arcs[2 * (m - gap)] > lo) {
                arcs[2 * m] = arcs[2 * (m - gap)];
                arcs[2 * m + 1] = arcs[2 * (m - gap) + 1];
                m -= gap;
            }
            arcs[2 * m] = lo;
            arcs[2 * m + 1] = hi;
        }
    }
}

/* src/sasa_lr.c:389-408: sweep the sorted arcs, add every gap between the running supremum and
 * the next infimum, then the gap from the last supremum back to 2pi; n == 0 -> 2pi. */
double oracle_exposed_arc(double *arcs, int n)
{
    if (n == 0) return TWO_PI;
    order_arcs(arcs, n);
    double sum = arcs[0], sup = arcs[1];
    for (int k = 1; k < n; ++k) {
        if (sup < arcs[2 * k]) sum += arcs[2 * k] - sup;
        if (arcs[2 * k + 1] > sup) sup = arcs[2 * k + 1];
    }
    return sum + TWO_PI - sup;
}

/* One atom, src/sasa_lr.c:270-364.  R includes the probe.  scratch holds 4 doubles per
 * neighbour (an arc crossing zero is stored as two). */
static double lr_atom(const double *xyz, const double *R, const int *nb, int nn, int i,
                      int n_slices, double *scratch)
{
    const double xi = xyz[3 * i], yi = xyz[3 * i + 1], zi = xyz[3 * i + 2], Ri = R[i];
    const double delta = 2 * Ri / n_slices; /* :304 */
    double z = zi - Ri - 0.5 * delta;       /* :305 */
    double area = 0;

    for (int s = 0; s < n_slices; ++s) {
        z += delta;                            /* :307  slice centre, accumulated */
        double di = fabs(zi - z);              /* :308 */
        double Ri_p2 = Ri * Ri - di * di;      /* :309 */
        if (Ri_p2 < 0) continue;               /* :310 */
        double Ri_p = sqrt(Ri_p2);             /* :311 */
        if (Ri_p <= 0) continue;               /* :312 */
        int n_arcs = 0, buried = 0;
        for (int k = 0; k < nn; ++k) {
            int j = nb[k];
            double zj = xyz[3 * j + 2], Rj = R[j];
            double dj = fabs(zj - z);          /* :317 */
            if (!(dj < Rj)) continue;          /* :320 */
            double Rj_p2 = Rj * Rj - dj * dj;  /* :321 */
            double Rj_p = sqrt(Rj_p2);         /* :322 */
            double xd = xyz[3 * j] - xi, yd = xyz[3 * j + 1] - yi; /* src/nb.c:445-448,484-485 */
            double dij = sqrt(xd * xd + yd * yd);                  /* src/nb.c:440 */
            if (dij >= Ri_p + Rj_p) continue;  /* :324  circles do not touch */
            if (dij + Ri_p < Rj_p) {           /* :327  circle i inside circle j */
                buried = 1;
                break;
            }
            if (dij + Rj_p < Ri_p) continue;   /* :331  circle j inside circle i */
            double alpha = acos((Ri_p2 + dij * dij - Rj_p2) / (2.0 * Ri_p * dij)); /* :335 */
            double beta = atan2(yd, xd) + M_PI;                                     /* :337 */
            double inf = beta - alpha, sup = beta + alpha;
            if (inf < 0) inf += TWO_PI;        /* :340 */
            if (sup > 2 * M_PI) sup -= TWO_PI; /* :341 */
            double *a = scratch + 2 * n_arcs;
            if (sup < inf) {                   /* :344-351  arc passes through zero: split */
                a[0] = 0;
                a[1] = sup;
                a[2] = inf;
                a[3] = TWO_PI;
                n_arcs += 2;
            } else {
                a[0] = inf;
                a[1] = sup;
                n_arcs += 1;
            }
        }
        if (!buried) area += delta * Ri * oracle_exposed_arc(scratch, n_arcs); /* :359-361 */
    }
    return area;
}

int oracle_lee_richards(double *sasa, const double *xyz, const double *radii, int n, double probe,
                        int n_slices)
{
    if (n <= 0 || n_slices <= 0) return ORACLE_FAIL;
    double *R = malloc(sizeof(double) * (size_t)n);
    int *start = NULL, *list = NULL, rc = ORACLE_OK;
    if (!R) return ORACLE_FAIL;
    for (int i = 0; i < n; ++i) R[i] = radii[i] + probe; /* src/sasa_lr.c:135-138 */
    if (oracle_neighbours(xyz, R, n, &start, &list)) {
        free(R);
        return ORACLE_FAIL;
    }
    int max_nn = 0;
    for (int i = 0; i < n; ++i)
        if (start[i + 1] - start[i] > max_nn) max_nn = start[i + 1] - start[i];
#pragma omp parallel
    {
        double *scratch = malloc(sizeof(double) * 4 * (size_t)(max_nn + 1));
        if (!scratch) {
#pragma omp atomic write
            rc = ORACLE_FAIL;
        } else {
#pragma omp for schedule(dynamic, 64)
            for (int i = 0; i < n; ++i)
                sasa[i] = lr_atom(xyz, R, list + start[i], start[i + 1] - start[i], i, n_slices,
                                  scratch);
        }
        free(scratch);
    }
    free(start);
    free(list);
    free(R);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Shrake & Rupley                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* src/sasa_sr.c:56-90.  z and longitude are ACCUMULATED, exactly as there, so the points are
 * bit-identical (first point at z = 1 - dz/2, longitude 0). */
void oracle_test_points(int n_points, double *out)
{
    const double dlong = M_PI * (3 - sqrt(5)), dz = 2.0 / n_points;
    double longitude = 0, z = 1 - dz / 2;
    for (int k = 0; k < n_points; ++k) {
        double r = sqrt(1 - z * z);
        out[3 * k] = cos(longitude) * r;
        out[3 * k + 1] = sin(longitude) * r;
        out[3 * k + 2] = z;
        z -= dz;
        longitude += dlong;
    }
}

/* One atom, src/sasa_sr.c:276-338.  A test point p = u*R_i + x_i (scale, then translate:
 * src/sasa_sr.c:297-299 via src/coord.c:306-342) is buried iff some neighbour a has
 * |p - x_a|^2 <= R_a^2; the reference's "start with the neighbour that hid the previous point"
 * loop (:305-328) is only an early-out and yields the same count. */
static double sr_atom(const double *xyz, const double *R, const double *R2, const int *nb, int nn,
                      int i, const double *unit, int n_points)
{
    const double ri = R[i];
    const double *vi = xyz + 3 * i;
    int exposed = 0, last = 0;
    for (int q = 0; q < n_points; ++q) {
        double px = unit[3 * q] * ri, py = unit[3 * q + 1] * ri, pz = unit[3 * q + 2] * ri;
        px += vi[0];
        py += vi[1];
        pz += vi[2];
        int hidden = 0;
        for (int t = 0; t < nn && !hidden; ++t) {
            int k = (t + last) % nn; /* begin with the last occluder */
            int a = nb[k];
            double dx = px - xyz[3 * a], dy = py - xyz[3 * a + 1], dz = pz - xyz[3 * a + 2];
            if (!(dx * dx + dy * dy + dz * dz > R2[a])) { /* :317,324 */
                hidden = 1;
                last = k;
            }
        }
        exposed += !hidden;
    }
    return (4.0 * M_PI * ri * ri * exposed) / n_points; /* :337 */
}

int oracle_shrake_rupley(double *sasa, const double *xyz, const double *radii, int n, double probe,
                         int n_points)
{
    if (n <= 0 || n_points <= 0) return ORACLE_FAIL;
    double *R = malloc(sizeof(double) * (size_t)n), *R2 = malloc(sizeof(double) * (size_t)n);
    double *unit = malloc(sizeof(double) * 3 * (size_t)n_points);
    int *start = NULL, *list = NULL;
    if (!R || !R2 || !unit) goto fail;
    for (int i = 0; i < n; ++i) { /* src/sasa_sr.c:143-147 */
        double ri = radii[i] + probe;
        R[i] = ri;
        R2[i] = ri * ri;
    }
    oracle_test_points(n_points, unit);
    if (oracle_neighbours(xyz, R, n, &start, &list)) goto fail;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i)
        sasa[i] = sr_atom(xyz, R, R2, list + start[i], start[i + 1] - start[i], i, unit, n_points);
    free(start);
    free(list);
    free(R);
    free(R2);
    free(unit);
    return ORACLE_OK;
fail:
    free(R);
    free(R2);
    free(unit);
    return ORACLE_FAIL;
}
