/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under dust_b200/ may include, link or call this.
 *
 * Independent cross-check of the HDDA oracle (SURVEY.md §4 "what the build must add" (ii)):
 * a brute-force Amanatides-Woo march, in double precision, over a DENSE bit grid of one tree —
 * no hierarchy, no leaf-local frames, no shared code with oracle_trace.c.  It answers "which
 * voxel does this object-space ray enter first, and at what t" so tests can check that the
 * hierarchical traversal finds the same voxel (t to ~1e-4 relative; rays that graze a voxel
 * edge within rounding may legitimately differ and are counted, not hidden).
 */
#include "oracle_vdb.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct ODense {
    uint32_t extent;         /* voxels per axis */
    uint64_t *bits;          /* extent^3 bits, index = (x*extent + y)*extent + z */
} ODense;

ODense *o_dense_from_tree(const OTree *t)
{
    ODense *g = (ODense *)calloc(1, sizeof *g);
    g->extent = 1u << t->extent_log2[t->n_levels - 1];
    uint64_t n = (uint64_t)g->extent * g->extent * g->extent;
    g->bits = (uint64_t *)calloc((n + 63) / 64, 8);
    uint64_t nv = o_tree_iter(t, NULL, 0);
    uint32_t *xyz = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (nv ? nv : 1));
    o_tree_iter(t, xyz, nv);
    for (uint64_t i = 0; i < nv; i++) {
        uint64_t idx = ((uint64_t)xyz[3 * i] * g->extent + xyz[3 * i + 1]) * g->extent + xyz[3 * i + 2];
        g->bits[idx >> 6] |= (uint64_t)1 << (idx & 63);
    }
    free(xyz);
    return g;
}

void o_dense_free(ODense *g) { if (g) { free(g->bits); free(g); } }

static int dense_get(const ODense *g, int64_t x, int64_t y, int64_t z)
{
    uint64_t idx = ((uint64_t)x * g->extent + (uint64_t)y) * g->extent + (uint64_t)z;
    return (int)((g->bits[idx >> 6] >> (idx & 63)) & 1u);
}

/* returns 1 on hit; xyz = voxel coordinate, t = entry parameter (>= tmin) */
int o_dense_march(const ODense *g, const double o[3], const double d[3], double tmin, double tmax,
                  int32_t xyz[3], double *t_out)
{
    double E = (double)g->extent;
    double t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        if (d[a] == 0.0) { if (o[a] < 0.0 || o[a] >= E) return 0; continue; }
        double ta = (0.0 - o[a]) / d[a], tb = (E - o[a]) / d[a];
        if (ta > tb) { double s = ta; ta = tb; tb = s; }
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    if (!(t0 <= t1)) return 0;
    int64_t p[3]; double tnext[3], dt[3]; int st[3];
    for (int a = 0; a < 3; a++) {
        double x = o[a] + d[a] * t0;
        int64_t ip = (int64_t)floor(x);
        if (ip < 0) ip = 0;
        if (ip > (int64_t)g->extent - 1) ip = (int64_t)g->extent - 1;
        p[a] = ip;
        st[a] = d[a] > 0.0 ? 1 : (d[a] < 0.0 ? -1 : 0);
        if (st[a] == 0) { tnext[a] = INFINITY; dt[a] = INFINITY; }
        else {
            double plane = (double)(ip + (st[a] > 0 ? 1 : 0));
            tnext[a] = (plane - o[a]) / d[a];
            dt[a] = fabs(1.0 / d[a]);
        }
    }
    double t = t0;
    for (;;) {
        if (dense_get(g, p[0], p[1], p[2])) { xyz[0] = (int32_t)p[0]; xyz[1] = (int32_t)p[1]; xyz[2] = (int32_t)p[2]; *t_out = t; return 1; }
        int a = tnext[0] <= tnext[1] ? (tnext[0] <= tnext[2] ? 0 : 2) : (tnext[1] <= tnext[2] ? 1 : 2);
        t = tnext[a];
        if (t > t1) return 0;
        p[a] += st[a];
        if (p[a] < 0 || p[a] >= (int64_t)g->extent) return 0;
        tnext[a] += dt[a];
    }
}
