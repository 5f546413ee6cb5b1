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
    int32_t lo[3];           /* the grid covers the voxel box [lo, lo + size) — the occupied bounds rounded out to whole leaves */
    uint32_t size[3];
    uint64_t *bits;          /* size^3 bits, index = ((x-lo.x)*size.y + (y-lo.y))*size.z + (z-lo.z) */
} ODense;

ODense *o_dense_from_tree(const OTree *t)
{
    ODense *g = (ODense *)calloc(1, sizeof *g);
    uint64_t nv = o_tree_iter(t, NULL, 0);
    uint32_t *xyz = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (nv ? nv : 1));
    o_tree_iter(t, xyz, nv);
    uint32_t mn[3] = { 0, 0, 0 }, mx[3] = { 4, 4, 4 };
    for (uint64_t i = 0; i < nv; i++)
        for (int a = 0; a < 3; a++) {
            uint32_t v = xyz[3 * i + a];
            if (i == 0 || (v & ~3u) < mn[a]) mn[a] = v & ~3u;
            if (i == 0 || (v | 3u) + 1u > mx[a]) mx[a] = (v | 3u) + 1u;
        }
    for (int a = 0; a < 3; a++) { g->lo[a] = (int32_t)mn[a]; g->size[a] = mx[a] - mn[a]; }
    uint64_t n = (uint64_t)g->size[0] * g->size[1] * g->size[2];
    if (n > ((uint64_t)1 << 34)) { free(xyz); free(g); return NULL; }          /* 2 GiB of bits: refuse, the caller picks a smaller domain */
    g->bits = (uint64_t *)calloc((n + 63) / 64, 8);
    for (uint64_t i = 0; i < nv; i++) {
        uint64_t idx = ((uint64_t)(xyz[3 * i] - mn[0]) * g->size[1] + (xyz[3 * i + 1] - mn[1])) * g->size[2] + (xyz[3 * i + 2] - mn[2]);
        g->bits[idx >> 6] |= (uint64_t)1 << (idx & 63);
    }
    free(xyz);
    return g;
}

void o_dense_free(ODense *g) { if (g) { free(g->bits); free(g); } }

static int dense_get(const ODense *g, int64_t x, int64_t y, int64_t z)
{
    uint64_t idx = ((uint64_t)(x - g->lo[0]) * g->size[1] + (uint64_t)(y - g->lo[1])) * g->size[2] + (uint64_t)(z - g->lo[2]);
    return (int)((g->bits[idx >> 6] >> (idx & 63)) & 1u);
}

/* exit time of the 4x4x4 leaf box that contains voxel p, for the ray (o, d) */
static double leaf_exit_time(const int64_t p[3], const double o[3], const double d[3])
{
    double te = INFINITY;
    for (int a = 0; a < 3; a++) {
        if (d[a] == 0.0) continue;
        double base = (double)(p[a] & ~(int64_t)3);
        double plane = d[a] > 0.0 ? base + 4.0 : base;
        double ta = (plane - o[a]) / d[a];
        if (ta < te) te = ta;
    }
    return te;
}

/* Returns 1 on hit; xyz = voxel coordinate, t = entry parameter (>= tmin).
 * guard != 0: the marcher also applies the ONE rule of the reference's intersection shader that is geometry, not arithmetic
 * (assets/shaders/primary/hit.rint:107): inside a 4x4x4 leaf the DDA looks at the next voxel only while `t + 0.001 < t_exit(leaf)`,
 * so a voxel entered within 0.001 of the leaf's exit is skipped (the first voxel of a leaf is always examined).  The rule is
 * restated here from the shader, in double precision and without any code of oracle_trace.c. */
static int dense_march(const ODense *g, const double o[3], const double d[3], double tmin, double tmax, int guard,
                       int32_t xyz[3], double *t_out)
{
    double t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; a++) {
        double L = (double)g->lo[a], H = (double)g->lo[a] + (double)g->size[a];
        if (d[a] == 0.0) { if (o[a] < L || o[a] >= H) return 0; continue; }
        double ta = (L - o[a]) / d[a], tb = (H - o[a]) / d[a];
        if (ta > tb) { double s = ta; ta = tb; tb = s; }
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    if (!(t0 <= t1)) return 0;
    int64_t p[3]; double tnext[3], dt[3]; int st[3];
    for (int a = 0; a < 3; a++) {
        double x = o[a] + d[a] * t0;
        int64_t ip = (int64_t)floor(x);
        if (ip < g->lo[a]) ip = g->lo[a];
        if (ip > (int64_t)g->lo[a] + (int64_t)g->size[a] - 1) ip = (int64_t)g->lo[a] + (int64_t)g->size[a] - 1;
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
    int64_t leaf[3] = { p[0] >> 2, p[1] >> 2, p[2] >> 2 };
    double t_leaf_exit = leaf_exit_time(p, o, d);
    int skip = 0;                                       /* guard: the rest of this leaf is not examined */
    for (;;) {
        if (!skip && dense_get(g, p[0], p[1], p[2])) {
            if (t > tmax) return 0;
            xyz[0] = (int32_t)p[0]; xyz[1] = (int32_t)p[1]; xyz[2] = (int32_t)p[2]; *t_out = t; return 1;
        }
        int a = tnext[0] <= tnext[1] ? (tnext[0] <= tnext[2] ? 0 : 2) : (tnext[1] <= tnext[2] ? 1 : 2);
        t = tnext[a];
        if (t > t1) return 0;
        p[a] += st[a];
        if (p[a] < g->lo[a] || p[a] >= (int64_t)g->lo[a] + (int64_t)g->size[a]) return 0;
        tnext[a] += dt[a];
        if ((p[a] >> 2) != leaf[a]) {                   /* a new leaf: its first voxel is always examined */
            leaf[a] = p[a] >> 2;
            t_leaf_exit = leaf_exit_time(p, o, d);
            skip = 0;
        } else if (guard && t + 0.001 >= t_leaf_exit) skip = 1;
    }
}

int o_dense_march(const ODense *g, const double o[3], const double d[3], double tmin, double tmax, int32_t xyz[3], double *t_out)
{
    return dense_march(g, o, d, tmin, tmax, 0, xyz, t_out);
}

/* n rays (o, d: n x 3 doubles, object space of this grid's tree) in one call; hit[i] = 0 / 1 */
void o_dense_march_batch(const ODense *g, const double *o, const double *d, uint64_t n, double tmin, double tmax, int guard,
                         int32_t *xyz, double *t, uint8_t *hit)
{
    for (uint64_t i = 0; i < n; i++) {
        t[i] = INFINITY;
        hit[i] = (uint8_t)dense_march(g, o + 3 * i, d + 3 * i, tmin, tmax, guard, xyz + 3 * i, &t[i]);
    }
}
