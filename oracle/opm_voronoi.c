/* opm_voronoi.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, like opm_oracle.c) of the Voronoi fluence estimator, the reference's
 * default for FluenceDetector (nodes/fluence_detector/mod.rs:61) and for the ghost-focus LIDT check (analyzers/ghostfocus.rs:61):
 *
 *   RaysHitMap::calc_fluence_with_voronoi   surface/hit_map/rays_hit_map.rs:407-521
 *   create_voronoi_cells                    utils/griddata.rs:265-345
 *   calc_closed_poly_area                   utils/griddata.rs:112-134
 *   interpolate_3d_triangulated_scatter_data utils/griddata.rs:387-436
 *   all_points_on_same_line                 utils/griddata.rs:437-456
 *
 * PARITY UNPINNED beyond the reference's two known answers (hit_map/mod.rs:942-982: centre bin 2 and 4 J/m^2): the geometry
 * itself is computed by two crates that are NOT under /root/reference -- `voronator 0.2.1` (Cargo.lock:8019-8020: Delaunay
 * triangulation by `delaunator`, Voronoi cells from the circumcentres, clipped against a bounding polygon) and `spade 2.14.0`
 * (Cargo.lock:7008-7009: natural-neighbour interpolation).  They are restated here by their published DEFINITIONS, by brute
 * force, with no triangulation at all:
 *   - the Voronoi cell of site i = a rectangle around everything clipped by the half-plane "closer to s_i than to s_j" of every
 *     other site j (what voronator's cells are whenever its helper points lie far enough outside), then clipped against the
 *     bounding polygon as voronator does: Sutherland-Hodgman with every EDGE of that polygon as a half-plane (the polygon is not
 *     convex for elongated hulls; the clip then keeps what lies inside all edge half-planes);
 *   - the number of Delaunay triangles of n sites with k of them on the boundary of the convex hull = 2n - 2 - k (Euler);
 *   - Sibson's natural-neighbour coordinates of a query point q = the shares of the cell q would get, were it inserted, that
 *     it takes from each old cell (spade's `NaturalNeighbor::interpolate`): None outside the convex hull, the vertex value on a
 *     site, linear between the end points on a hull edge.
 * What stays open: delaunator's treatment of duplicate and of collinear hull points (hull vertex list, triangle count), voronator's
 * helper-point placement, spade's exact-arithmetic tie breaking -- none of which the reference's tests exercise. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "opm_oracle.h"

typedef struct { double x, y; } P2;

/* griddata.rs:112-134 */
static int poly_area(const P2* p, int n, double* out) {
    if (n < 3) return ORC_ERR_ARG;
    double area = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        if (!isfinite(p[i].x) || !isfinite(p[i].y)) return ORC_ERR_ARG;
        area += p[i].x * p[j].y;
        area -= p[i].y * p[j].x;
    }
    *out = fabs(area) / 2.0;
    return ORC_OK;
}

/* Sutherland-Hodgman against the half-plane d(v) = (v - m) . nrm <= 0; ids[k] = generator of the edge poly[k] -> poly[k+1]
 * (may be NULL); the edge the cut creates gets `id`.  Returns the new vertex count (0: nothing left). */
static int clip_halfplane(const P2* in, const int* in_id, int n, P2 m, P2 nrm, int id, P2* out, int* out_id, int cap) {
    int k = 0;
    for (int a = 0; a < n; ++a) {
        int b = (a + 1) % n;
        double da = (in[a].x - m.x) * nrm.x + (in[a].y - m.y) * nrm.y;
        double db = (in[b].x - m.x) * nrm.x + (in[b].y - m.y) * nrm.y;
        int ia = da <= 0.0, ib = db <= 0.0;
        if (ia) {
            if (k < cap) { out[k] = in[a]; if (out_id) out_id[k] = ib ? (in_id ? in_id[a] : -1) : (in_id ? in_id[a] : -1); }
            ++k;
        }
        if (ia != ib) {
            double t = da / (da - db);
            P2 x = { in[a].x + t * (in[b].x - in[a].x), in[a].y + t * (in[b].y - in[a].y) };
            if (k < cap) {
                out[k] = x;
                /* leaving the half-plane: the next edge runs along the cut; entering: the rest of the old edge follows */
                if (out_id) out_id[k] = ia ? id : (in_id ? in_id[a] : -1);
            }
            ++k;
        }
    }
    return k > cap ? -1 : k;
}

static double cross3(P2 o, P2 a, P2 b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); }

typedef struct { double x, y; int idx; } HP;
static int hp_cmp(const void* a, const void* b) {
    const HP* p = (const HP*)a; const HP* q = (const HP*)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}
/* convex hull (Andrew's monotone chain), counter-clockwise, strict vertices only; returns the vertex count, hull[] = indices.
 * *on_boundary = number of DISTINCT input points on the hull's boundary (vertices and points on its edges). */
static int convex_hull(const P2* s, int n, int* hull, int* on_boundary) {
    HP* v = (HP*)malloc((size_t)n * sizeof(HP));
    for (int i = 0; i < n; ++i) { v[i].x = s[i].x; v[i].y = s[i].y; v[i].idx = i; }
    qsort(v, (size_t)n, sizeof(HP), hp_cmp);
    int m = 0;
    for (int i = 0; i < n; ++i)  /* distinct points */
        if (m == 0 || v[i].x != v[m - 1].x || v[i].y != v[m - 1].y) v[m++] = v[i];
    int* st = (int*)malloc((size_t)(2 * m + 2) * sizeof(int));
    int k = 0;
    for (int i = 0; i < m; ++i) {
        while (k >= 2 && cross3((P2){v[st[k - 2]].x, v[st[k - 2]].y}, (P2){v[st[k - 1]].x, v[st[k - 1]].y}, (P2){v[i].x, v[i].y}) <= 0.0) --k;
        st[k++] = i;
    }
    for (int i = m - 2, t = k + 1; i >= 0; --i) {
        while (k >= t && cross3((P2){v[st[k - 2]].x, v[st[k - 2]].y}, (P2){v[st[k - 1]].x, v[st[k - 1]].y}, (P2){v[i].x, v[i].y}) <= 0.0) --k;
        st[k++] = i;
    }
    int h = k > 1 ? k - 1 : k;
    for (int i = 0; i < h; ++i) hull[i] = v[st[i]].idx;
    if (on_boundary) {
        int cnt = 0;
        for (int i = 0; i < m; ++i) {
            P2 q = { v[i].x, v[i].y };
            int on = 0;
            for (int e = 0; e < h && !on; ++e) {
                P2 a = s[hull[e]], b = s[hull[(e + 1) % h]];
                if (cross3(a, b, q) == 0.0 && (q.x - a.x) * (q.x - b.x) <= 0.0 && (q.y - a.y) * (q.y - b.y) <= 0.0) on = 1;
            }
            cnt += on;
        }
        *on_boundary = cnt;
    }
    free(st); free(v);
    return h;
}

/* griddata.rs:437-456 (abs_diff_ne! with the default epsilon f64::EPSILON) */
static int all_on_same_line(const P2* s, int n) {
    if (n < 3) return 1;
    double l1x = s[1].x - s[0].x, l1y = s[1].y - s[0].y;
    for (int i = 2; i < n; ++i) {
        double l2x = s[i].x - s[0].x, l2y = s[i].y - s[0].y;
        if (fabs(fma(l1x, l2y, -(l1y * l2x)) - 0.0) > 2.220446049250313e-16) return 0;
    }
    return 1;
}

/* the bounding polygon of create_voronoi_cells (griddata.rs:290-343): hull vertices and edge midpoints pushed away from their
 * centroid by half the height of the equilateral triangle that has the mean area of the Delaunay triangles.  Returns the number
 * of polygon vertices (2 h). */
static int bounding_polygon(const P2* s, const int* hull, int h, int n_triangles, P2* poly, double* area_hull_out) {
    P2* hp = (P2*)malloc((size_t)h * sizeof(P2));
    for (int i = 0; i < h; ++i) hp[i] = s[hull[h - 1 - i]];  /* convex_hull.iter().rev() */
    double area_hull = 0.0;
    poly_area(hp, h, &area_hull);
    double area_triangle = area_hull / (double)n_triangles;
    double side = sqrt(area_triangle * 4.0 / sqrt(3.0));
    double base_height = sqrt(fma(side / 2.0, -(side / 2.0), side * side));
    int m = 0;
    for (int i = 0; i < h; ++i) {
        int j = (i + 1) % h;
        P2 mid = { hp[i].x + (hp[j].x - hp[i].x) / 2.0, hp[i].y + (hp[j].y - hp[i].y) / 2.0 };
        poly[m++] = hp[i];
        poly[m++] = mid;
    }
    P2 c = { 0.0, 0.0 };
    for (int i = 0; i < m; ++i) { c.x = c.x + poly[i].x; c.y = c.y + poly[i].y; }
    c.x /= (double)m; c.y /= (double)m;
    for (int i = 0; i < m; ++i) {
        double vx = poly[i].x - c.x, vy = poly[i].y - c.y;
        double dist = sqrt(vx * vx + vy * vy);
        double scale = (dist + base_height / 2.0) / dist;
        poly[i].x = fma(scale, vx, c.x);
        poly[i].y = fma(scale, vy, c.y);
    }
    if (area_hull_out) *area_hull_out = area_hull;
    free(hp);
    return m;
}

/* Per-site fluence E_i / area(cell_i) in J/cm^2 (rays_hit_map.rs:432-458); x, y in cm.  NaN for a cell of fewer than 3 vertices. */
int orc_voronoi_site_fluence(const double* x_cm, const double* y_cm, const double* e, uint64_t n64, double* fluence, double* area_hull) {
    const int n = (int)n64;
    if (n < 3) return ORC_ERR_ARG;  /* "Too few points (<3) on hitmap to calculate fluence!" */
    P2* s = (P2*)malloc((size_t)n * sizeof(P2));
    for (int i = 0; i < n; ++i) {
        if (!isfinite(x_cm[i]) || !isfinite(y_cm[i])) { free(s); return ORC_ERR_ARG; }
        s[i].x = x_cm[i]; s[i].y = y_cm[i];
    }
    if (all_on_same_line(s, n)) { free(s); return ORC_ERR_ARG; }
    int* hull = (int*)malloc((size_t)n * sizeof(int));
    int on_b = 0;
    const int h = convex_hull(s, n, hull, &on_b);
    if (h < 3) { free(hull); free(s); return ORC_ERR_ARG; }
    const int n_tri = 2 * n - 2 - on_b;
    P2* poly = (P2*)malloc((size_t)(2 * h) * sizeof(P2));
    const int m = bounding_polygon(s, hull, h, n_tri, poly, area_hull);
    const int cap = m + n + 8;
    P2* a = (P2*)malloc((size_t)cap * sizeof(P2));
    P2* b = (P2*)malloc((size_t)cap * sizeof(P2));
    /* a rectangle that contains the bounding polygon; the polygon's orientation (its interior lies on one side of EVERY edge) */
    double bx0 = poly[0].x, bx1 = poly[0].x, by0 = poly[0].y, by1 = poly[0].y, a2 = 0.0;
    for (int q = 0; q < m; ++q) {
        const int r = (q + 1) % m;
        a2 += poly[q].x * poly[r].y - poly[q].y * poly[r].x;
        bx0 = fmin(bx0, poly[q].x); bx1 = fmax(bx1, poly[q].x); by0 = fmin(by0, poly[q].y); by1 = fmax(by1, poly[q].y);
    }
    const double orient = a2 >= 0.0 ? 1.0 : -1.0;
    int rc = ORC_OK;
    for (int i = 0; i < n && !rc; ++i) {
        /* the Voronoi cell of site i: every other site's half-plane ... */
        int k = 4;
        a[0] = (P2){ bx0, by0 }; a[1] = (P2){ bx1, by0 }; a[2] = (P2){ bx1, by1 }; a[3] = (P2){ bx0, by1 };
        for (int j = 0; j < n && k >= 3; ++j) {
            if (j == i) continue;
            P2 nrm = { s[j].x - s[i].x, s[j].y - s[i].y };
            if (nrm.x == 0.0 && nrm.y == 0.0) continue;  /* a duplicate site has no bisector */
            P2 mid = { s[i].x + 0.5 * nrm.x, s[i].y + 0.5 * nrm.y };
            k = clip_halfplane(a, NULL, k, mid, nrm, j, b, NULL, cap);
            if (k < 0) { rc = ORC_ERR_ARG; break; }
            P2* t = a; a = b; b = t;
        }
        /* ... clipped against the bounding polygon the way voronator does it: Sutherland-Hodgman, one EDGE of the polygon after the
         * other as a half-plane.  For a convex polygon that is the intersection; the reference's polygon is not convex for elongated
         * hulls (the pushed-out edge midpoints stick out), and then this keeps what lies inside every edge's half-plane. */
        for (int q = 0; q < m && k >= 3 && !rc; ++q) {
            const int r = (q + 1) % m;
            const double ex = poly[r].x - poly[q].x, ey = poly[r].y - poly[q].y;
            P2 nrm = { orient * ey, -orient * ex };  /* outward normal */
            k = clip_halfplane(a, NULL, k, poly[q], nrm, -1, b, NULL, cap);
            if (k < 0) { rc = ORC_ERR_ARG; break; }
            P2* t = a; a = b; b = t;
        }
        double ar;
        if (!rc) fluence[i] = (k >= 3 && poly_area(a, k, &ar) == ORC_OK) ? e[i] / ar : NAN;
    }
    free(a); free(b); free(poly); free(hull); free(s);
    return rc;
}

/* Sibson interpolation of site values z at (qx, qy): spade's NaturalNeighbor::interpolate over the sites with finite z. */
static double natural_neighbor(const P2* s, const double* z, int n, const int* hull, int h, P2 q, double big, P2* a, int* aid, P2* b, int* bid,
                               int cap) {
    /* inside, on the boundary of, or outside the convex hull (counter-clockwise) */
    for (int e = 0; e < h; ++e) {
        P2 p0 = s[hull[e]], p1 = s[hull[(e + 1) % h]];
        double c = cross3(p0, p1, q);
        if (c < 0.0) return NAN;  /* OutsideOfConvexHull -> None */
        if (c == 0.0) {           /* on the line of a hull edge */
            double ex = p1.x - p0.x, ey = p1.y - p0.y;
            double t = ((q.x - p0.x) * ex + (q.y - p0.y) * ey) / (ex * ex + ey * ey);
            if (t < 0.0 || t > 1.0) return NAN;
            if (t == 0.0) return z[hull[e]];
            if (t == 1.0) return z[hull[(e + 1) % h]];
            return (1.0 - t) * z[hull[e]] + t * z[hull[(e + 1) % h]];  /* two-point interpolation along a hull edge */
        }
    }
    for (int i = 0; i < n; ++i)
        if (s[i].x == q.x && s[i].y == q.y) return z[i];  /* OnVertex */
    /* the cell q would get: a large square around q clipped by the bisectors with all sites; edge ids = its natural neighbours */
    int k = 4;
    a[0] = (P2){ q.x - big, q.y - big }; a[1] = (P2){ q.x + big, q.y - big }; a[2] = (P2){ q.x + big, q.y + big }; a[3] = (P2){ q.x - big, q.y + big };
    for (int t = 0; t < 4; ++t) aid[t] = -1;
    for (int j = 0; j < n && k >= 3; ++j) {
        P2 nrm = { s[j].x - q.x, s[j].y - q.y };
        P2 mid = { q.x + 0.5 * nrm.x, q.y + 0.5 * nrm.y };
        k = clip_halfplane(a, aid, k, mid, nrm, j, b, bid, cap);
        if (k < 0) return NAN;
        P2* t = a; a = b; b = t;
        int* ti = aid; aid = bid; bid = ti;
    }
    if (k < 3) return NAN;
    double total = 0.0, acc = 0.0;
    int nn[64], n_nn = 0;
    for (int t = 0; t < k; ++t) {
        int id = aid[t], seen = 0;
        if (id < 0) return NAN;  /* the cell reaches the far square: q is (numerically) outside the hull */
        for (int u = 0; u < n_nn; ++u) seen |= nn[u] == id;
        if (!seen && n_nn < 64) nn[n_nn++] = id;
    }
    P2 c0[128], c1[128];
    for (int u = 0; u < n_nn; ++u) {
        const int i = nn[u];
        int m = k;
        P2 *pa = c0, *pb = c1;
        memcpy(pa, a, (size_t)k * sizeof(P2));
        for (int w = 0; w < n_nn && m >= 3; ++w) {
            if (w == u) continue;
            const int j = nn[w];
            P2 nrm = { s[j].x - s[i].x, s[j].y - s[i].y };
            if (nrm.x == 0.0 && nrm.y == 0.0) continue;
            P2 mid = { s[i].x + 0.5 * nrm.x, s[i].y + 0.5 * nrm.y };
            m = clip_halfplane(pa, NULL, m, mid, nrm, j, pb, NULL, 128);
            if (m < 0) return NAN;
            P2* t = pa; pa = pb; pb = t;
        }
        double ar = 0.0;
        if (m >= 3) poly_area(pa, m, &ar);
        total += ar;
        acc += ar * z[i];
    }
    return total > 0.0 ? acc / total : NAN;
}

/* RaysHitMap::calc_fluence_with_voronoi (rays_hit_map.rs:407-521).  hits in metres / joules; ranges: NULL (data limits) or
 * {ax1.start, ax1.end, ax2.start, ax2.end} in metres.  map: nx * nx doubles (BOTH axes take nr_of_points.0 points, :470 and :489),
 * column-major like the reference's DMatrix (element (row = y index, column = x index) at x_index * nx + y_index), in J/m^2 (the
 * SI value of the reference's J/cm^2 quantity); NaN outside the convex hull of the hits.  lrtb_out = (left, right, top, bottom) in
 * metres as FluenceData keeps them. */
int orc_fluence_voronoi(const OrcHit* hits, uint64_t n64, uint64_t nx, const double* ranges, double* map, double lrtb_out[4]) {
    const int n = (int)n64;
    if (n < 3) return ORC_ERR_ARG;
    double* x = (double*)malloc((size_t)n * 4 * sizeof(double));
    double *y = x + n, *e = y + n, *fl = e + n;
    for (int i = 0; i < n; ++i) { x[i] = hits[i].x / 1.0E-2; y[i] = hits[i].y / 1.0E-2; e[i] = hits[i].e; }  /* get::<centimeter>() */
    int rc = orc_voronoi_site_fluence(x, y, e, n64, fl, NULL);
    if (rc) { free(x); return rc; }
    double x0, x1, y0, y1;
    if (ranges) { x0 = ranges[0] / 1.0E-2; x1 = ranges[1] / 1.0E-2; y0 = ranges[2] / 1.0E-2; y1 = ranges[3] / 1.0E-2; }
    else {  /* AxLims::finite_from_dvector: minimum and maximum of the finite values */
        x0 = y0 = INFINITY; x1 = y1 = -INFINITY;
        for (int i = 0; i < n; ++i) { x0 = fmin(x0, x[i]); x1 = fmax(x1, x[i]); y0 = fmin(y0, y[i]); y1 = fmax(y1, y[i]); }
    }
    double* ax = (double*)malloc((size_t)nx * 2 * sizeof(double));
    double* ay = ax + nx;
    if (orc_linspace(x0, x1, nx, ax) || orc_linspace(y0, y1, nx, ay)) { free(ax); free(x); return ORC_ERR_ARG; }
    /* sites with a finite value only (griddata.rs:412) */
    P2* s = (P2*)malloc((size_t)n * sizeof(P2));
    double* z = (double*)malloc((size_t)n * sizeof(double));
    int m = 0;
    double ext = 0.0;
    for (int i = 0; i < n; ++i)
        if (isfinite(fl[i])) { s[m].x = x[i]; s[m].y = y[i]; z[m] = fl[i]; ++m; }
    for (int i = 0; i < m; ++i) { ext = fmax(ext, fabs(s[i].x)); ext = fmax(ext, fabs(s[i].y)); }
    int* hull = (int*)malloc((size_t)(m > 0 ? m : 1) * sizeof(int));
    const int h = m >= 3 ? convex_hull(s, m, hull, NULL) : 0;
    const int cap = m + 16;
    P2* a = (P2*)malloc((size_t)cap * 2 * sizeof(P2));
    int* aid = (int*)malloc((size_t)cap * 2 * sizeof(int));
    const double big = 1.0e3 * (fmax(x1 - x0, y1 - y0) + ext + 1.0);
    for (uint64_t xi = 0; xi < nx; ++xi)
        for (uint64_t yi = 0; yi < nx; ++yi) {
            double v = NAN;
            if (h >= 3) v = natural_neighbor(s, z, m, hull, h, (P2){ ax[xi], ay[yi] }, big, a, aid, a + cap, aid + cap, cap);
            map[xi * nx + yi] = v * 1.0E4;  /* J_per_cm2!(v): the SI value */
        }
    if (lrtb_out) { lrtb_out[0] = x0 * 1.0E-2; lrtb_out[1] = x1 * 1.0E-2; lrtb_out[2] = y1 * 1.0E-2; lrtb_out[3] = y0 * 1.0E-2; }
    if (ranges && lrtb_out) { lrtb_out[0] = ranges[0]; lrtb_out[1] = ranges[1]; lrtb_out[2] = ranges[3]; lrtb_out[3] = ranges[2]; }
    free(aid); free(a); free(hull); free(z); free(s); free(ax); free(x);
    return ORC_OK;
}

/* RaysHitMap::calc_fluence_with_helper_rays (rays_hit_map.rs:627-722): hits whose `e` is a FLUENCE in J/m^2 (FluenceHitPoint).
 * The sites of create_voronoi_cells take the fluences in J/cm^2, the sites voronator adds take 0 (:700-701), natural-neighbour
 * interpolation over all of them (griddata.rs:387-436).  voronator 0.2 (VoronoiDiagram::with_bounding_polygon -> helper_points)
 * appends FOUR sites: left, right, below and above the box of the clip polygon, one box width / height away -- restated from
 * the crate as remembered, its source is not under /root/reference: parity unpinned beyond the reference's known answer
 * (hit_map/mod.rs:1030-1068, a value AT a site).  map, ranges, lrtb_out as orc_fluence_voronoi. */
int orc_fluence_helper_rays(const OrcHit* hits, uint64_t n64, uint64_t nx, const double* ranges, double* map, double lrtb_out[4]) {
    const int n = (int)n64;
    if (n < 3) return ORC_ERR_ARG;  /* "Too few points (<3) on hitmap to calculate fluence!" */
    P2* s = (P2*)malloc((size_t)(n + 4) * sizeof(P2));
    double* z = (double*)malloc((size_t)(n + 4) * sizeof(double));
    for (int i = 0; i < n; ++i) {
        s[i].x = hits[i].x / 1.0E-2; s[i].y = hits[i].y / 1.0E-2; z[i] = hits[i].e / 1.0E4;  /* get::<centimeter>(), get::<joule_per_square_centimeter>() */
        if (!isfinite(s[i].x) || !isfinite(s[i].y)) { free(z); free(s); return ORC_ERR_ARG; }
    }
    double x0, x1, y0, y1;
    if (ranges) { x0 = ranges[0] / 1.0E-2; x1 = ranges[1] / 1.0E-2; y0 = ranges[2] / 1.0E-2; y1 = ranges[3] / 1.0E-2; }
    else {
        x0 = y0 = INFINITY; x1 = y1 = -INFINITY;
        for (int i = 0; i < n; ++i) { x0 = fmin(x0, s[i].x); x1 = fmax(x1, s[i].x); y0 = fmin(y0, s[i].y); y1 = fmax(y1, s[i].y); }
    }
    if (all_on_same_line(s, n)) { free(z); free(s); return ORC_ERR_ARG; }
    int* hull = (int*)malloc((size_t)(n + 4) * sizeof(int));
    int on_b = 0;
    int h = convex_hull(s, n, hull, &on_b);
    if (h < 3) { free(hull); free(z); free(s); return ORC_ERR_ARG; }
    P2* poly = (P2*)malloc((size_t)(2 * h) * sizeof(P2));
    const int m = bounding_polygon(s, hull, h, 2 * n - 2 - on_b, poly, NULL);
    double bx0 = poly[0].x, bx1 = poly[0].x, by0 = poly[0].y, by1 = poly[0].y;
    for (int q = 1; q < m; ++q) { bx0 = fmin(bx0, poly[q].x); bx1 = fmax(bx1, poly[q].x); by0 = fmin(by0, poly[q].y); by1 = fmax(by1, poly[q].y); }
    const double w = bx1 - bx0, hg = by1 - by0;
    s[n] = (P2){ bx0 - w, by0 + hg / 2.0 }; s[n + 1] = (P2){ bx1 + w, by0 + hg / 2.0 };
    s[n + 2] = (P2){ bx0 + w / 2.0, by0 - hg }; s[n + 3] = (P2){ bx0 + w / 2.0, by1 + hg };
    for (int k = 0; k < 4; ++k) z[n + k] = 0.0;
    /* finite values only (griddata.rs:412) */
    int mm = 0;
    for (int i = 0; i < n + 4; ++i) if (isfinite(z[i])) { s[mm] = s[i]; z[mm] = z[i]; ++mm; }
    double* ax = (double*)malloc((size_t)nx * 2 * sizeof(double));
    double* ay = ax + nx;
    if (orc_linspace(x0, x1, nx, ax) || orc_linspace(y0, y1, nx, ay)) { free(ax); free(poly); free(hull); free(z); free(s); return ORC_ERR_ARG; }
    double ext = 0.0;
    for (int i = 0; i < mm; ++i) { ext = fmax(ext, fabs(s[i].x)); ext = fmax(ext, fabs(s[i].y)); }
    h = mm >= 3 ? convex_hull(s, mm, hull, NULL) : 0;
    const int cap = mm + 16;
    P2* a = (P2*)malloc((size_t)cap * 2 * sizeof(P2));
    int* aid = (int*)malloc((size_t)cap * 2 * sizeof(int));
    const double big = 1.0e3 * (fmax(x1 - x0, y1 - y0) + ext + 1.0);
    for (uint64_t xi = 0; xi < nx; ++xi)
        for (uint64_t yi = 0; yi < nx; ++yi) {
            double v = NAN;
            if (h >= 3) v = natural_neighbor(s, z, mm, hull, h, (P2){ ax[xi], ay[yi] }, big, a, aid, a + cap, aid + cap, cap);
            map[xi * nx + yi] = v * 1.0E4;
        }
    if (lrtb_out) {
        if (ranges) { lrtb_out[0] = ranges[0]; lrtb_out[1] = ranges[1]; lrtb_out[2] = ranges[3]; lrtb_out[3] = ranges[2]; }
        else { lrtb_out[0] = x0 * 1.0E-2; lrtb_out[1] = x1 * 1.0E-2; lrtb_out[2] = y1 * 1.0E-2; lrtb_out[3] = y0 * 1.0E-2; }
    }
    free(aid); free(a); free(ax); free(poly); free(hull); free(z); free(s);
    return ORC_OK;
}
