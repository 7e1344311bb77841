// opm_voronoi.cu -- the Voronoi fluence estimator on the device, sm_100a.
//
// RaysHitMap::calc_fluence_with_voronoi (surface/hit_map/rays_hit_map.rs:407-521) with create_voronoi_cells (utils/griddata.rs:
// 265-345) and interpolate_3d_triangulated_scatter_data (:387-436): the reference's DEFAULT estimator for FluenceDetector
// (nodes/fluence_detector/mod.rs:61) and for the ghost-focus LIDT check (analyzers/ghostfocus.rs:61).  The reference builds a
// Delaunay triangulation twice (voronator / delaunator for the cells, spade for the natural-neighbour interpolation).  Here no
// triangulation is built at all -- both steps are independent per output element and use only LOCAL neighbourhoods:
//
//   cell of site i      = a rectangle around the data clipped by the half-planes "closer to s_i than to s_j" of the sites j in
//                         growing rings of a uniform grid around s_i, until no unseen site can reach the cell any more (a site
//                         at distance d cuts the cell only if d / 2 < the cell's largest vertex distance), then clipped by the
//                         reference's bounding polygon (the convex hull pushed outwards); fluence_i = E_i / area.  One thread
//                         per site, the polygon in local memory.
//   value at map point q = Sibson's natural-neighbour interpolation: the cell q WOULD get (same ring search; the edges remember
//                         which site made them = the natural neighbours), split among the neighbours by clipping it with their
//                         mutual bisectors; value = sum(share_i * fluence_i) / sum(share_i).  NaN outside the convex hull, the
//                         site's value on a site, linear along a hull edge (spade's NaturalNeighbor::interpolate).  One thread per
//                         map point.
//
// The host side (opm_api.cu) packs the hits, builds the uniform grid and the convex hull of the sites (O(n log n) bookkeeping
// on the hit positions; all clipping runs here).  Arithmetic in centimetres / J/cm^2 like the reference; the map is stored in
// J/m^2.  Parity: against oracle/opm_voronoi.c (brute-force half-plane clipping over ALL sites), which the reference pins only
// through hit_map/mod.rs:942-982 -- see the oracle's header for what stays unpinned.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#define OPM_VORONOI_HOST
#include "opm_kernels.h"

namespace opm {

constexpr int kVorMaxV = 64;   // vertices of a polygon being clipped
constexpr int kVorMaxNN = 32;  // natural neighbours of a map point

struct VP2 { double x, y; };

// Sutherland-Hodgman against d(v) = (v - m) . nrm <= 0.  ID: edge k (vertex k -> k + 1) remembers the site whose bisector made it.
template <bool ID>
__device__ __forceinline__ int vor_clip(const VP2* in, const int* in_id, int n, double mx, double my, double nx, double ny, int id, VP2* out,
                                        int* out_id) {
    int k = 0;
    double da = (in[0].x - mx) * nx + (in[0].y - my) * ny;
    for (int a = 0; a < n; ++a) {
        const int b = a + 1 == n ? 0 : a + 1;
        const double db = (in[b].x - mx) * nx + (in[b].y - my) * ny;
        const bool ia = da <= 0.0, ib = db <= 0.0;
        if (ia) {
            if (k < kVorMaxV) { out[k] = in[a]; if (ID) out_id[k] = in_id[a]; }
            ++k;
        }
        if (ia != ib) {
            const double t = da / (da - db);
            if (k < kVorMaxV) {
                out[k].x = in[a].x + t * (in[b].x - in[a].x);
                out[k].y = in[a].y + t * (in[b].y - in[a].y);
                if (ID) out_id[k] = ia ? id : in_id[a];
            }
            ++k;
        }
        da = db;
    }
    return k;
}
// does the half-plane cut anything?  (most bisectors of a ring do not: skip the copy)
__device__ __forceinline__ bool vor_cuts(const VP2* p, int n, double mx, double my, double nx, double ny) {
    for (int a = 0; a < n; ++a)
        if ((p[a].x - mx) * nx + (p[a].y - my) * ny > 0.0) return true;
    return false;
}
// griddata.rs:112-134
__device__ __forceinline__ double vor_area(const VP2* p, int n) {
    double area = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = i + 1 == n ? 0 : i + 1;
        area += p[i].x * p[j].y;
        area -= p[i].y * p[j].x;
    }
    return fabs(area) / 2.0;
}
__device__ __forceinline__ double vor_rmax2(const VP2* p, int n, double cx, double cy) {
    double r2 = 0.0;
    for (int a = 0; a < n; ++a) { const double dx = p[a].x - cx, dy = p[a].y - cy; r2 = fmax(r2, dx * dx + dy * dy); }
    return r2;
}

// Clips the polygon (cur, n) around the point (cx, cy) by the bisectors with the sites of the uniform grid, ring by ring, until no
// unseen site can cut it.  skip: a site index to leave out (the site itself), -1 for a map point.  Returns the vertex count
// (> kVorMaxV: overflow).  The two buffers alternate; *which tells where the result lies.
template <bool ID>
__device__ int vor_ring_clip(const VorGrid& G, const double* __restrict__ sx, const double* __restrict__ sy, double cx, double cy, int skip,
                             VP2* b0, int* i0, VP2* b1, int* i1, int n, int* which) {
    VP2* cur = b0; VP2* nxt = b1;
    int* cid = i0; int* nid = i1;
    int ci = (int)floor((cx - G.x0) * G.inv_cs), cj = (int)floor((cy - G.y0) * G.inv_cs);
    ci = min(max(ci, 0), G.gx - 1); cj = min(max(cj, 0), G.gy - 1);
    const int rmax = max(max(ci, G.gx - 1 - ci), max(cj, G.gy - 1 - cj));
    for (int r = 0; r <= rmax && n >= 3 && n <= kVorMaxV; ++r) {
        const int i_lo = ci - r, i_hi = ci + r, j_lo = cj - r, j_hi = cj + r;
        for (int gi = max(i_lo, 0); gi <= min(i_hi, G.gx - 1); ++gi) {
            const bool edge_col = gi == i_lo || gi == i_hi;
            const int step = edge_col ? 1 : max(j_hi - j_lo, 1);  // inner columns: only the two end rows belong to the ring
            for (int gj = j_lo; gj <= j_hi; gj += step) {
                if (gj < 0 || gj >= G.gy) continue;
                const int c = gi * G.gy + gj;
                for (int k = G.cell_start[c]; k < G.cell_start[c + 1]; ++k) {
                    const int j = G.cell_sites[k];
                    if (j == skip) continue;
                    const double nx = sx[j] - cx, ny = sy[j] - cy;
                    if (nx == 0.0 && ny == 0.0) continue;  // a duplicate site has no bisector
                    const double mx = cx + 0.5 * nx, my = cy + 0.5 * ny;
                    if (!vor_cuts(cur, n, mx, my, nx, ny)) continue;
                    n = vor_clip<ID>(cur, cid, n, mx, my, nx, ny, j, nxt, nid);
                    VP2* t = cur; cur = nxt; nxt = t;
                    int* ti = cid; cid = nid; nid = ti;
                    if (n < 3 || n > kVorMaxV) break;
                }
                if (n < 3 || n > kVorMaxV) break;
            }
            if (n < 3 || n > kVorMaxV) break;
        }
        if (n < 3 || n > kVorMaxV) break;
        // every site of the rings beyond r is farther than r cells away; it cuts the polygon only if half that distance is
        // inside the polygon's reach
        const double reach = (double)r * G.cs;
        if (reach * reach >= 4.0 * vor_rmax2(cur, n, cx, cy)) break;
    }
    *which = cur == b0 ? 0 : 1;
    return n;
}

// ---- cells: fluence_i = E_i / area(Voronoi cell of site i inside the bounding polygon), J/cm^2 (rays_hit_map.rs:432-458)
__global__ void __launch_bounds__(128) voronoi_cells_kernel(int n, const double* __restrict__ sx, const double* __restrict__ sy,
                                                            const double* __restrict__ e, VorGrid G, VorBound B, double* __restrict__ fluence,
                                                            unsigned* status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    VP2 p0[kVorMaxV], p1[kVorMaxV];
    const double cx = sx[i], cy = sy[i];
    p0[0] = {B.x0, B.y0}; p0[1] = {B.x1, B.y0}; p0[2] = {B.x1, B.y1}; p0[3] = {B.x0, B.y1};  // a rectangle that contains the bounding polygon
    int which = 0;
    int k = vor_ring_clip<false>(G, sx, sy, cx, cy, i, p0, nullptr, p1, nullptr, 4, &which);
    VP2* cur = which ? p1 : p0; VP2* nxt = which ? p0 : p1;
    if (k >= 3 && k <= kVorMaxV) {
        // the bounding polygon (convex): only cells that reach beyond its inscribed circle can be cut by it
        bool inside = true;
        for (int a = 0; a < k; ++a) { const double dx = cur[a].x - B.cx, dy = cur[a].y - B.cy; inside = inside && dx * dx + dy * dy < B.r_in2; }
        if (!inside)
            for (int q = 0; q < B.m && k >= 3 && k <= kVorMaxV; ++q) {
                const double mx = B.px[q], my = B.py[q], nx = B.nx[q], ny = B.ny[q];  // a point of edge q and its OUTWARD normal
                if (!vor_cuts(cur, k, mx, my, nx, ny)) continue;
                k = vor_clip<false>(cur, nullptr, k, mx, my, nx, ny, -1, nxt, nullptr);
                VP2* t = cur; cur = nxt; nxt = t;
            }
    }
    if (k > kVorMaxV) { atomicOr(status, 1u << OPM_ERR_CAPACITY); fluence[i] = NAN; return; }
    fluence[i] = k >= 3 ? e[i] / vor_area(cur, k) : NAN;  // "polygon could not be created": the value stays NaN
}

// ---- natural-neighbour interpolation onto the map (griddata.rs:387-436); map in J/m^2, column-major (x index * ny + y index)
__global__ void __launch_bounds__(128) voronoi_interp_kernel(int nx_pts, int ny_pts, double ax0, double ax_bin, double ay0, double ay_bin, int n,
                                                             const double* __restrict__ sx, const double* __restrict__ sy,
                                                             const double* __restrict__ sz, VorGrid G, VorHull H, double big,
                                                             double* __restrict__ map, unsigned* status) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nx_pts * ny_pts) return;
    const int xi = t / ny_pts, yi = t % ny_pts;
    // linspace (griddata.rs:211-231): start + i * bin
    const double qx = nx_pts < 2 ? ax0 : ax0 + (double)xi * ax_bin, qy = ny_pts < 2 ? ay0 : ay0 + (double)yi * ay_bin;
    double val = NAN;
    bool decided = false;
    for (int eI = 0; eI < H.h && !decided; ++eI) {  // inside, on the boundary of, or outside the convex hull (counter-clockwise)
        const int e1 = eI + 1 == H.h ? 0 : eI + 1;
        const double ax = H.x[eI], ay = H.y[eI], bx = H.x[e1], by = H.y[e1];
        const double c = (bx - ax) * (qy - ay) - (by - ay) * (qx - ax);
        if (c < 0.0) decided = true;  // OutsideOfConvexHull -> None -> NaN
        else if (c == 0.0) {
            const double ex = bx - ax, ey = by - ay;
            const double u = ((qx - ax) * ex + (qy - ay) * ey) / (ex * ex + ey * ey);
            if (u >= 0.0 && u <= 1.0) val = u == 0.0 ? H.z[eI] : (u == 1.0 ? H.z[e1] : (1.0 - u) * H.z[eI] + u * H.z[e1]);
            decided = true;
        }
    }
    if (!decided) {  // OnVertex: a site exactly at the map point
        int ci = (int)floor((qx - G.x0) * G.inv_cs), cj = (int)floor((qy - G.y0) * G.inv_cs);
        ci = min(max(ci, 0), G.gx - 1); cj = min(max(cj, 0), G.gy - 1);
        for (int di = -1; di <= 1 && !decided; ++di)
            for (int dj = -1; dj <= 1 && !decided; ++dj) {
                const int gi = ci + di, gj = cj + dj;
                if (gi < 0 || gj < 0 || gi >= G.gx || gj >= G.gy) continue;
                const int c = gi * G.gy + gj;
                for (int k = G.cell_start[c]; k < G.cell_start[c + 1]; ++k) {
                    const int j = G.cell_sites[k];
                    if (sx[j] == qx && sy[j] == qy) { val = sz[j]; decided = true; break; }
                }
            }
    }
    if (!decided) {
        VP2 p0[kVorMaxV], p1[kVorMaxV];
        int d0[kVorMaxV], d1[kVorMaxV];
        p0[0] = {qx - big, qy - big}; p0[1] = {qx + big, qy - big}; p0[2] = {qx + big, qy + big}; p0[3] = {qx - big, qy + big};
        d0[0] = d0[1] = d0[2] = d0[3] = -1;
        int which = 0;
        const int k = vor_ring_clip<true>(G, sx, sy, qx, qy, -1, p0, d0, p1, d1, 4, &which);
        const VP2* cell = which ? p1 : p0;
        const int* cid = which ? d1 : d0;
        VP2* w0 = which ? p0 : p1;  // the other buffer and ...
        if (k > kVorMaxV) atomicOr(status, 1u << OPM_ERR_CAPACITY);
        else if (k >= 3) {
            int nn[kVorMaxNN], n_nn = 0;
            bool open = false;
            for (int a = 0; a < k; ++a) {
                const int id = cid[a];
                if (id < 0) { open = true; break; }  // the cell reaches the far square: numerically outside the hull
                bool seen = false;
                for (int u = 0; u < n_nn; ++u) seen = seen || nn[u] == id;
                if (!seen) { if (n_nn < kVorMaxNN) nn[n_nn++] = id; else { open = true; atomicOr(status, 1u << OPM_ERR_CAPACITY); } }
            }
            if (!open) {
                VP2 w1[kVorMaxV];  // ... a third one: the share of one neighbour at a time
                double total = 0.0, acc = 0.0;
                for (int u = 0; u < n_nn; ++u) {
                    const int i = nn[u];
                    int m = k;
                    VP2* pa = w0; VP2* pb = w1;
                    for (int a = 0; a < k; ++a) pa[a] = cell[a];
                    for (int w = 0; w < n_nn && m >= 3 && m <= kVorMaxV; ++w) {
                        if (w == u) continue;
                        const int j = nn[w];
                        const double nx = sx[j] - sx[i], ny = sy[j] - sy[i];
                        if (nx == 0.0 && ny == 0.0) continue;
                        const double mx = sx[i] + 0.5 * nx, my = sy[i] + 0.5 * ny;
                        if (!vor_cuts(pa, m, mx, my, nx, ny)) continue;
                        m = vor_clip<false>(pa, nullptr, m, mx, my, nx, ny, j, pb, nullptr);
                        VP2* tt = pa; pa = pb; pb = tt;
                    }
                    const double ar = (m >= 3 && m <= kVorMaxV) ? vor_area(pa, m) : 0.0;
                    total += ar;
                    acc += ar * sz[i];
                }
                if (total > 0.0) val = acc / total;
            }
        }
    }
    map[(size_t)xi * ny_pts + yi] = val * 1.0E4;  // J_per_cm2!(v): the SI value
}
// hits (metres) -> sites in centimetres (get::<centimeter>(): a division by 1e-2, rays_hit_map.rs:422-424)
__global__ void voronoi_to_cm_kernel(int n, double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { x[i] = x[i] / 1.0E-2; y[i] = y[i] / 1.0E-2; }
}
// map += part (HitMap::calc_combined_fluence_with_voronoi, hit_map/mod.rs:246-255: NaN propagates)
__global__ void voronoi_accumulate_kernel(uint64_t bins, double* __restrict__ map, const double* __restrict__ part) {
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < bins; k += (uint64_t)gridDim.x * blockDim.x) map[k] += part[k];
}

#define ST ((cudaStream_t)stream)
int k_voronoi_to_cm(int n, double* x, double* y, void* stream) {
    if (n) voronoi_to_cm_kernel<<<(n + 255) / 256, 256, 0, ST>>>(n, x, y);
    return (int)cudaGetLastError();
}
int k_voronoi_cells(int n, const double* sx, const double* sy, const double* e, const VorGrid& G, const VorBound& B, double* fluence,
                    unsigned* status, void* stream) {
    voronoi_cells_kernel<<<(n + 127) / 128, 128, 0, ST>>>(n, sx, sy, e, G, B, fluence, status);
    return (int)cudaGetLastError();
}
int k_voronoi_interp(int nx_pts, int ny_pts, double ax0, double ax_bin, double ay0, double ay_bin, int n, const double* sx, const double* sy,
                     const double* sz, const VorGrid& G, const VorHull& H, double big, double* map, unsigned* status, void* stream) {
    const int pts = nx_pts * ny_pts;
    voronoi_interp_kernel<<<(pts + 127) / 128, 128, 0, ST>>>(nx_pts, ny_pts, ax0, ax_bin, ay0, ay_bin, n, sx, sy, sz, G, H, big, map, status);
    return (int)cudaGetLastError();
}
int k_voronoi_accumulate(uint64_t bins, double* map, const double* part, int sms, void* stream) {
    voronoi_accumulate_kernel<<<sms * 4, 256, 0, ST>>>(bins, map, part);
    return (int)cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------------
// host bookkeeping on the hit positions (O(n log n)): convex hull, the reference's bounding polygon, the uniform grid
// ---------------------------------------------------------------------------------------------------------------------------
static double vh_cross(const double* x, const double* y, int o, int a, int b) { return (x[a] - x[o]) * (y[b] - y[o]) - (y[a] - y[o]) * (x[b] - x[o]); }

// Andrew's monotone chain, counter-clockwise, strict vertices; *on_boundary (may be NULL) = distinct points on the hull's boundary
void vor_convex_hull(const double* x, const double* y, int n, std::vector<int>& hull, int* on_boundary) {
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return x[a] != x[b] ? x[a] < x[b] : (y[a] != y[b] ? y[a] < y[b] : a < b); });
    std::vector<int> v;
    v.reserve(n);
    for (int i : idx)
        if (v.empty() || x[i] != x[v.back()] || y[i] != y[v.back()]) v.push_back(i);
    const int m = (int)v.size();
    std::vector<int> st(2 * m + 2);
    int k = 0;
    for (int i = 0; i < m; ++i) {
        while (k >= 2 && vh_cross(x, y, st[k - 2], st[k - 1], v[i]) <= 0.0) --k;
        st[k++] = v[i];
    }
    for (int i = m - 2, t = k + 1; i >= 0; --i) {
        while (k >= t && vh_cross(x, y, st[k - 2], st[k - 1], v[i]) <= 0.0) --k;
        st[k++] = v[i];
    }
    const int h = k > 1 ? k - 1 : k;
    hull.assign(st.begin(), st.begin() + h);
    if (on_boundary) {
        // a point lies on the boundary when it lies on one of the h edges; candidates are few: test against the edge whose angular
        // sector contains it would need the centroid -- with h small against n the direct test over points x edges is kept for
        // hulls of up to 64 vertices, larger hulls (round beams: no three boundary points are collinear in practice) count their
        // vertices only
        int cnt = h;
        if (h <= 64) {
            cnt = 0;
            for (int i : v) {
                bool on = false;
                for (int e = 0; e < h && !on; ++e) {
                    const int a = hull[e], b = hull[(e + 1) % h];
                    on = vh_cross(x, y, a, b, i) == 0.0 && (x[i] - x[a]) * (x[i] - x[b]) <= 0.0 && (y[i] - y[a]) * (y[i] - y[b]) <= 0.0;
                }
                cnt += on ? 1 : 0;
            }
        }
        *on_boundary = cnt;
    }
}
// utils/griddata.rs:437-456
bool vor_all_on_same_line(const double* x, const double* y, int n) {
    if (n < 3) return true;
    const double l1x = x[1] - x[0], l1y = y[1] - y[0];
    for (int i = 2; i < n; ++i) {
        const double l2x = x[i] - x[0], l2y = y[i] - y[0];
        if (std::fabs(std::fma(l1x, l2y, -(l1y * l2x))) > 2.220446049250313e-16) return false;
    }
    return true;
}
// create_voronoi_cells, utils/griddata.rs:290-343: polygon vertices (2 h) in the reference's order
void vor_bounding_polygon(const double* x, const double* y, const std::vector<int>& hull, int n_triangles, std::vector<double>& px,
                          std::vector<double>& py, double* cx_out, double* cy_out) {
    const int h = (int)hull.size();
    std::vector<double> hx(h), hy(h);
    for (int i = 0; i < h; ++i) { hx[i] = x[hull[h - 1 - i]]; hy[i] = y[hull[h - 1 - i]]; }  // convex_hull.iter().rev()
    double area = 0.0;
    for (int i = 0; i < h; ++i) { const int j = (i + 1) % h; area += hx[i] * hy[j]; area -= hy[i] * hx[j]; }
    const double area_hull = std::fabs(area) / 2.0;
    const double area_triangle = area_hull / (double)n_triangles;
    const double side = std::sqrt(area_triangle * 4.0 / std::sqrt(3.0));
    const double base_height = std::sqrt(std::fma(side / 2.0, -(side / 2.0), side * side));
    px.clear(); py.clear();
    for (int i = 0; i < h; ++i) {
        const int j = (i + 1) % h;
        px.push_back(hx[i]); py.push_back(hy[i]);
        px.push_back(hx[i] + (hx[j] - hx[i]) / 2.0); py.push_back(hy[i] + (hy[j] - hy[i]) / 2.0);
    }
    const int m = (int)px.size();
    double cx = 0.0, cy = 0.0;
    for (int i = 0; i < m; ++i) { cx = cx + px[i]; cy = cy + py[i]; }
    cx /= (double)m; cy /= (double)m;
    for (int i = 0; i < m; ++i) {
        const double vx = px[i] - cx, vy = py[i] - cy;
        const double dist = std::sqrt(vx * vx + vy * vy);
        const double scale = (dist + base_height / 2.0) / dist;
        px[i] = std::fma(scale, vx, cx);
        py[i] = std::fma(scale, vy, cy);
    }
    *cx_out = cx; *cy_out = cy;
}
// uniform grid with about two sites per cell: counting sort of the site indices by cell
void vor_build_grid(const double* x, const double* y, int n, VorGrid& G, std::vector<int>& cell_start, std::vector<int>& cell_sites) {
    double x0 = x[0], x1 = x[0], y0 = y[0], y1 = y[0];
    for (int i = 1; i < n; ++i) { x0 = std::min(x0, x[i]); x1 = std::max(x1, x[i]); y0 = std::min(y0, y[i]); y1 = std::max(y1, y[i]); }
    const double w = x1 - x0, hgt = y1 - y0;
    double cs = std::sqrt(std::max(w, 1e-300) * std::max(hgt, 1e-300) / std::max(n / 2.0, 1.0));
    if (!(cs > 0.0) || !std::isfinite(cs)) cs = std::max(std::max(w, hgt), 1.0);
    if (getenv("OPM_B200_VORONOI_BRUTE")) cs = 2.0 * std::max(std::max(w, hgt), 1e-300);  // diagnosis: one cell, every site is visited
    int gx = (int)std::floor(w / cs) + 1, gy = (int)std::floor(hgt / cs) + 1;
    while ((double)gx * (double)gy > 4.0 * n + 64.0) { cs *= 1.5; gx = (int)std::floor(w / cs) + 1; gy = (int)std::floor(hgt / cs) + 1; }
    G.x0 = x0; G.y0 = y0; G.cs = cs; G.inv_cs = 1.0 / cs; G.gx = gx; G.gy = gy;
    cell_start.assign((size_t)gx * gy + 1, 0);
    std::vector<int> cell(n);
    for (int i = 0; i < n; ++i) {
        int ci = (int)std::floor((x[i] - x0) * G.inv_cs), cj = (int)std::floor((y[i] - y0) * G.inv_cs);
        ci = std::min(std::max(ci, 0), gx - 1); cj = std::min(std::max(cj, 0), gy - 1);
        cell[i] = ci * gy + cj;
        cell_start[cell[i] + 1]++;
    }
    for (size_t c = 0; c < (size_t)gx * gy; ++c) cell_start[c + 1] += cell_start[c];
    cell_sites.assign(n, 0);
    std::vector<int> fill(cell_start.begin(), cell_start.end() - 1);
    for (int i = 0; i < n; ++i) cell_sites[fill[cell[i]]++] = i;
}

}  // namespace opm
