// opm_scene.cpp -- host-side mirror of the reference's node / analyzer layer (see include/opossum_b200_scene.h).
//
// Pure host bookkeeping, O(#nodes): it compiles optic nodes into surface descriptors and `Rays` op programs
// and drives them through the public C ABI (opossum_b200.h) exactly as a Rust host would through FFI.  No ray
// arithmetic happens here.  Reference file:line citations are relative to opossum/src/.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/opossum_b200_scene.h"

namespace {

struct HitEntry { uint64_t uuid; OpmHitMap* map; };

// One bundle of the ghost-focus analysis (struct Rays: uuid, parent_id; rays.rs:62-73)
struct Bundle {
    OpmRays* rays = nullptr;
    uint64_t uuid = 0;
    uint64_t parent = 0;
    // ghost focus: a clone of a cached bundle that has not been written yet -- its first launch reads `lazy_src` and writes `rays`
    // (opm_scene_gf.inc); lazy_limits: the clone still owes the ray limits of the node it was injected behind
    std::shared_ptr<Bundle> lazy_src;
    bool lazy_limits = false;
};

// surface/optic_surface.rs:29-43
struct SurfaceState {
    OpmSurface surf{};
    OpmIsometry eff_iso{};   // effective_surface_iso = effective_node_iso o anchor_point_iso (optic_node.rs:372-382)
    OpmAperture aperture{};
    double lidt = 1.0e4;     // 1 J/cm^2 in J/m^2
    std::vector<HitEntry> hits;
    std::vector<std::shared_ptr<Bundle>> fwd_cache, bwd_cache;  // optic_surface.rs:92-143
    std::vector<OpmCriticalFluence> critical;
    double peak_seen = -INFINITY;
};

struct Node {
    OpmNodeDesc d{};
    int parent = -1;
    SurfaceState s[2];          // input_1, output_1
    std::vector<double> tris[2];  // polygon aperture storage
    OpmRays* stored = nullptr;  // bundle kept by a detector (LightData::Geometric, analyzers/raytrace.rs:186-205)
    OpmRays* split = nullptr;   // second output of a beam splitter
    std::vector<std::shared_ptr<Bundle>> stored_gf;  // LightData::GhostFocus kept by a detector
    int main_child = -1, split_child = -1;
    bool placed = true;         // false: waits for opm_scene_calc_node_positions (a node without an isometry of its own)
    double distance = 0.0;      // edge distance from the predecessor (NodeGroup::connect_nodes)
    int align_like = -1;        // NodeAttr::align_like_node_at_distance: take that node's isometry, moved by align_distance along its z
    double align_distance = 0.0;
    bool side_inlined = false;  // the second output's side chain ran inside the main launch (OPM_OP_FORK .. OPM_OP_JOIN)
};

OpmIsometry iso_z(double z) {
    OpmIsometry o;
    double t[3] = {0.0, 0.0, z}, e[3] = {0.0, 0.0, 0.0};
    opm_isometry_new(t, e, &o);
    return o;
}
OpmIsometry iso_append(const OpmIsometry& a, const OpmIsometry& b) {
    OpmIsometry o;
    opm_isometry_append(&a, &b, &o);
    return o;
}

// a per-bundle fluence check of the ghost-focus analysis whose result still sits on the device (opm_scene_gf.inc)
struct GfCheck { SurfaceState* surface; uint64_t uuid; uint32_t lvl; };

}  // namespace

struct OpmScene {
    OpmContext* ctx = nullptr;
    OpmIndexModel ambient{};
    std::vector<std::unique_ptr<Node>> nodes;
    uint64_t next_uuid = 1;
    std::vector<std::vector<std::shared_ptr<Bundle>>> passes;  // ray_collection of every ghost-focus pass
    std::vector<OpmRays*> owned_rays;                           // every device bundle the scene allocated for GF
    OpmRays* main_out = nullptr;                                // output of the main chain with OPM_SCENE_KEEP_SOURCE
    // device buffers are recycled across analyses (cudaMalloc / cudaFree are synchronous and cost milliseconds)
    std::vector<OpmRays*> pool_rays;
    std::vector<OpmHitMap*> pool_hits;
    // ghost focus: device scratch of the per-bundle 101x101 fluence checks (map | one slot of range[4] | peak,total[2] | status
    // per check) and the pinned mirror of the slots: a check costs no synchronisation and no allocation
    void* gf_scratch = nullptr;
    void* gf_host = nullptr;
    void* gf_vor_map = nullptr;  // 101 x 101 map + the bounce-level key of a Voronoi LIDT check
    std::vector<GfCheck> gf_pending;
    void* gf_batch_mem = nullptr;               // maps and accumulators of the deferred checks (opm_scene_gf.inc)
    std::vector<OpmHitMap*> gf_batch_maps;      // deferred checks not launched yet: their hit maps (flat) ...
    std::vector<int32_t> gf_batch_counts;       // ... how many of them each check takes ...
    std::vector<double*> gf_batch_slots;        // ... and the checks' result slots
};

namespace {

OpmStatus acquire_rays(OpmScene* sc, uint64_t cap, OpmRays** out) {
    int best = -1;
    for (size_t k = 0; k < sc->pool_rays.size(); ++k)
        if (opm_rays_capacity(sc->pool_rays[k]) >= cap &&
            (best < 0 || opm_rays_capacity(sc->pool_rays[k]) < opm_rays_capacity(sc->pool_rays[best]))) best = (int)k;
    if (best >= 0) {
        *out = sc->pool_rays[best];
        sc->pool_rays.erase(sc->pool_rays.begin() + best);
        OpmStatus st = opm_rays_disable_position_history(*out);  // a pooled bundle starts without the log of its last use
        return st ? st : opm_rays_clear(*out);
    }
    return opm_rays_create(sc->ctx, cap ? cap : 1, out);
}
OpmStatus acquire_hitmap(OpmScene* sc, uint64_t cap, OpmHitMap** out) {
    int best = -1;
    for (size_t k = 0; k < sc->pool_hits.size(); ++k)
        if (opm_hitmap_capacity(sc->pool_hits[k]) >= cap &&
            (best < 0 || opm_hitmap_capacity(sc->pool_hits[k]) < opm_hitmap_capacity(sc->pool_hits[best]))) best = (int)k;
    if (best >= 0) {
        *out = sc->pool_hits[best];
        sc->pool_hits.erase(sc->pool_hits.begin() + best);
        return OPM_OK;
    }
    return opm_hitmap_create(sc->ctx, cap ? cap : 1, out);
}

void free_node_data(OpmScene* sc, Node& n) {
    for (auto& s : n.s) {
        for (auto& h : s.hits) sc->pool_hits.push_back(h.map);
        s.hits.clear(); s.fwd_cache.clear(); s.bwd_cache.clear(); s.critical.clear(); s.peak_seen = -INFINITY;
    }
    if (n.stored) { sc->pool_rays.push_back(n.stored); n.stored = nullptr; }
    if (n.split) { sc->pool_rays.push_back(n.split); n.split = nullptr; }
    n.stored_gf.clear();
}

// ---- node -> surfaces (OpticNode::update_surfaces)
OpmSurface make_surface(int geom, double param, const OpmIsometry& iso, const OpmCoatingDesc& c) {
    OpmSurface s;
    memset(&s, 0, sizeof s);
    s.geom = geom; s.geom_param = param; s.iso = iso; s.coating = c.type; s.coating_r = c.reflectivity;
    return s;
}

OpmStatus update_surfaces(Node& n) {
    const OpmNodeDesc& d = n.d;
    OpmIsometry node_iso = d.has_alignment ? iso_append(d.iso, d.alignment) : d.iso;  // effective_node_iso (optic_node.rs:356-363)
    OpmIsometry ident;
    opm_isometry_identity(&ident);
    OpmIsometry anchor[2] = {ident, ident};
    const OpmCoatingDesc* coat[2] = {&d.coating_in, &d.coating_out};
    switch (d.type) {
        case OPM_NODE_LENS:
        case OPM_NODE_CYLINDRIC_LENS: {  // nodes/lens/mod.rs:226-293, nodes/cylindric_lens/mod.rs:146-214
            const int curved = d.type == OPM_NODE_LENS ? OPM_GEOM_SPHERE : OPM_GEOM_CYLINDER;
            const double front = d.p[0], rear = d.p[1], thick = d.p[2];
            if (std::isnan(front) || std::isnan(rear) || front == 0.0 || rear == 0.0) return OPM_ERR_ARG;
            if (!(thick >= 0.0) || !std::isfinite(thick)) return OPM_ERR_ARG;  // nodes/lens/mod.rs:110-116
            if (std::isinf(front)) n.s[0].surf = make_surface(OPM_GEOM_PLANE, 0.0, node_iso, *coat[0]);
            else {
                anchor[0] = iso_z(front);
                n.s[0].surf = make_surface(curved, front, iso_append(node_iso, anchor[0]), *coat[0]);
            }
            if (std::isinf(rear)) {
                anchor[1] = iso_z(thick);
                n.s[1].surf = make_surface(OPM_GEOM_PLANE, 0.0, iso_append(node_iso, anchor[1]), *coat[1]);
            } else {
                anchor[1] = iso_z(rear + thick);
                n.s[1].surf = make_surface(curved, rear, iso_append(node_iso, anchor[1]), *coat[1]);
            }
            break;
        }
        case OPM_NODE_WEDGE: {  // nodes/wedge/mod.rs:116-160
            const double thick = d.p[0], angle = d.p[1];
            if (!(thick >= 0.0) || !std::isfinite(thick) || !std::isfinite(angle) || std::fabs(angle) >= M_PI / 2) return OPM_ERR_ARG;
            n.s[0].surf = make_surface(OPM_GEOM_PLANE, 0.0, node_iso, *coat[0]);
            OpmIsometry wedge_iso;
            double t0[3] = {0, 0, 0}, e[3] = {angle, 0.0, 0.0};
            opm_isometry_new(t0, e, &wedge_iso);
            anchor[1] = iso_append(iso_z(thick), wedge_iso);
            n.s[1].surf = make_surface(OPM_GEOM_PLANE, 0.0, iso_append(node_iso, anchor[1]), *coat[1]);
            break;
        }
        case OPM_NODE_THIN_MIRROR: {  // nodes/thin_mirror.rs:121-155: both ports share the geometry
            const double curv = d.p[0];
            if (std::isnan(curv) || curv == 0.0) return OPM_ERR_ARG;
            for (int k = 0; k < 2; ++k) {
                if (std::isinf(curv)) n.s[k].surf = make_surface(OPM_GEOM_PLANE, 0.0, node_iso, *coat[k]);
                else {
                    anchor[k] = iso_z(curv);
                    n.s[k].surf = make_surface(OPM_GEOM_SPHERE, curv, iso_append(node_iso, anchor[k]), *coat[k]);
                }
            }
            break;
        }
        case OPM_NODE_PARABOLIC_MIRROR: {  // nodes/parabolic_mirror.rs:388-421: Parabola(-parent focal length) at the off-axis anchor
            double parent_f = 0.0;
            OpmIsometry off_axis;
            if (opm_parabola_off_axis_isometry(d.p[0], d.p[1], &d.p[2], d.p[4] != 0.0, &off_axis, &parent_f)) return OPM_ERR_ARG;
            for (int k = 0; k < 2; ++k) {
                anchor[k] = off_axis;
                n.s[k].surf = make_surface(OPM_GEOM_PARABOLA, -1.0 * parent_f, iso_append(node_iso, anchor[k]), *coat[k]);
            }
            break;
        }
        default:  // optic_node.rs:62-80 update_flat_single_surfaces
            for (int k = 0; k < 2; ++k) n.s[k].surf = make_surface(OPM_GEOM_PLANE, 0.0, node_iso, *coat[k]);
            break;
    }
    const OpmAperture* ap[2] = {&d.aperture_in, &d.aperture_out};
    for (int k = 0; k < 2; ++k) {
        n.s[k].eff_iso = iso_append(node_iso, anchor[k]);
        n.s[k].aperture = *ap[k];
        if (ap[k]->n_tris > 0 && ap[k]->tris) {  // the polygon triangles are caller memory: keep a copy
            n.tris[k].assign(ap[k]->tris, ap[k]->tris + (size_t)ap[k]->n_tris * 6);
            n.s[k].aperture.tris = n.tris[k].data();
        }
        n.s[k].lidt = d.lidt > 0.0 ? d.lidt : 1.0e4;
    }
    return OPM_OK;
}

// ---- op builders (one op = one `Rays` method call of the reference)
OpmOp op_refract(const SurfaceState& s, const OpmIndexModel* index, bool intended, int strategy, int flags, OpmHitMap* hits,
                 OpmRays* children) {
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_REFRACT;
    op.u.refract.surface = s.surf;
    if (index) { op.u.refract.index = *index; op.u.refract.has_index = 1; }
    op.u.refract.refraction_intended = intended ? 1 : 0;
    op.u.refract.strategy = strategy;
    op.u.refract.flags = flags;
    op.u.refract.hits = hits;
    op.u.refract.children = children;
    return op;
}
OpmOp op_apodize(const OpmAperture& ap, const OpmIsometry& iso) {
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_APODIZE; op.u.apodize.aperture = ap; op.u.apodize.iso = iso;
    return op;
}
OpmOp op_threshold(double e) {
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_THRESHOLD; op.u.threshold.min_energy = e;
    return op;
}
OpmOp op_limits(uint64_t max_bounces, uint64_t max_refr, bool check_refr) {
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_LIMITS; op.u.limits.max_bounces = max_bounces; op.u.limits.max_refractions = max_refr;
    op.u.limits.check_refractions = check_refr ? 1 : 0;
    return op;
}

// ReflectiveGrating RT analyze (nodes/reflective_grating.rs:209-262): grating vector 2 pi * line density * (effective
// surface iso applied to x), vacuum behind the surface, refraction_intended = false, the diffracted bundle continues
OpmOp op_diffract(const Node& n, const SurfaceState& s, int extra_flags, OpmRays* children) {
    OpmOp op;
    memset(&op, 0, sizeof op);
    op.kind = OPM_OP_DIFFRACT;
    OpmDiffractOp& g = op.u.diffract;
    g.surface = s.surf;
    g.index.type = OPM_IDX_CONST; g.index.c[0] = 1.0;  // refr_index_vaccuum()
    const double dens = n.d.p[0];
    for (int k = 0; k < 3; ++k) g.grating_vector[k] = 2. * M_PI * dens * s.eff_iso.fwd.rot[3 * k + 0];  // iso.transform_vector(x)
    g.order = (int32_t)n.d.p[1];
    g.refraction_intended = 0;
    g.flags = OPM_REFRACT_CONTINUE_AS_CHILD | extra_flags;
    g.children = children;
    return op;
}

// Runs an op list, splitting it into several launches when it exceeds OPM_MAX_PROGRAM_OPS or when every reference
// call should be its own launch (OPM_SCENE_PER_SURFACE: the un-fused baseline).
struct Runner {
    OpmRays* rays;
    int flags;
    OpmRays* first_out = nullptr;  // OPM_SCENE_KEEP_SOURCE: the first launch reads `rays` and writes here, later ones run on it in place
    OpmTraceStats total{};
    std::vector<OpmOp> ops;
    OpmStatus status = OPM_OK;

    void push(const OpmOp& op) {
        if (status) return;
        ops.push_back(op);
        if ((flags & OPM_SCENE_PER_SURFACE) || ops.size() == OPM_MAX_PROGRAM_OPS) flush();
    }
    void flush() {
        if (status || ops.empty()) return;
        OpmTraceStats st;
        memset(&st, 0, sizeof st);
        const bool async = (flags & OPM_SCENE_ASYNC) != 0;
        status = opm_trace_sequence(rays, ops.data(), (int32_t)ops.size(), flags & OPM_TRACE_STRICT, first_out, async ? nullptr : &st);
        if (first_out) { rays = first_out; first_out = nullptr; }
        ops.clear();
        total.interactions += st.interactions; total.children += st.children; total.hits += st.hits; total.missed += st.missed;
        total.apodized += st.apodized; total.valid_out = st.valid_out; total.alive_out = st.alive_out; total.status_bits |= st.status_bits;
        total.first_valid_key = st.first_valid_key;
    }
};

OpmHitMap* new_hitmap(OpmScene* sc, SurfaceState& s, uint64_t uuid, uint64_t n, OpmStatus* st) {
    OpmHitMap* h = nullptr;
    *st = acquire_hitmap(sc, n ? n : 1, &h);
    if (*st) return nullptr;
    s.hits.push_back({uuid, h});
    return h;
}

// `AnalysisRayTrace::analyze` of one node, forward direction, appended to the runner's program
OpmStatus compile_node_rt(OpmScene* sc, Node& n, const OpmRayTraceConfig& cfg, int flags, uint64_t n_rays, Runner& R, int depth = 0);

// the side chain hanging off a splitter's second output, inlined between OPM_OP_FORK and OPM_OP_JOIN: true when the side
// chain (no further splitter with a second output in it) and the rest fit comfortably into the current launch
bool can_inline_side_chain(OpmScene* sc, const Node& sp, int flags, const Runner& R, int depth) {
    if ((flags & (OPM_SCENE_PER_SURFACE | OPM_SCENE_POSITION_HISTORY)) || depth != 0 || sp.split_child < 0) return false;
    size_t nodes = 0;
    for (int i = sp.split_child; i >= 0; i = sc->nodes[i]->main_child) {
        if (sc->nodes[i]->d.type == OPM_NODE_BEAM_SPLITTER && sc->nodes[i]->split_child >= 0) return false;
        ++nodes;
    }
    return R.ops.size() + 8 * (nodes + 1) + 8 <= OPM_MAX_PROGRAM_OPS;
}

OpmStatus compile_node_rt(OpmScene* sc, Node& n, const OpmRayTraceConfig& cfg, int flags, uint64_t n_rays, Runner& R, int depth) {
    const int strat = cfg.missed_surface_strategy;
    const double thr = cfg.min_energy_per_ray;
    OpmStatus st = OPM_OK;
    auto hits_for = [&](SurfaceState& s) -> OpmHitMap* {
        // by default only where a report reads the hit map: FluenceDetector (the other detectors report from their copy of the bundle)
        if (!(flags & OPM_SCENE_RECORD_ALL_HITS) && n.d.type != OPM_NODE_FLUENCE_DETECTOR) return nullptr;
        return new_hitmap(sc, s, 1, n_rays, &st);
    };
    switch (n.d.type) {
        case OPM_NODE_SOURCE: {  // nodes/source.rs:189-235
            OpmOp op;
            memset(&op, 0, sizeof op);
            op.kind = OPM_OP_TRANSFORM; op.u.transform.iso = n.s[0].eff_iso; op.u.transform.inverse = 0;
            R.push(op);
            R.push(op_apodize(n.s[1].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_LENS:
        case OPM_NODE_CYLINDRIC_LENS:
        case OPM_NODE_WEDGE: {  // nodes/lens/analysis_raytrace.rs:11-55 -> pass_through_surface x2 (analyzers/raytrace.rs:96-140)
            OpmHitMap* h0 = hits_for(n.s[0]);
            if (st) return st;
            R.push(op_refract(n.s[0], &n.d.index, true, strat, 0, h0, nullptr));
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            OpmHitMap* h1 = hits_for(n.s[1]);
            if (st) return st;
            R.push(op_refract(n.s[1], &sc->ambient, true, strat, 0, h1, nullptr));
            R.push(op_apodize(n.s[1].aperture, n.s[1].eff_iso));
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_THIN_MIRROR:
        case OPM_NODE_PARABOLIC_MIRROR: {  // nodes/thin_mirror.rs:209-250, nodes/parabolic_mirror.rs:474-517
            OpmHitMap* h0 = hits_for(n.s[0]);
            if (st) return st;
            R.push(op_refract(n.s[0], nullptr, false, strat, OPM_REFRACT_CONTINUE_AS_CHILD, h0, nullptr));
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_REFLECTIVE_GRATING: {  // nodes/reflective_grating.rs:209-262
            R.push(op_diffract(n, n.s[0], 0, nullptr));
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_BEAM_SPLITTER: {  // nodes/beam_splitter/mod.rs:158-293 (one input)
            OpmHitMap* h0 = hits_for(n.s[0]);
            if (st) return st;
            R.push(op_refract(n.s[0], nullptr, true, strat, 0, h0, nullptr));
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            if (n.split_child >= 0 && !n.split) {
                st = acquire_rays(sc, n_rays ? n_rays : 1, &n.split);
                if (st) return st;
            }
            OpmOp op;
            memset(&op, 0, sizeof op);
            n.side_inlined = can_inline_side_chain(sc, n, flags, R, depth);
            if (n.side_inlined) {
                // the second output's side chain runs in this launch on the split-off rays (same ops the separate launch
                // would run: out2 aperture in the frame of out1, threshold, limits, then the side chain's nodes)
                op.kind = OPM_OP_FORK; op.u.fork.ratio = n.d.p[0]; op.u.fork.spectrum = n.d.spectrum;
                R.push(op);
                R.push(op_apodize(n.s[1].aperture, n.s[1].eff_iso));
                R.push(op_threshold(thr));
                R.push(op_limits(cfg.max_number_of_bounces, cfg.max_number_of_refractions, true));
                for (int i = n.split_child; i >= 0; i = sc->nodes[i]->main_child) {
                    st = compile_node_rt(sc, *sc->nodes[i], cfg, flags, n_rays, R, depth + 1);
                    if (st) return st;
                }
                memset(&op, 0, sizeof op);
                op.kind = OPM_OP_JOIN; op.u.join.out = n.split;
                R.push(op);
            } else {
                op.kind = OPM_OP_SPLIT; op.u.split.ratio = n.d.p[0]; op.u.split.split_out = n.split; op.u.split.spectrum = n.d.spectrum;
                R.push(op);
            }
            R.push(op_apodize(n.s[1].aperture, n.s[1].eff_iso));  // out1 aperture, iso of out1_port
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_DUMMY:
        case OPM_NODE_IDEAL_FILTER:
        case OPM_NODE_PARAXIAL_SURFACE: {  // nodes/dummy.rs:83-135, ideal_filter.rs:203-250, paraxial_surface.rs:169-230
            OpmHitMap* h0 = hits_for(n.s[0]);
            if (st) return st;
            R.push(op_refract(n.s[0], nullptr, true, strat, 0, h0, nullptr));
            if (n.d.type == OPM_NODE_IDEAL_FILTER) {
                OpmOp op;
                memset(&op, 0, sizeof op);
                op.kind = OPM_OP_FILTER; op.u.filter.transmission = n.d.p[0]; op.u.filter.spectrum = n.d.spectrum;
                R.push(op);
            } else if (n.d.type == OPM_NODE_PARAXIAL_SURFACE) {
                OpmOp op;
                memset(&op, 0, sizeof op);
                op.kind = OPM_OP_PARAXIAL; op.u.paraxial.focal_length = n.d.p[0]; op.u.paraxial.iso = n.s[0].eff_iso;
                R.push(op);
            }
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            R.push(op_apodize(n.s[1].aperture, n.s[0].eff_iso));  // output aperture in the INPUT surface frame
            R.push(op_threshold(thr));
            break;
        }
        case OPM_NODE_SPOT_DIAGRAM:
        case OPM_NODE_RAY_PROPAGATION_VISUALIZER:
        case OPM_NODE_SPECTROMETER:
        case OPM_NODE_WAVEFRONT:
        case OPM_NODE_ENERGY_METER:
        case OPM_NODE_FLUENCE_DETECTOR: {  // pass_through_detector_surface analyzers/raytrace.rs:150-207
            OpmHitMap* h0 = hits_for(n.s[0]);
            if (st) return st;
            R.push(op_refract(n.s[0], nullptr, true, strat, 0, h0, nullptr));
            R.push(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
            R.push(op_threshold(thr));
            // the detector's `light_data` (analyzers/raytrace.rs:186-205).  FluenceDetector::node_report reads only the hit map
            // (nodes/fluence_detector/mod.rs:91-135): its copy of the bundle is written only on request
            const bool keep = n.d.type != OPM_NODE_FLUENCE_DETECTOR || (flags & OPM_SCENE_KEEP_LIGHT_DATA);
            if (keep && !(flags & OPM_SCENE_NO_SNAPSHOTS)) {
                if (!n.stored) {
                    st = acquire_rays(sc, n_rays ? n_rays : 1, &n.stored);
                    if (st) return st;
                }
                OpmOp op;
                memset(&op, 0, sizeof op);
                op.kind = OPM_OP_SNAPSHOT; op.u.snapshot.out = n.stored;
                if (n.d.type == OPM_NODE_SPOT_DIAGRAM) {  // the report is opm_scene_node_spot: statistics in the detector's frame
                    op.u.snapshot.want_spot_sums = 1;
                    op.u.snapshot.spot_frame = &n.s[0].eff_iso;
                    if ((flags & OPM_SCENE_LEAN_SPOT_DATA) && !(flags & OPM_SCENE_POSITION_HISTORY)) op.u.snapshot.fields = OPM_SNAPSHOT_SPOT;
                }
                R.push(op);
            }
            break;
        }
        default: return OPM_ERR_ARG;
    }
    // filter_ray_limits after every node (nodes/node_group/analysis_raytrace.rs:20-27,66)
    R.push(op_limits(cfg.max_number_of_bounces, cfg.max_number_of_refractions, true));
    return R.status;
}

OpmStatus link_nodes(OpmScene* sc) {
    const int N = (int)sc->nodes.size();
    for (int i = 0; i < N; ++i) { sc->nodes[i]->main_child = -1; sc->nodes[i]->split_child = -1; }
    for (int i = 0; i < N; ++i) {
        Node& n = *sc->nodes[i];
        n.parent = n.d.parent < 0 ? i - 1 : n.d.parent;
        if (i == 0) { n.parent = -1; continue; }
        if (n.parent < 0 || n.parent >= i) return OPM_ERR_ARG;
        Node& p = *sc->nodes[n.parent];
        if (n.d.parent_port == 1) {
            if (p.d.type != OPM_NODE_BEAM_SPLITTER || p.split_child >= 0) return OPM_ERR_ARG;
            p.split_child = i;
        } else {
            if (p.main_child >= 0) return OPM_ERR_ARG;
            p.main_child = i;
        }
    }
    return OPM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" void opm_node_desc_init(OpmNodeDesc* d, int32_t type) {
    memset(d, 0, sizeof *d);
    d->type = type;
    d->parent = -1;
    opm_isometry_identity(&d->iso);
    opm_isometry_identity(&d->alignment);
    d->index.type = OPM_IDX_CONST; d->index.c[0] = 1.5;  // nodes/lens/mod.rs:81, nodes/wedge/mod.rs:63
    d->coating_in.type = d->coating_out.type = OPM_COAT_IDEAL_AR;  // surface/optic_surface.rs:54
    if (type == OPM_NODE_THIN_MIRROR || type == OPM_NODE_PARABOLIC_MIRROR) {  // nodes/thin_mirror.rs:67-82
        d->coating_in.type = d->coating_out.type = OPM_COAT_CONSTANT_R;
        d->coating_in.reflectivity = d->coating_out.reflectivity = 1.0;
    }
    switch (type) {
        case OPM_NODE_LENS: case OPM_NODE_CYLINDRIC_LENS: d->p[0] = 500.0 * 1e-3; d->p[1] = -500.0 * 1e-3; d->p[2] = 10.0 * 1e-3; break;
        case OPM_NODE_WEDGE: d->p[0] = 10.0 * 1e-3; d->p[1] = 0.0; break;
        case OPM_NODE_THIN_MIRROR: d->p[0] = INFINITY; break;
        case OPM_NODE_PARABOLIC_MIRROR: d->p[0] = 1.0; break;
        case OPM_NODE_BEAM_SPLITTER: d->p[0] = 0.5; break;
        case OPM_NODE_IDEAL_FILTER: d->p[0] = 1.0; break;
        case OPM_NODE_PARAXIAL_SURFACE: d->p[0] = 10.0 * 1e-3; break;
        case OPM_NODE_REFLECTIVE_GRATING: d->p[0] = 1740.0 * 1.0e3; d->p[1] = -1.0; break;  // nodes/reflective_grating.rs:60-75
        default: break;
    }
    d->lidt = 1.0e4;
}

extern "C" OpmStatus opm_scene_create(OpmContext* ctx, OpmScene** out) {
    if (!ctx || !out) return OPM_ERR_ARG;
    OpmScene* s = new OpmScene();
    s->ctx = ctx;
    memset(&s->ambient, 0, sizeof s->ambient);
    s->ambient.type = OPM_IDX_CONST; s->ambient.c[0] = 1.0;  // vacuum (optic_scenery_rsc.rs:14-19)
    *out = s;
    return OPM_OK;
}
extern "C" OpmStatus opm_scene_reset_data(OpmScene* sc) {
    if (!sc) return OPM_ERR_ARG;
    for (auto& n : sc->nodes) free_node_data(sc, *n);
    sc->passes.clear();
    for (OpmRays* r : sc->owned_rays) sc->pool_rays.push_back(r);
    sc->owned_rays.clear();
    if (sc->main_out) { sc->pool_rays.push_back(sc->main_out); sc->main_out = nullptr; }
    return OPM_OK;
}
extern "C" void opm_scene_destroy(OpmScene* sc) {
    if (!sc) return;
    opm_scene_reset_data(sc);
    for (OpmRays* r : sc->pool_rays) opm_rays_destroy(r);
    for (OpmHitMap* h : sc->pool_hits) opm_hitmap_destroy(h);
    if (sc->gf_scratch) opm_device_free(sc->ctx, sc->gf_scratch);
    if (sc->gf_vor_map) opm_device_free(sc->ctx, sc->gf_vor_map);
    if (sc->gf_batch_mem) opm_device_free(sc->ctx, sc->gf_batch_mem);
    if (sc->gf_host) opm_host_free_pinned(sc->gf_host);
    delete sc;
}
extern "C" OpmStatus opm_scene_set_ambient(OpmScene* sc, const OpmIndexModel* a) {
    if (!sc || !a) return OPM_ERR_ARG;
    sc->ambient = *a;
    return OPM_OK;
}
extern "C" OpmStatus opm_scene_add_node(OpmScene* sc, const OpmNodeDesc* d, int32_t* idx) {
    if (!sc || !d) return OPM_ERR_ARG;
    if (sc->nodes.empty() && d->type != OPM_NODE_SOURCE) return OPM_ERR_ARG;  // the chain starts at a Source
    auto n = std::make_unique<Node>();
    n->d = *d;
    OpmStatus st = update_surfaces(*n);
    if (st) return st;
    sc->nodes.push_back(std::move(n));
    if (idx) *idx = (int32_t)sc->nodes.size() - 1;
    return link_nodes(sc);
}
extern "C" int32_t opm_scene_node_count(const OpmScene* sc) { return sc ? (int32_t)sc->nodes.size() : 0; }

// ---------------------------------------------------------------------------------------------
// node positioning from the optical axis
// ---------------------------------------------------------------------------------------------
extern "C" OpmStatus opm_scene_add_node_at_distance(OpmScene* sc, const OpmNodeDesc* d, double distance, int32_t* idx) {
    if (!sc || !d || !std::isfinite(distance)) return OPM_ERR_ARG;  // "propagation length must be finite" (ray.rs:414-418)
    if (sc->nodes.empty()) return OPM_ERR_ARG;                      // the source carries an isometry of its own
    OpmNodeDesc copy = *d;
    opm_isometry_identity(&copy.iso);  // placeholder: validates the parameters, replaced when the node is placed
    int32_t i = -1;
    OpmStatus st = opm_scene_add_node(sc, &copy, &i);
    if (st) return st;
    sc->nodes[i]->placed = false;
    sc->nodes[i]->distance = distance;
    if (idx) *idx = i;
    return OPM_OK;
}
extern "C" OpmStatus opm_scene_add_node_aligned_like(OpmScene* sc, const OpmNodeDesc* d, int32_t like_node, double distance,
                                                     double axis_distance, int32_t* idx) {
    if (!sc || like_node < 0 || like_node >= (int)sc->nodes.size() || !std::isfinite(distance)) return OPM_ERR_ARG;
    int32_t i = -1;
    OpmStatus st = opm_scene_add_node_at_distance(sc, d, axis_distance, &i);
    if (st) return st;
    sc->nodes[i]->align_like = like_node;
    sc->nodes[i]->align_distance = distance;
    if (idx) *idx = i;
    return OPM_OK;
}
extern "C" OpmStatus opm_scene_node_isometry(const OpmScene* sc, int32_t node, OpmIsometry* out) {
    if (!sc || !out || node < 0 || node >= (int)sc->nodes.size() || !sc->nodes[node]->placed) return OPM_ERR_ARG;
    *out = sc->nodes[node]->d.iso;
    return OPM_OK;
}

namespace {
struct AxisRay { OpmRay ray; bool has_prev = false; double prev_dir[3] = {0, 0, 0}; bool set = false; };

// Ray::define_up_direction (ray.rs:818-838)
void define_up_direction(const double dir_in[3], double up[3]) {
    auto cross_norm = [](const double a[3], const double b[3]) {
        const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        return std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    };
    const double len = std::sqrt(dir_in[0] * dir_in[0] + dir_in[1] * dir_in[1] + dir_in[2] * dir_in[2]);
    const double dir[3] = {dir_in[0] / len, dir_in[1] / len, dir_in[2] / len};
    const double X[3] = {1, 0, 0}, Y[3] = {0, 1, 0}, Z[3] = {0, 0, 1};
    const double eps = 2.220446049250313e-16;
    if (cross_norm(dir, X) < eps || cross_norm(dir, Z) < eps) { up[0] = 0; up[1] = 1; up[2] = 0; return; }
    if (cross_norm(dir, Y) < eps) { up[0] = 1; up[1] = 0; up[2] = 0; return; }
    const double k = dir[1];  // y . dir
    double pr[3] = {-k * dir[0], 1.0 - k * dir[1], -k * dir[2]};
    const double pl = std::sqrt(pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2]);
    for (int c = 0; c < 3; ++c) up[c] = pr[c] / pl;
}
// Ray::calc_new_up_direction (ray.rs:842-860): up <- Rotation3::rotation_between(prev_dir, dir) * up when the direction changed
// by more than 1000 eps (axis = unit(a x b), angle = acos(a . b), Rodrigues matrix; parallel: identity, anti-parallel: untouched)
void calc_new_up_direction(const double prev[3], const double dir[3], double up[3]) {
    const double eps = 2.220446049250313e-16;
    auto norm = [](const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    const double d[3] = {prev[0] - dir[0], prev[1] - dir[1], prev[2] - dir[2]};
    if (!(norm(d) > eps * 1000.)) return;
    const double la = norm(prev), lb = norm(dir);
    if (!(la > 0.0) || !(lb > 0.0)) return;
    const double a[3] = {prev[0] / la, prev[1] / la, prev[2] / la}, b[3] = {dir[0] / lb, dir[1] / lb, dir[2] / lb};
    const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const double cn = norm(c);
    if (!(cn > eps)) return;
    const double ux = c[0] / cn, uy = c[1] / cn, uz = c[2] / cn;
    const double angle = std::acos(a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
    if (angle == 0.0) return;
    const double sn = std::sin(angle), cs = std::cos(angle), omc = 1.0 - cs;
    const double m[9] = {ux * ux + (1.0 - ux * ux) * cs, ux * uy * omc - uz * sn,       ux * uz * omc + uy * sn,
                         ux * uy * omc + uz * sn,       uy * uy + (1.0 - uy * uy) * cs, uy * uz * omc - ux * sn,
                         ux * uz * omc - uy * sn,       uy * uz * omc + ux * sn,       uz * uz + (1.0 - uz * uz) * cs};
    const double v[3] = {up[0], up[1], up[2]};
    for (int r = 0; r < 3; ++r) up[r] = m[3 * r] * v[0] + m[3 * r + 1] * v[1] + m[3 * r + 2] * v[2];
}

// passes the one-ray bundle through `ops`, one op per launch; the ray is read back after every op that can change its
// direction, so that `prev_dir` (the direction before the last surface that was reached: ray.rs:450,541,628-631,676) is known
OpmStatus run_axis_ops(OpmRays* bundle, const std::vector<OpmOp>& ops, int flags, AxisRay& a) {
    for (const OpmOp& op : ops) {
        const bool turns = op.kind == OPM_OP_REFRACT || op.kind == OPM_OP_DIFFRACT || op.kind == OPM_OP_PARAXIAL;
        OpmTraceStats st;
        memset(&st, 0, sizeof st);
        OpmStatus rc = opm_trace_sequence(bundle, &op, 1, flags & OPM_TRACE_STRICT, nullptr, &st);
        if (rc) return rc;
        if (!turns) continue;
        OpmRay after;
        uint64_t n = 0;
        rc = opm_rays_download(bundle, &after, 1, &n);
        if (rc) return rc;
        if (n != 1) return OPM_ERR_ARG;  // "no rays in this ray bundle. cannot position nodes" (optic_graph.rs:919-923)
        const bool was_valid = a.ray.valid != 0;
        const bool reached = op.kind == OPM_OP_PARAXIAL ? was_valid
                                                        : (was_valid && st.interactions == 1 && (st.missed == 0 || after.bounces != a.ray.bounces));
        if (reached) { a.has_prev = true; memcpy(a.prev_dir, a.ray.dir, sizeof a.prev_dir); }
        a.ray = after;
    }
    uint64_t n = 0;
    OpmStatus rc = opm_rays_download(bundle, &a.ray, 1, &n);  // energy / validity after the trailing apodize, threshold, limits
    if (rc) return rc;
    return n == 1 ? OPM_OK : OPM_ERR_ARG;
}
}  // namespace

extern "C" OpmStatus opm_scene_calc_node_positions(OpmScene* sc, double wavelength, const OpmRayTraceConfig* cfg_in) {
    if (!sc || sc->nodes.empty() || !(wavelength > 0.0) || !std::isfinite(wavelength)) return OPM_ERR_ARG;
    OpmRayTraceConfig cfg;
    if (cfg_in) cfg = *cfg_in; else opm_raytrace_config_default(&cfg);
    OpmStatus st = opm_scene_reset_data(sc);
    if (st) return st;
    const int N = (int)sc->nodes.size();
    std::vector<AxisRay> out_main(N), out_split(N);
    double up[3] = {0.0, 1.0, 0.0};  // analysis_raytrace.rs:103
    OpmRays* bundle = nullptr;
    st = acquire_rays(sc, 1, &bundle);
    if (st) return st;
    auto done = [&](OpmStatus rc) { sc->pool_rays.push_back(bundle); opm_scene_reset_data(sc); return rc; };
    const int flags = 0;
    for (int i = 0; i < N; ++i) {
        Node& n = *sc->nodes[i];
        AxisRay a;
        if (i == 0) {  // nodes/source.rs:236-268: Ray::new_collimated(origin, wavelength, 1 J) moved into the source's frame
            memset(&a.ray, 0, sizeof a.ray);
            a.ray.dir[2] = 1.0; a.ray.e = 1.0; a.ray.wvl = wavelength; a.ray.n = 1.0; a.ray.valid = 1;
            st = opm_rays_upload(bundle, &a.ray, 1);
            if (st) return done(st);
            OpmOp op;
            memset(&op, 0, sizeof op);
            op.kind = OPM_OP_TRANSFORM; op.u.transform.iso = n.s[0].eff_iso; op.u.transform.inverse = 0;
            st = run_axis_ops(bundle, {op}, flags, a);
            if (st) return done(st);
            define_up_direction(a.ray.dir, up);
            a.set = true;
            out_main[0] = a;
            continue;
        }
        const AxisRay& in = (n.d.parent_port == 1 ? out_split : out_main)[n.parent];
        if (!in.set) return done(OPM_ERR_ARG);
        a = in;  // prev_dir is a field of the ray: it survives a node whose surface the ray does not reach
        if (!n.placed && n.align_like >= 0 && sc->nodes[n.align_like]->placed) {
            // analysis_raytrace.rs:142-155: the reference node's isometry (without its local alignment) o T(0, 0, distance)
            n.d.iso = iso_append(sc->nodes[n.align_like]->d.iso, iso_z(n.align_distance));
            st = update_surfaces(n);
            if (st) return done(st);
            n.placed = true;
        }
        if (!n.placed) {  // optic_graph.rs:878-926: ray.propagate(distance), Isometry::new_from_view(position, direction, up)
            double pos[3];
            for (int c = 0; c < 3; ++c) pos[c] = a.ray.pos[c] + n.distance * a.ray.dir[c];
            st = opm_isometry_from_view(pos, a.ray.dir, up, &n.d.iso);
            if (st) return done(st);
            st = update_surfaces(n);
            if (st) return done(st);
            n.placed = true;
        }
        st = opm_rays_upload(bundle, &a.ray, 1);
        if (st) return done(st);
        Runner R{bundle, flags};
        int edges = 1;
        if (n.d.type == OPM_NODE_BEAM_SPLITTER) {  // beam_splitter/analysis_raytrace.rs:40-93: the input surface only
            R.ops.push_back(op_refract(n.s[0], nullptr, true, cfg.missed_surface_strategy, 0, nullptr, nullptr));
            edges = 2;
        } else {
            st = compile_node_rt(sc, n, cfg, OPM_SCENE_NO_SNAPSHOTS, 1, R);  // ops are collected, not launched (a node has < 128 ops)
            if (st) return done(st);
        }
        std::vector<OpmOp> ops;
        ops.swap(R.ops);
        for (OpmOp& op : ops)  // the axis ray records nothing (the reference resets the data afterwards)
            if (op.kind == OPM_OP_REFRACT) op.u.refract.hits = nullptr;
        st = run_axis_ops(bundle, ops, flags, a);
        if (st) return done(st);
        a.set = true;
        out_main[i] = a;
        if (edges == 2) out_split[i] = a;
        for (int e = 0; e < edges; ++e) {  // once per outgoing edge (analysis_raytrace.rs:199-208)
            if (!a.has_prev) return done(OPM_ERR_ARG);  // "no previous direction of ray defined to calculate new up-direction!"
            calc_new_up_direction(a.prev_dir, a.ray.dir, up);
        }
    }
    return done(OPM_OK);
}

// RayTracingAnalyzer::analyze (analyzers/raytrace.rs:36-49): reset_data, then the topological walk
extern "C" OpmStatus opm_scene_raytrace(OpmScene* sc, OpmRays* rays, const OpmRayTraceConfig* cfg_in, int32_t flags,
                                        OpmTraceStats* stats) {
    if (!sc || !rays || sc->nodes.empty()) return OPM_ERR_ARG;
    OpmRayTraceConfig cfg;
    if (cfg_in) cfg = *cfg_in; else opm_raytrace_config_default(&cfg);
    for (auto& n : sc->nodes) if (!n->placed) return OPM_ERR_ARG;  // "no effective isometry defined" (analyzers/raytrace.rs:108-112)
    OpmStatus st = opm_scene_reset_data(sc);
    if (st) return st;
    const uint64_t n_rays = opm_rays_len(rays);
    OpmTraceStats total;
    memset(&total, 0, sizeof total);
    // main chain, then the side chains hanging off beam splitters
    struct Job { int start; OpmRays* rays; int splitter; };
    std::vector<Job> jobs;
    jobs.push_back({0, rays, -1});
    if (flags & OPM_SCENE_KEEP_SOURCE) {
        st = acquire_rays(sc, n_rays ? n_rays : 1, &sc->main_out);
        if (st) return st;
    }
    if (flags & OPM_SCENE_POSITION_HISTORY) {  // the working bundle logs; stored and split-off copies inherit (opm_trace_sequence)
        st = opm_rays_enable_position_history(sc->main_out ? sc->main_out : rays, 2 * (uint32_t)sc->nodes.size() + 4);
        if (st) return st;
    }
    for (size_t j = 0; j < jobs.size(); ++j) {
        Runner R{jobs[j].rays, flags};
        if (j == 0) R.first_out = sc->main_out;
        if (jobs[j].splitter >= 0) {  // second output: apodize(out2 aperture, iso of out1) + threshold + limits
            Node& sp = *sc->nodes[jobs[j].splitter];
            R.push(op_apodize(sp.s[1].aperture, sp.s[1].eff_iso));
            R.push(op_threshold(cfg.min_energy_per_ray));
            R.push(op_limits(cfg.max_number_of_bounces, cfg.max_number_of_refractions, true));
        }
        for (int i = jobs[j].start; i >= 0; i = sc->nodes[i]->main_child) {
            Node& n = *sc->nodes[i];
            st = compile_node_rt(sc, n, cfg, flags, n_rays, R);
            if (st) return st;
            if (n.d.type == OPM_NODE_BEAM_SPLITTER && n.split_child >= 0 && !n.side_inlined) jobs.push_back({n.split_child, n.split, i});
        }
        R.flush();
        if (R.status) return R.status;
        total.interactions += R.total.interactions; total.children += R.total.children; total.hits += R.total.hits;
        total.missed += R.total.missed; total.apodized += R.total.apodized; total.status_bits |= R.total.status_bits;
        if (j == 0) { total.valid_out = R.total.valid_out; total.alive_out = R.total.alive_out; total.first_valid_key = R.total.first_valid_key; }
    }
    if (stats) *stats = total;
    return OPM_OK;
}

extern "C" OpmRays* opm_scene_output_rays(OpmScene* sc) { return sc ? sc->main_out : nullptr; }
extern "C" OpmRays* opm_scene_node_rays(OpmScene* sc, int32_t node) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size()) return nullptr;
    Node& n = *sc->nodes[node];
    return n.d.type == OPM_NODE_BEAM_SPLITTER ? n.split : n.stored;
}
extern "C" OpmStatus opm_scene_node_spot(OpmScene* sc, int32_t node, OpmSpotStats* out) {
    if (!sc || !out || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    return opm_rays_spot_stats(n.stored, &n.s[0].eff_iso, out);  // nodes/spot_diagram.rs:106-112
}
extern "C" OpmStatus opm_scene_node_spot_partial(OpmScene* sc, int32_t node, int32_t phase, OpmSpotPartial* inout) {
    if (!sc || !inout || node < 0 || node >= (int)sc->nodes.size() || (phase != 1 && phase != 2)) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    return phase == 1 ? opm_rays_spot_partial_sums(n.stored, &n.s[0].eff_iso, inout)
                      : opm_rays_spot_partial_radii(n.stored, &n.s[0].eff_iso, inout);
}
extern "C" OpmStatus opm_scene_node_spot_async(OpmScene* sc, int32_t node, int32_t phase, double* sums_dev, double* radii_dev) {
    if (!sc || !sums_dev || node < 0 || node >= (int)sc->nodes.size() || (phase != 1 && phase != 2)) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    return phase == 1 ? opm_rays_spot_sums_async(n.stored, &n.s[0].eff_iso, sums_dev)
                      : opm_rays_spot_radii_async(n.stored, &n.s[0].eff_iso, sums_dev, radii_dev);
}
// EnergyMeter::node_report (nodes/energy_meter.rs:136-151)
extern "C" OpmStatus opm_scene_node_energy(OpmScene* sc, int32_t node, double* energy) {
    if (!sc || !energy || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    OpmSpotStats ss;
    OpmStatus st = opm_rays_spot_stats(n.stored, nullptr, &ss);
    if (st) return st;
    *energy = ss.total_energy;
    return OPM_OK;
}
// Spectrometer::node_report (nodes/spectrometer.rs:155-175): to_spectrum(0.2 nm) of the kept bundle
extern "C" OpmStatus opm_scene_node_spectrum(OpmScene* sc, int32_t node, double resolution_m, const double* range_m, double* lambda_um,
                                             double* values, uint64_t cap, uint64_t* n_out) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    return opm_rays_to_spectrum(n.stored, resolution_m > 0.0 ? resolution_m : 0.2e-9, range_m, lambda_um, values, cap, n_out);
}
// WaveFront::node_report (nodes/wavefront.rs:173-235): monitor frame = effective_surface_iso("input_1")
extern "C" OpmStatus opm_scene_node_wavefront(OpmScene* sc, int32_t node, double* xyw_out, uint64_t cap, OpmWavefrontStats* out) {
    if (!sc || !out || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (!n.stored) return OPM_ERR_ARG;
    return opm_rays_wavefront_error(n.stored, &n.s[0].eff_iso, 0.0, xyw_out, cap, out);
}
extern "C" int32_t opm_scene_node_hitmaps(OpmScene* sc, int32_t node, int32_t surface, OpmHitMap** out, int32_t cap) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size() || surface < 0 || surface > 1) return 0;
    auto& hits = sc->nodes[node]->s[surface].hits;
    int32_t k = 0;
    for (auto& h : hits) { if (out && k < cap) out[k] = h.map; ++k; }
    return k;
}
extern "C" OpmStatus opm_scene_node_fluence(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                            double* peak, double* total) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    std::vector<OpmHitMap*> maps;
    for (auto& h : n.s[0].hits) maps.push_back(h.map);  // hit_maps()["input_1"], all bounces and bundles merged
    if (maps.empty()) return OPM_ERR_EMPTY_HITMAP;
    void* dev = nullptr;
    OpmStatus st = opm_device_alloc(sc->ctx, nx * ny * sizeof(double), &dev);
    if (st) return st;
    double box[4];
    st = opm_fluence_binning(maps.data(), (int32_t)maps.size(), -1, nx, ny, nullptr, nullptr, 0, (double*)dev, box);
    if (!st) st = opm_fluence_peak_total(sc->ctx, (const double*)dev, nx, ny, box, peak, total);
    if (!st && map_host) st = opm_device_to_host(sc->ctx, map_host, dev, nx * ny * sizeof(double));
    if (lrtb) memcpy(lrtb, box, sizeof box);
    opm_device_free(sc->ctx, dev);
    return st;
}

// FluenceDetector::node_report with FluenceEstimator::KDE (nodes/fluence_detector/mod.rs:91-135, rays_hit_map.rs:575-613)
extern "C" OpmStatus opm_scene_node_fluence_kde(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                                double* peak, double* total, double* band_width) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size()) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    std::vector<OpmHitMap*> maps;
    for (auto& h : n.s[0].hits) maps.push_back(h.map);
    if (maps.empty()) return OPM_ERR_EMPTY_HITMAP;
    void* dev = nullptr;
    OpmStatus st = opm_device_alloc(sc->ctx, nx * ny * sizeof(double), &dev);
    if (st) return st;
    double box[4];
    st = opm_fluence_kde(maps.data(), (int32_t)maps.size(), -1, nx, ny, nullptr, (double*)dev, box, band_width);
    if (!st) st = opm_fluence_peak_total(sc->ctx, (const double*)dev, nx, ny, box, peak, total);
    if (!st && map_host) st = opm_device_to_host(sc->ctx, map_host, dev, nx * ny * sizeof(double));
    if (lrtb) memcpy(lrtb, box, sizeof box);
    opm_device_free(sc->ctx, dev);
    return st;
}

// BeamSplitter::analyze_raytrace with BOTH inputs (nodes/beam_splitter/mod.rs:158-293; analysis_raytrace.rs:12-39): each input
// passes the node's flat input surface and input aperture and is split (Rays::split, rays.rs:1253-1264); output 1 = what input 1
// keeps merged with what input 2 splits off (Rays::merge appends), output 2 the other way round; both outputs are apodised by
// the output aperture in the frame of output 1 and thresholded.  The scene graph itself is a tree (one input per node): this is
// the node-level call for a caller that holds two bundles -- either input may be NULL (the reference's Rays::default()).
extern "C" OpmStatus opm_scene_beam_splitter_two_inputs(OpmScene* sc, int32_t node, OpmRays* in1, OpmRays* in2, const OpmRayTraceConfig* cfg,
                                                        OpmRays* out1, OpmRays* out2) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size() || !out1 || !out2 || out1 == out2) return OPM_ERR_ARG;
    Node& n = *sc->nodes[node];
    if (n.d.type != OPM_NODE_BEAM_SPLITTER || !n.placed) return OPM_ERR_ARG;
    OpmRayTraceConfig def;
    opm_raytrace_config_default(&def);
    if (!cfg) cfg = &def;
    OpmStatus st = opm_rays_clear(out1);
    if (!st) st = opm_rays_clear(out2);
    if (st || (!in1 && !in2)) return st;  // no input: no output (mod.rs:169-171)
    OpmRays* split[2] = {nullptr, nullptr};  // what input k splits off
    OpmRays* in[2] = {in1, in2};
    OpmRays* keep[2] = {out1, out2};
    for (int k = 0; k < 2 && !st; ++k) {
        if (!in[k]) continue;
        st = opm_rays_copy(in[k], keep[k]);  // the reference clones the incoming bundle
        if (!st) st = acquire_rays(sc, opm_rays_len(in[k]) ? opm_rays_len(in[k]) : 1, &split[k]);
        if (st) break;
        std::vector<OpmOp> ops;
        ops.push_back(op_refract(n.s[0], nullptr, true, cfg->missed_surface_strategy, 0, nullptr, nullptr));
        ops.push_back(op_apodize(n.s[0].aperture, n.s[0].eff_iso));
        OpmOp sp;
        memset(&sp, 0, sizeof sp);
        sp.kind = OPM_OP_SPLIT; sp.u.split.ratio = n.d.p[0]; sp.u.split.split_out = split[k]; sp.u.split.spectrum = n.d.spectrum;
        ops.push_back(sp);
        OpmTraceStats ts;
        st = opm_trace_sequence(keep[k], ops.data(), (int32_t)ops.size(), 0, nullptr, &ts);
    }
    if (!st && split[1]) st = opm_rays_append(out1, split[1]);  // in_ray1.merge(&split2)
    if (!st && split[0]) st = opm_rays_append(out2, split[0]);  // in_ray2.merge(&split1)
    for (int k = 0; k < 2 && !st; ++k) {
        if (opm_rays_len(keep[k]) == 0) continue;
        std::vector<OpmOp> ops{op_apodize(n.s[1].aperture, n.s[1].eff_iso), op_threshold(cfg->min_energy_per_ray)};
        OpmTraceStats ts;
        st = opm_trace_sequence(keep[k], ops.data(), (int32_t)ops.size(), 0, nullptr, &ts);
    }
    for (OpmRays* r : split) if (r) sc->pool_rays.push_back(r);  // back to the scene's pool
    return st;
}

// FluenceDetector::node_report with FluenceEstimator::Voronoi, the node's DEFAULT (nodes/fluence_detector/mod.rs:61,91-135):
// HitMap::calc_combined_fluence_with_voronoi (hit_map/mod.rs:232-262) -- the bounding box of all hits of the surface, one estimate
// per rays hit map, i.e. per (bundle, bounce level), on that common grid, summed.  The reference adds nx x nx estimates (both axes
// take nr_of_points.0 points, rays_hit_map.rs:470,489) to an nx x ny matrix and panics unless nx == ny -- its own report asks for
// (100, 83): refused here with OPM_ERR_ARG instead of a panic.
static OpmStatus node_fluence_sites(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4], double* peak,
                                    double* total, bool helper_rays);
extern "C" OpmStatus opm_scene_node_fluence_voronoi(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                                    double* peak, double* total) {
    return node_fluence_sites(sc, node, nx, ny, map_host, lrtb, peak, total, false);
}
// FluenceDetector::node_report with FluenceEstimator::HelperRays after the trace of a bundle with fluence helper rays:
// HitMap::calc_combined_fluence_with_helper_rays (hit_map/mod.rs:283-310), same sum over the rays hit maps on the common box
extern "C" OpmStatus opm_scene_node_fluence_helper_rays(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4],
                                                        double* peak, double* total) {
    return node_fluence_sites(sc, node, nx, ny, map_host, lrtb, peak, total, true);
}
static OpmStatus node_fluence_sites(OpmScene* sc, int32_t node, uint64_t nx, uint64_t ny, double* map_host, double lrtb[4], double* peak,
                                    double* total, bool helper_rays) {
    if (!sc || node < 0 || node >= (int)sc->nodes.size() || nx == 0) return OPM_ERR_ARG;
    if (nx != ny) return OPM_ERR_ARG;  // "Matrix addition/subtraction dimensions mismatch" in the reference
    Node& n = *sc->nodes[node];
    std::vector<OpmHitMap*> maps;
    for (auto& h : n.s[0].hits) maps.push_back(h.map);
    if (maps.empty()) return OPM_ERR_EMPTY_HITMAP;
    double box[4];
    uint64_t n_hits = 0;
    OpmStatus st = opm_hitmaps_bounding_box(maps.data(), (int32_t)maps.size(), -1, box, &n_hits);  // get_bounding_box, hit_map/mod.rs:178-203
    if (st) return st;
    void* dev = nullptr;
    st = opm_device_alloc(sc->ctx, nx * nx * sizeof(double), &dev);
    if (st) return st;
    st = opm_device_memset(sc->ctx, dev, 0, nx * nx * sizeof(double));
    const double ranges[4] = {box[0], box[1], box[3], box[2]};  // ax_1 = left..right, ax_2 = bottom..top
    // one rays hit map per (bundle uuid, bounce level): the scene files a bundle's hits under its uuid, several passes of the same
    // bundle share it
    std::vector<uint64_t> uuids;
    for (auto& h : n.s[0].hits) { bool seen = false; for (uint64_t u : uuids) seen = seen || u == h.uuid; if (!seen) uuids.push_back(h.uuid); }
    for (uint64_t u : uuids) {
        std::vector<OpmHitMap*> mine;
        uint64_t levels = 0;
        for (auto& h : n.s[0].hits)
            if (h.uuid == u) { mine.push_back(h.map); uint64_t m = 0; if (!st) st = opm_hitmap_bounce_levels(h.map, &m); levels |= m; }
        for (int b = 0; b < 64 && !st; ++b)
            if ((levels >> b) & 1)
                st = helper_rays ? opm_fluence_helper_rays(mine.data(), (int32_t)mine.size(), b, nx, ranges, 1, (double*)dev, nullptr)
                                 : opm_fluence_voronoi(mine.data(), (int32_t)mine.size(), b, nx, ranges, 1, (double*)dev, nullptr);
    }
    if (!st) st = opm_fluence_peak_total(sc->ctx, (const double*)dev, nx, nx, box, peak, total);
    if (!st && map_host) st = opm_device_to_host(sc->ctx, map_host, dev, nx * nx * sizeof(double));
    if (lrtb) memcpy(lrtb, box, sizeof box);
    opm_device_free(sc->ctx, dev);
    return st;
}

// ---------------------------------------------------------------------------------------------
// ghost focus: see opm_scene_gf.inc
// ---------------------------------------------------------------------------------------------
#include "opm_scene_gf.inc"
