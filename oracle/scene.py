"""ORACLE scene layer (test infrastructure only): the reference's node -> surface builders, `pass_through_surface`,
the ray-trace walk and the ghost-focus pass loop, restated in Python over the C oracle's bundle functions.

Only O(#nodes) control flow lives here; every per-ray loop is in opm_oracle.c.  Written from the reference
(file:line cited, relative to /root/reference/opossum/src), independently of the product's C++ scene layer.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import oracle as orc

(SOURCE, LENS, CYLINDRIC_LENS, WEDGE, THIN_MIRROR, PARABOLIC_MIRROR, BEAM_SPLITTER, DUMMY, IDEAL_FILTER, PARAXIAL_SURFACE,
 SPOT_DIAGRAM, FLUENCE_DETECTOR, REFLECTIVE_GRATING, RAY_PROPAGATION_VISUALIZER, SPECTROMETER, WAVEFRONT, ENERGY_METER) = range(17)


@dataclass
class RayTraceConfig:  # analyzers/raytrace.rs:302-323
    min_energy_per_ray: float = 1.0 * 1.0e-12
    max_number_of_bounces: int = 1000
    max_number_of_refractions: int = 1000
    missed_surface_strategy: int = orc.MISS_STOP


@dataclass
class Bundle:  # struct Rays (rays.rs:62-73)
    rays: np.ndarray
    uuid: int
    parent: int = 0

    def clone(self) -> "Bundle":
        return Bundle(self.rays.copy(), self.uuid, self.parent)


class OSurface:
    """surface/optic_surface.rs:29-43"""

    def __init__(self, surf: orc.Surface, eff_iso: orc.Isometry, aperture: orc.Aperture, lidt: float):
        self.surf, self.eff_iso, self.aperture, self.lidt = surf, eff_iso, aperture, lidt
        self.reset()

    def reset(self):
        self.hit_map: dict[tuple[int, int], list[np.ndarray]] = {}
        self.fwd_cache: list[Bundle] = []
        self.bwd_cache: list[Bundle] = []
        self.critical: dict[int, tuple[float, int]] = {}
        self.peak_seen = -math.inf

    def add_hits(self, hits: np.ndarray, uuid: int):  # hit_map/mod.rs:101-115
        for b in np.unique(hits["bounce"]):
            self.hit_map.setdefault((int(b), uuid), []).append(hits[hits["bounce"] == b])

    def merged_hits(self) -> np.ndarray:  # hit_map/mod.rs:161-175
        parts = [h for lst in self.hit_map.values() for h in lst]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=orc.HIT_DTYPE)

    estimator = "binning"  # GhostFocusConfig::fluence_estimator of the running analysis ("binning" | "voronoi")

    def evaluate_fluence_of_ray_bundle(self, b: Bundle):  # optic_surface.rs:202-224 -> get_max_fluence (rays_hit_map.rs:777-783), 101x101
        lst = self.hit_map.get((orc.bounce_lvl(b.rays), b.uuid))
        if not lst:
            return
        try:
            if OSurface.estimator == "voronoi":
                fmap, _ = orc.fluence_voronoi(np.concatenate(lst), 101)
            else:
                fmap, _ = orc.fluence_binning(np.concatenate(lst), 101, 101)
        except orc.OracleError:
            return
        peak = orc.fluence_peak(fmap)
        self.peak_seen = max(self.peak_seen, peak)
        if peak > self.lidt:
            self.critical[b.uuid] = (peak, orc.bounce_lvl(b.rays))


def parabola_off_axis_isometry(focal_length, oa_angle=0.0, dir_x=1.0, dir_y=0.0, collimating=0.0):
    """ParabolicMirror::calc_off_axis_isometry / calc_parent_focal_length (nodes/parabolic_mirror.rs:286-320)"""
    if dir_x == 0.0 and dir_y == 0.0:
        dir_x = 1.0  # the default off-axis direction (parabolic_mirror.rs:64-70)
    import ctypes as C
    tot, parent_f = orc.Isometry(), C.c_double()
    orc.lib().orc_parabola_off_axis_isometry(C.c_double(focal_length), C.c_double(oa_angle), C.c_double(dir_x), C.c_double(dir_y),
                                             int(bool(collimating)), C.byref(tot), C.byref(parent_f))
    parent_f = parent_f.value
    return tot, parent_f


def _surface(geom, param, iso, coating):
    return orc.Surface(geom=geom, param=param, iso=iso, coating=coating[0], coating_r=coating[1])


class ONode:
    def __init__(self, node_type, iso=None, p=(), index=None, coating_in=None, coating_out=None, aperture_in=None,
                 aperture_out=None, alignment=None, lidt=1.0e4, parent=-1, parent_port=0, spectrum=None, distance=None, align_like=None):
        # align_like = (node index, distance): NodeAttr::align_like_node_at_distance (analysis_raytrace.rs:142-163)
        # distance: NodeGroup::connect_nodes(.., distance) of a node WITHOUT an isometry of its own -- OScene.calc_node_positions
        # places it (then `iso` is only a placeholder)
        self._args = dict(node_type=node_type, p=p, index=index, coating_in=coating_in, coating_out=coating_out, aperture_in=aperture_in,
                          aperture_out=aperture_out, alignment=alignment, lidt=lidt, parent=parent, parent_port=parent_port,
                          spectrum=spectrum)
        self.distance, self.placed, self.iso = distance, distance is None and align_like is None, iso or orc.Isometry.identity()
        self.align_like = align_like
        self.type, self.p, self.index = node_type, tuple(p), index
        self.spectrum = spectrum  # (slots [um], values): FilterType::Spectrum / SplittingConfig::Spectrum
        self.parent, self.parent_port = parent, parent_port
        mirror = node_type in (THIN_MIRROR, PARABOLIC_MIRROR)
        default_coat = (orc.COAT_CONSTANT_R, 1.0) if mirror else (orc.COAT_IDEAL_AR, 0.0)  # thin_mirror.rs:67-82
        coats = [coating_in or default_coat, coating_out or default_coat]
        aps = [aperture_in or orc.ap_none(), aperture_out or orc.ap_none()]
        node_iso = iso or orc.Isometry.identity()
        if alignment is not None:  # effective_node_iso (optic_node.rs:356-363)
            node_iso = node_iso.append(alignment)
        ident = orc.Isometry.identity()
        anchor = [ident, ident]
        T = orc.Isometry.new_along_z
        if node_type in (LENS, CYLINDRIC_LENS):  # nodes/lens/mod.rs:226-293, cylindric_lens/mod.rs:146-214
            curved = orc.GEOM_SPHERE if node_type == LENS else orc.GEOM_CYLINDER
            front, rear, thick = self.p
            if math.isinf(front):
                s0 = _surface(orc.GEOM_PLANE, 0.0, node_iso, coats[0])
            else:
                anchor[0] = T(front)
                s0 = _surface(curved, front, node_iso.append(anchor[0]), coats[0])
            if math.isinf(rear):
                anchor[1] = T(thick)
                s1 = _surface(orc.GEOM_PLANE, 0.0, node_iso.append(anchor[1]), coats[1])
            else:
                anchor[1] = T(rear + thick)
                s1 = _surface(curved, rear, node_iso.append(anchor[1]), coats[1])
        elif node_type == WEDGE:  # nodes/wedge/mod.rs:116-160
            thick, angle = self.p
            s0 = _surface(orc.GEOM_PLANE, 0.0, node_iso, coats[0])
            anchor[1] = T(thick).append(orc.Isometry.new((0.0, 0.0, 0.0), (angle, 0.0, 0.0)))
            s1 = _surface(orc.GEOM_PLANE, 0.0, node_iso.append(anchor[1]), coats[1])
        elif node_type == THIN_MIRROR:  # nodes/thin_mirror.rs:121-155
            (curv,) = self.p or (math.inf,)
            if math.isinf(curv):
                s0, s1 = (_surface(orc.GEOM_PLANE, 0.0, node_iso, c) for c in coats)
            else:
                anchor = [T(curv), T(curv)]
                s0, s1 = (_surface(orc.GEOM_SPHERE, curv, node_iso.append(anchor[0]), c) for c in coats)
        elif node_type == PARABOLIC_MIRROR:  # nodes/parabolic_mirror.rs:388-421: p = (f, oa angle, oa dir x, oa dir y, collimating)
            a_iso, parent_f = parabola_off_axis_isometry(*self.p)
            anchor = [a_iso, a_iso]
            s0, s1 = (_surface(orc.GEOM_PARABOLA, -1.0 * parent_f, node_iso.append(a_iso), c) for c in coats)
        else:  # optic_node.rs:62-80
            s0, s1 = (_surface(orc.GEOM_PLANE, 0.0, node_iso, c) for c in coats)
        self.s = [OSurface(s0, node_iso.append(anchor[0]), aps[0], lidt), OSurface(s1, node_iso.append(anchor[1]), aps[1], lidt)]
        self.stored: np.ndarray | None = None
        self.split: np.ndarray | None = None
        self.stored_hist = self.split_hist = None  # orc.History of those bundles (raytrace(history=True))

    def place(self, iso: "orc.Isometry"):
        """OpticNode::set_isometry: rebuild the surfaces at the given node isometry"""
        distance, align_like = self.distance, self.align_like
        self.__init__(iso=iso, **self._args)
        self.distance, self.align_like, self.placed = distance, align_like, True


class OScene:
    def __init__(self, nodes: list[ONode], ambient: orc.IndexModel | None = None, n_threads: int = 1):
        self.nodes = nodes
        self.ambient = ambient or orc.IndexModel.const(1.0)
        self.nt = n_threads
        self.interactions = 0
        self._uuid = 1
        self.passes: list[list[Bundle]] = []

    def reset_data(self):
        for n in self.nodes:
            for s in n.s:
                s.reset()
            n.stored = n.split = None
            n.stored_hist = n.split_hist = None
        self.passes = []
        self.interactions = 0
        self.chain_out: dict[int, np.ndarray] = {}

    # ---- shared helpers
    def _refract(self, rays, s: OSurface, idx, intended, strategy, uuid=1, record=True, want_children=True):
        ch, hits, st = orc.rays_refract_on_surface(rays, s.surf, idx, intended, strategy, self.nt, want_children=want_children,
                                                   want_hits=record)
        self.interactions += st.n_interactions
        if record:
            s.add_hits(hits, uuid)
        return ch

    def _rebind_history(self, parents: np.ndarray, clones: np.ndarray) -> np.ndarray:
        """the reflected / diffracted clones continue as the bundle (rays.rs:1016-1019): each carries the pos_hist of the
        ray it was cloned from (same id), the sink moves to the new array"""
        if self.hist is None:
            return clones
        row = {int(i): k for k, i in enumerate(parents["id"])}
        assert len(row) == len(parents), "history tracking needs unique ray ids within a bundle"
        orc.History.detach()
        self.hist = self.hist.select([row[int(i)] for i in clones["id"]])
        clones = np.ascontiguousarray(clones)
        self.hist.attach(clones)
        return clones

    def _children_of(self, i, port=0):
        return [k for k, n in enumerate(self.nodes) if k > 0 and (n.parent if n.parent >= 0 else k - 1) == i and n.parent_port == port]

    # ---- ray tracing (analyzers/raytrace.rs:36-49, nodes/node_group/analysis_raytrace.rs:29-91)
    def _analyze_rt(self, n: ONode, rays: np.ndarray, cfg: RayTraceConfig, record_all=True) -> np.ndarray:
        strat, thr = cfg.missed_surface_strategy, cfg.min_energy_per_ray
        t = n.type
        if t == SOURCE:  # nodes/source.rs:189-235
            orc.rays_transform(rays, n.s[0].eff_iso)
            orc.rays_apodize(rays, n.s[1].aperture, n.s[0].eff_iso, self.nt)
            orc.rays_threshold(rays, thr, self.nt)
        elif t in (LENS, CYLINDRIC_LENS, WEDGE):  # lens/analysis_raytrace.rs + pass_through_surface (raytrace.rs:96-140)
            for s, idx in ((n.s[0], n.index), (n.s[1], self.ambient)):
                self._refract(rays, s, idx, True, strat, record=record_all, want_children=False)
                orc.rays_apodize(rays, s.aperture, s.eff_iso, self.nt)
                orc.rays_threshold(rays, thr, self.nt)
        elif t in (THIN_MIRROR, PARABOLIC_MIRROR):  # thin_mirror.rs:209-250
            rays = self._rebind_history(rays, self._refract(rays, n.s[0], None, False, strat, record=record_all))
            orc.rays_apodize(rays, n.s[0].aperture, n.s[0].eff_iso, self.nt)
            orc.rays_threshold(rays, thr, self.nt)
        elif t == REFLECTIVE_GRATING:  # nodes/reflective_grating.rs:209-262
            iso = n.s[0].eff_iso
            gx = iso.transform_vector((1.0, 0.0, 0.0))
            gvec = [2. * math.pi * n.p[0] * c for c in gx]
            out, st = orc.rays_diffract_on_periodic_surface(rays, n.s[0].surf, orc.IndexModel.const(1.0), gvec, int(n.p[1]), False)
            rays = self._rebind_history(rays, out)
            self.interactions += st.n_interactions
            orc.rays_apodize(rays, n.s[0].aperture, iso, self.nt)
            orc.rays_threshold(rays, thr, self.nt)
        elif t == BEAM_SPLITTER:  # beam_splitter/mod.rs:158-293
            self._refract(rays, n.s[0], None, True, strat, record=record_all, want_children=False)
            orc.rays_apodize(rays, n.s[0].aperture, n.s[0].eff_iso, self.nt)
            split = orc.rays_split_spectrum(rays, *n.spectrum) if n.spectrum is not None else orc.rays_split(rays, n.p[0])
            orc.rays_apodize(rays, n.s[1].aperture, n.s[1].eff_iso, self.nt)
            orc.rays_threshold(rays, thr, self.nt)
            if self.hist is not None:  # the split-off clones carry the history up to here; their own pushes go to their copy
                n.split_hist = self.hist.copy()
                orc.History.detach()
                n.split_hist.attach(split)
            orc.rays_apodize(split, n.s[1].aperture, n.s[1].eff_iso, self.nt)
            orc.rays_threshold(split, thr, self.nt)
            if self.hist is not None:
                orc.History.detach()
                self.hist.attach(rays)
            n.split = split
        elif t in (DUMMY, IDEAL_FILTER, PARAXIAL_SURFACE):  # dummy.rs:83-135, ideal_filter.rs:203-250, paraxial_surface.rs:169-230
            self._refract(rays, n.s[0], None, True, strat, record=record_all, want_children=False)
            if t == IDEAL_FILTER:
                if n.spectrum is not None:
                    orc.rays_filter_energy_spectrum(rays, *n.spectrum)
                else:
                    orc.rays_filter_energy(rays, n.p[0])
            if t == PARAXIAL_SURFACE:
                orc.rays_refract_paraxial(rays, n.p[0], n.s[0].eff_iso)
            for ap in (n.s[0].aperture, n.s[1].aperture):
                orc.rays_apodize(rays, ap, n.s[0].eff_iso, self.nt)
                orc.rays_threshold(rays, thr, self.nt)
        else:  # detectors: pass_through_detector_surface (raytrace.rs:150-207)
            self._refract(rays, n.s[0], None, True, strat, record=True, want_children=False)
            orc.rays_apodize(rays, n.s[0].aperture, n.s[0].eff_iso, self.nt)
            orc.rays_threshold(rays, thr, self.nt)
            n.stored = rays.copy()
            n.stored_hist = self.hist.copy() if self.hist is not None else None
        return rays

    def beam_splitter_two_inputs(self, node: int, in1: np.ndarray | None, in2: np.ndarray | None, cfg: RayTraceConfig | None = None):
        """BeamSplitter::analyze_raytrace with both inputs (nodes/beam_splitter/mod.rs:158-293): each input passes the node's flat
        surface and input aperture and is split; output 1 = what input 1 keeps + what input 2 splits off, output 2 the other way
        round; both outputs see output 1's aperture frame and the energy threshold.  (The product's scene layer is a tree and takes
        one input; this restates the reference for its two-input test.)  Returns (out1, out2) or (None, None)."""
        cfg = cfg or RayTraceConfig()
        n = self.nodes[node]
        if in1 is None and in2 is None:
            return None, None
        empty = np.zeros(0, dtype=orc.RAY_DTYPE)

        def one(rays):
            if rays is None:
                return empty.copy(), empty.copy()
            rays = np.ascontiguousarray(rays.copy())
            self._refract(rays, n.s[0], None, True, cfg.missed_surface_strategy, record=False, want_children=False)
            orc.rays_apodize(rays, n.s[0].aperture, n.s[0].eff_iso, self.nt)
            split = orc.rays_split_spectrum(rays, *n.spectrum) if n.spectrum is not None else orc.rays_split(rays, n.p[0])
            return rays, split
        r1, sp1 = one(in1)
        r2, sp2 = one(in2)
        out1, out2 = np.concatenate([r1, sp2]), np.concatenate([r2, sp1])  # Rays::merge appends
        for out in (out1, out2):
            orc.rays_apodize(out, n.s[1].aperture, n.s[1].eff_iso, self.nt)
            orc.rays_threshold(out, cfg.min_energy_per_ray, self.nt)
        return out1, out2

    def _limits_rt(self, rays, cfg):  # node_group/analysis_raytrace.rs:20-27
        orc.rays_filter_bounces(rays, cfg.max_number_of_bounces)
        orc.rays_filter_refractions(rays, cfg.max_number_of_refractions)

    def raytrace(self, source_rays: np.ndarray, cfg: RayTraceConfig | None = None, record_all=True, history=False) -> np.ndarray:
        """history=True: every bundle's pos_hist is tracked (orc.History); see chain_hist, ONode.stored_hist"""
        cfg = cfg or RayTraceConfig()
        self.reset_data()
        self.chain_hist: dict[int, orc.History] = {}
        jobs = [(0, np.ascontiguousarray(source_rays.copy()), orc.History(len(source_rays)) if history else None)]
        main_out = None
        try:
            return self._raytrace_jobs(jobs, cfg, record_all)
        finally:
            if history:
                orc.lib().orc_history_detach()
            self.hist = None

    hist = None  # History of the bundle being traced

    def _raytrace_jobs(self, jobs, cfg, record_all):
        main_out = None
        while jobs:
            i, rays, self.hist = jobs.pop(0)
            if self.hist is not None:
                self.hist.attach(rays)
            start = i
            first = main_out is None
            while i is not None:
                n = self.nodes[i]
                rays = self._analyze_rt(n, rays, cfg, record_all)
                self._limits_rt(rays, cfg)
                if n.type == BEAM_SPLITTER:
                    self._limits_rt(n.split, cfg)
                    for k in self._children_of(i, 1):
                        jobs.append((k, np.ascontiguousarray(n.split.copy()), n.split_hist.copy() if n.split_hist is not None else None))
                nxt = self._children_of(i, 0)
                i = nxt[0] if nxt else None
            self.chain_out[start] = rays  # final bundle of every chain, keyed by its first node
            if self.hist is not None:
                orc.History.detach()
                self.chain_hist[start] = self.hist
            if first:
                main_out = rays
        return main_out

    # ---- node positioning (nodes/node_group/analysis_raytrace.rs:92-212, optic_graph.rs:878-926)
    def calc_node_positions(self, wavelength: float, cfg: RayTraceConfig | None = None) -> list:
        """Places every node that has no isometry of its own at its edge distance behind its predecessor along the optical
        axis: ONE axis ray (Rays::get_optical_axis_ray: origin, +z, `wavelength`, 1 J; nodes/source.rs:236-268) is passed from
        node to node with each node's own `analyze` (beam splitter: only its input surface, both outputs carry the ray,
        beam_splitter/analysis_raytrace.rs:40-93); an unplaced node gets Isometry::new_from_view(ray.propagate(distance),
        up) (optic_graph.rs:903-905); `up` starts as Ray::define_up_direction of the source's axis ray and is rotated by every
        node's last direction change (ray.rs:842-860).  Walk order = node index order.  Returns the node isometries."""
        cfg = cfg or RayTraceConfig()
        self.reset_data()
        self.hist = None
        out: dict[tuple[int, int], tuple[np.ndarray, np.ndarray | None]] = {}
        up = np.array([0.0, 1.0, 0.0])  # analysis_raytrace.rs:103
        for i, n in enumerate(self.nodes):
            if n.type == SOURCE:
                ray = orc.ray_new((0.0, 0.0, 0.0), (0.0, 0.0, 1.0), wavelength, 1.0)
                orc.rays_transform(ray, n.s[0].eff_iso)  # axis_ray.transformed_ray(effective_surface_iso("input_1"))
                out[(i, 0)] = (ray, None)
                up = orc.define_up_direction(ray["dir"][0])
                continue
            parent = n.parent if n.parent >= 0 else i - 1
            ray, prev_in = out[(parent, n.parent_port)]
            ray = ray.copy()
            if not n.placed and n.align_like is not None and self.nodes[n.align_like[0]].placed:
                n.place(self.nodes[n.align_like[0]].iso.append(orc.Isometry.new_along_z(n.align_like[1])))
            if not n.placed:
                r = ray[0]
                pos = r["pos"] + n.distance * r["dir"]  # Ray::propagate (ray.rs:408-427)
                n.place(orc.Isometry.from_view(pos, r["dir"], up))
            prev_dir = prev_in  # a field of the ray: survives a node whose surface the ray does not reach

            def refract(rays, s, idx, intended):
                nonlocal prev_dir
                before = rays[0].copy()
                ch = self._refract(rays, s, idx, intended, cfg.missed_surface_strategy, record=False)
                # prev_dir is set whenever the surface was reached (ray.rs:628-631,676), also on total reflection
                moved = before["valid"] and (len(ch) > 0 or rays[0]["bounces"] != before["bounces"])
                if moved:
                    prev_dir = before["dir"].copy()
                return ch

            t = n.type
            if t == BEAM_SPLITTER:
                refract(ray, n.s[0], None, True)
                out[(i, 0)] = out[(i, 1)] = (ray, prev_dir)
                edges = 2
            else:
                if t in (LENS, CYLINDRIC_LENS, WEDGE):
                    for s, idx in ((n.s[0], n.index), (n.s[1], self.ambient)):
                        refract(ray, s, idx, True)
                        orc.rays_apodize(ray, s.aperture, s.eff_iso, self.nt)
                        orc.rays_threshold(ray, cfg.min_energy_per_ray, self.nt)
                elif t in (THIN_MIRROR, PARABOLIC_MIRROR):
                    ray = np.ascontiguousarray(refract(ray, n.s[0], None, False))
                    orc.rays_apodize(ray, n.s[0].aperture, n.s[0].eff_iso, self.nt)
                    orc.rays_threshold(ray, cfg.min_energy_per_ray, self.nt)
                elif t == REFLECTIVE_GRATING:
                    raise NotImplementedError("axis ray through a grating")
                else:
                    refract(ray, n.s[0], None, True)
                    if t == PARAXIAL_SURFACE:
                        if ray[0]["valid"]:
                            prev_dir = ray[0]["dir"].copy()  # ray.rs:450
                        orc.rays_refract_paraxial(ray, n.p[0], n.s[0].eff_iso)
                    if t == IDEAL_FILTER:
                        orc.rays_filter_energy(ray, n.p[0])
                    orc.rays_apodize(ray, n.s[0].aperture, n.s[0].eff_iso, self.nt)
                    orc.rays_threshold(ray, cfg.min_energy_per_ray, self.nt)
                self._limits_rt(ray, cfg)
                out[(i, 0)] = (ray, prev_dir)
                edges = 1
            if len(ray) == 0:
                raise RuntimeError("no rays in this ray bundle. cannot position nodes")
            for _ in range(edges):  # calc_new_up_direction once per outgoing edge (analysis_raytrace.rs:199-208)
                if prev_dir is None:
                    raise RuntimeError("no previous direction of ray defined to calculate new up-direction!")
                up = orc.calc_new_up_direction(prev_dir, ray[0]["dir"], up)
        self.reset_data()
        return [n.iso for n in self.nodes]

    def node_fluence(self, node: int, nx=100, ny=83):  # fluence_detector/mod.rs:91-135 (Binning)
        fmap, lrtb = orc.fluence_binning(self.nodes[node].s[0].merged_hits(), nx, ny)
        return fmap, lrtb, orc.fluence_peak(fmap), orc.fluence_total_energy(fmap, lrtb)

    def node_energy(self, node: int) -> float:  # nodes/energy_meter.rs:136-151
        return orc.total_energy(self.nodes[node].stored)

    def node_spectrum(self, node: int, resolution=0.2e-9):  # nodes/spectrometer.rs:155-175
        return orc.rays_to_spectrum(self.nodes[node].stored, resolution)

    def node_wavefront(self, node: int):  # nodes/wavefront.rs:173-235, rays.rs:713-726
        n = self.nodes[node]
        wvl = orc.spectrum_center_wavelength(*orc.rays_to_spectrum(n.stored, 1.0e-9))
        xyw = orc.rays_wavefront_error(n.stored, wvl, n.s[0].eff_iso)
        ptv, rms = orc.wavefront_statistics(xyw)
        return {"wavelength": wvl, "ptv": ptv, "rms": rms, "map": xyw}

    def node_spot(self, node: int):  # spot_diagram.rs:101-163
        n = self.nodes[node]
        loc = n.stored.copy()
        orc.rays_transform(loc, n.s[0].eff_iso, inverse=True)
        return {"centroid": orc.energy_weighted_centroid(loc), "geo": orc.beam_radius_geo(loc),
                "rms": orc.energy_weighted_beam_radius_rms(loc), "n_valid": orc.nr_of_rays(loc),
                "total_energy": orc.total_energy(loc)}

    # ---- ghost focus (analyzers/ghostfocus.rs:82-122, node_group/analysis_ghostfocus.rs:22-109)
    def _pass_through_surface_gf(self, s: OSurface, idx, bundles: list[Bundle], backward: bool):  # raytrace.rs:96-140
        for b in bundles:
            ch = self._refract(b.rays, s, idx, True, orc.MISS_IGNORE, uuid=b.uuid)
            child = Bundle(ch, self._new_uuid(), b.uuid)
            s.evaluate_fluence_of_ray_bundle(b)
            (s.fwd_cache if backward else s.bwd_cache).append(child)  # add_to_rays_cache(reflected, backward)
            orc.rays_apodize(b.rays, s.aperture, s.eff_iso, self.nt)
        for c in (s.bwd_cache if backward else s.fwd_cache):
            bundles.append(c.clone())

    def _new_uuid(self):
        self._uuid += 1
        return self._uuid

    def _analyze_gf(self, n: ONode, bundles: list[Bundle], inverted: bool, gf_pass: int, source_rays) -> list[Bundle]:
        i_in, i_out = (1, 0) if inverted else (0, 1)
        t = n.type
        if t == SOURCE:  # source.rs:273-338
            if not inverted:
                bundles = []
                if gf_pass == 0:
                    r = source_rays.copy()
                    orc.rays_transform(r, n.s[1].eff_iso)
                    bundles = [Bundle(r, self._new_uuid())]
            for b in bundles:
                self._refract(b.rays, n.s[0], None, True, orc.MISS_IGNORE, uuid=b.uuid)
                n.s[0].evaluate_fluence_of_ray_bundle(b)
        elif t in (LENS, CYLINDRIC_LENS, WEDGE):  # lens/analysis_ghostfocus.rs:13-50
            self._pass_through_surface_gf(n.s[i_in], n.index, bundles, inverted)
            self._pass_through_surface_gf(n.s[i_out], self.ambient, bundles, inverted)
        elif t in (THIN_MIRROR, PARABOLIC_MIRROR):  # thin_mirror.rs:159-199
            s = n.s[i_in]
            for b in bundles:
                b.rays = self._refract(b.rays, s, None, False, orc.MISS_IGNORE, uuid=b.uuid)
                orc.rays_apodize(b.rays, s.aperture, s.eff_iso, self.nt)
                orc.rays_threshold(b.rays, 1.0 * 1.0e-12, self.nt)
            for b in bundles:
                s.evaluate_fluence_of_ray_bundle(b)
        elif t == BEAM_SPLITTER:  # beam_splitter/analysis_ghostfocus.rs:12-84 (in1 -> out1)
            s = n.s[i_in]
            for b in bundles:
                self._refract(b.rays, s, None, True, orc.MISS_IGNORE, uuid=b.uuid)
                orc.rays_apodize(b.rays, s.aperture, s.eff_iso, self.nt)
                orc.rays_split(b.rays, n.p[0])
                orc.rays_apodize(b.rays, n.s[i_out].aperture, n.s[i_out].eff_iso, self.nt)
                s.evaluate_fluence_of_ray_bundle(b)
        else:  # analyze_single_surface_node (ghostfocus.rs:250-281) / paraxial_surface.rs:95-150 / ideal_filter.rs:165-185
            s = n.s[i_in]
            for b in bundles:
                self._refract(b.rays, s, None, True, orc.MISS_IGNORE, uuid=b.uuid)
                if t == PARAXIAL_SURFACE:
                    orc.rays_refract_paraxial(b.rays, n.p[0], s.eff_iso)
                orc.rays_apodize(b.rays, s.aperture, s.eff_iso, self.nt)
                s.evaluate_fluence_of_ray_bundle(b)
                if t == IDEAL_FILTER:
                    orc.rays_filter_energy(b.rays, n.p[0])
        return bundles

    def ghostfocus(self, source_rays: np.ndarray, max_bounces: int = 1, estimator: str = "binning"):
        self.reset_data()
        OSurface.estimator = estimator
        chain, i = [], 0
        while i is not None:
            chain.append(i)
            nxt = self._children_of(i, 0)
            i = nxt[0] if nxt else None
        for bounce in range(max_bounces + 1):  # ghostfocus.rs:100-120
            inverted = bounce % 2 == 1
            bundles: list[Bundle] = []
            for i in (reversed(chain) if inverted else chain):
                bundles = self._analyze_gf(self.nodes[i], bundles, inverted, bounce, source_rays)
                for b in bundles:  # filter_ray_limits (analysis_ghostfocus.rs:14-20)
                    orc.rays_filter_bounces(b.rays, max_bounces)
            self.passes.append(bundles)
        return self.passes
