"""The force step with the interaction lists built on the GPU: the host hands
over the sorted particles and the tree topology; moments (cb200_build_moments), lists
(cb200_walk_device) and forces are all computed on the device, so per step only
~90 bytes per particle cross PCIe instead of the 8-byte-per-entry lists (~500 bytes per particle
at theta = 0.7).  torch is used for device buffers and H2D/D2H copies only."""
import ctypes as C

import numpy as np


class DeviceTreeStep:
    def __init__(self, hc, tree, theta=0.7, n_replicas=0, period=1.0, ewald=None, bucket_range=None,
                 rung=None, active_rung=0):
        """tree: changa_b200.tree.Tree (host-built topology).  ewald: None | dict(dEwCut, dEwhCut).
        rung (one byte per particle, TREE order) + active_rung: multistep step, see RawParticleStep."""
        import torch
        assert hc.L.cb200_real_bytes() == 4, "DeviceTreeStep holds float32 device buffers: use the float build"
        self.torch, self.hc, self.t = torch, hc, tree
        self.theta, self.nrep, self.period = float(theta), int(n_replicas), float(period)
        self.ewald = ewald
        self.range = bucket_range or (0, tree.num_buckets)
        self.ext = torch.cuda.Stream()          # torch owns it: its allocators record events on it at teardown
        self.stream = self.ext.cuda_stream
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        t = tree
        # what the host must provide every step (pinned, as TreePiece / DataManager would hold it)
        self.h = {
            "pos": pin(t.parts[:, 2:5]), "mass": pin(t.parts[:, 0]), "soft": pin(t.parts[:, 1]),
            "child0": pin(t.child0), "child1": pin(t.child1), "parent": pin(t.parent),
            "first": pin(t.first), "last": pin(t.last), "bfirst": pin(t.bucket_first), "bcount": pin(t.bucket_count),
            "bnode": pin(t.bucket_node), "geolo": pin(t.geolo), "geohi": pin(t.geohi),
            "boxlo": pin(t.boxlo), "boxhi": pin(t.boxhi),
            "parts32": pin(t.parts.astype(np.float32)),
        }
        self.active_rung = int(active_rung)
        if rung is not None:
            self.h["rung"] = pin(np.asarray(rung, dtype=np.uint8))
            self.h["bstarts"] = pin(np.asarray(t.bucket_starts, dtype=np.int32))
            self.h["bsizes"] = pin(np.asarray(t.bucket_sizes, dtype=np.int32))
        self.level_start = np.ascontiguousarray(t.level_start, dtype=np.int32)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.h.values())
        self.out = torch.zeros((t.n, 5), dtype=torch.float32).pin_memory()
        self.d2h_bytes = self.out.numel() * 4
        self.lists_info = None

    def run(self, keep_lists=False, phases=None):
        """phases: optional dict filled with the milliseconds of each phase (CUDA events on the
        step's stream: h2d, moments, walk, pack, ewald, forces, d2h)"""
        torch, hc, t, L = self.torch, self.hc, self.t, self.hc.L
        s = self.stream
        marks = []

        def mark(name):
            if phases is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.ext)
                marks.append((name, e))

        with torch.cuda.stream(self.ext):
            nn, nb, n = t.num_nodes, t.num_buckets, t.n
            if getattr(self, "dev", None) is None:  # device buffers live as long as the step object
                self.dev = {k: torch.empty_like(v, device="cuda") for k, v in self.h.items()}
                self.dev["mom32"] = torch.empty((nn, 27), dtype=torch.float32, device="cuda")
                self.dev["mom64"] = torch.empty((nn, 27), dtype=torch.float64, device="cuda")
                self.dev["pk_parts"] = torch.empty(n * L.cb200_packed_particle_bytes(), dtype=torch.uint8, device="cuda")
                self.dev["pk_mom"] = torch.empty(nn * L.cb200_packed_moment_bytes(), dtype=torch.uint8, device="cuda")
                self.dev["vars"] = torch.empty((n, 5), dtype=torch.float32, device="cuda")
            d = self.dev
            mark("start")
            for k, v in self.h.items():
                d[k].copy_(v, non_blocking=True)
            mark("h2d")
            mom32, mom64 = d["mom32"], d["mom64"]
            L.cb200_build_moments(d["pos"].data_ptr(), d["mass"].data_ptr(), d["soft"].data_ptr(), n,
                                  d["child0"].data_ptr(), d["child1"].data_ptr(), d["first"].data_ptr(),
                                  d["last"].data_ptr(), d["geolo"].data_ptr(), d["geohi"].data_ptr(),
                                  d["boxlo"].data_ptr(), d["boxhi"].data_ptr(), self.level_start.ctypes.data,
                                  t.num_levels, nn, mom32.data_ptr(), mom64.data_ptr(), s)
            mark("moments")
            lists = hc.T.Lists()
            active_ptr, self.n_act = None, n
            if "rung" in self.h:
                if "bucket_active" not in d:
                    d["bucket_active"] = torch.empty(nb, dtype=torch.uint8, device="cuda")
                    d["markers"] = torch.empty(n, dtype=torch.int32, device="cuda")
                counts = (C.c_int * 2)()
                L.cb200_active_sets_device(d["rung"].data_ptr(), None, n, d["bstarts"].data_ptr(),
                                           d["bsizes"].data_ptr(), nb, self.active_rung,
                                           d["bucket_active"].data_ptr(), d["markers"].data_ptr(), counts, s)
                active_ptr, self.n_act = d["bucket_active"].data_ptr(), int(counts[1])
                self.active = {"buckets": int(counts[0]), "particles": self.n_act}
            L.cb200_walk_device_active(nn, nb, t.num_levels, self.level_start.ctypes.data, d["child0"].data_ptr(),
                                       d["child1"].data_ptr(), d["parent"].data_ptr(), d["first"].data_ptr(),
                                       d["last"].data_ptr(), d["bfirst"].data_ptr(), d["bcount"].data_ptr(),
                                       d["bnode"].data_ptr(), d["boxlo"].data_ptr(), d["boxhi"].data_ptr(),
                                       mom64.data_ptr(), self.theta, self.nrep, self.period, self.range[0],
                                       self.range[1], active_ptr, C.byref(lists), s)
            mark("walk")
            if lists.error:
                raise RuntimeError(f"device walk: per-node capacity exceeded (error {lists.error})")
            pk_parts, pk_mom, vars_ = d["pk_parts"], d["pk_mom"], d["vars"]
            L.cb200_pack_particles_device(d["parts32"].data_ptr(), pk_parts.data_ptr(), n, s)
            L.cb200_pack_moments_device(mom32.data_ptr(), pk_mom.data_ptr(), nn, s)
            L.cb200_zero_vars_device(vars_.data_ptr(), n, s)
            P, V, M = pk_parts.data_ptr(), vars_.data_ptr(), pk_mom.data_ptr()
            fper = self.period if (self.nrep or self.ewald is not None) else 0.0
            mark("pack")
            if self.ewald is not None:
                self._ewald(mom64, P, V, s)
            mark("ewald")
            mx = int(t.bucket_sizes.max())
            L.cb200_cell_list_device_ex(P, V, M, lists.d_cell, lists.d_cellMarkers, lists.d_starts, lists.d_sizes,
                                        nb, fper, mx, s)
            L.cb200_part_list_device_ex(P, V, P, lists.d_part, lists.d_partMarkers, lists.d_starts, lists.d_sizes,
                                        nb, fper, mx, s)
            if lists.nSoft:
                L.cb200_part_list_device_ex(P, V, lists.d_nodeParticles, lists.d_soft, lists.d_softMarkers,
                                            lists.d_starts, lists.d_sizes, nb, fper, mx, s)
            mark("forces")
            self.out.copy_(vars_, non_blocking=True)
            mark("d2h")
            self.lists_info = {"nCell": int(lists.nCell), "nSoft": int(lists.nSoft), "nPart": int(lists.nPart)}
            if keep_lists:
                self.kept = self._download(lists, nb)
            L.cb200_lists_free(C.byref(lists), s)
            hc.stream_synchronize(s)
        if phases is not None:
            for (_, a), (name, b) in zip(marks, marks[1:]):
                phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
        return self.out.numpy()

    def _ewald(self, mom64, P, V, s):
        """root moments come back (27 doubles), the h-table is built on the host (EwaldInit, Ewald.cpp:285-375)"""
        from .tree import ewald_tables_fast as ewald_tables
        hc, t = self.hc, self.t
        self.hc.stream_synchronize(s)
        root = mom64[0].cpu().numpy()
        momc, ewt = ewald_tables(root, self.period, self.ewald.get("dEwhCut", 2.8))
        b0, b1 = self.range
        first, last = int(t.bucket_starts[b0]), int(t.bucket_starts[b1 - 1] + t.bucket_sizes[b1 - 1] - 1)
        if getattr(self, "_ew", None) is None:
            self._ew = hc.EwaldHostMemorySetup(1, len(ewt), 0)
        hc.fill_ewald(self._ew, root, momc, ewt, self.period, float(self.ewald.get("dEwCut", 2.6)), self.nrep,
                      active=None, first=first, last=last)
        if "rung" in self.h:  # large-phase form: device markers of my particle range
            torch = self.torch
            mk = self.dev["markers"][:self.n_act]
            i0, i1 = torch.searchsorted(mk, torch.tensor([first, last + 1], dtype=torch.int32, device="cuda")).tolist()
            if i1 > i0:
                hc.L.cb200_ewald_device(P, V, mk.data_ptr() + 4 * i0, i1 - i0, self._ew.cachedData, self._ew.ewt, s)
            return
        # small-phase form: a contiguous particle range, no marker array
        hc.L.cb200_EwaldHost(P, V, C.byref(self._ew), s, None, 0, 0)

    def _download(self, lists, nb):
        torch = self.torch
        def arr(ptr, n, cols):
            if n == 0:
                return np.zeros((0, cols) if cols > 1 else (0,), dtype=np.int32)
            buf = torch.empty(n * cols, dtype=torch.int32, device="cuda")
            self.hc.L.cb200_copy_device(buf.data_ptr(), ptr, n * cols * 4, self.stream)
            self.hc.stream_synchronize(self.stream)
            a = buf.cpu().numpy()
            return a.reshape(n, cols) if cols > 1 else a
        return {"cell": arr(lists.d_cell, int(lists.nCell), 2), "soft": arr(lists.d_soft, int(lists.nSoft), 2),
                "part": arr(lists.d_part, int(lists.nPart), 2), "cell_mark": arr(lists.d_cellMarkers, nb + 1, 1),
                "soft_mark": arr(lists.d_softMarkers, nb + 1, 1), "part_mark": arr(lists.d_partMarkers, nb + 1, 1),
                "starts": arr(lists.d_starts, nb, 1), "sizes": arr(lists.d_sizes, nb, 1)}

    def free(self):
        if getattr(self, "_ew", None) is not None:
            self.hc.EwaldHostMemoryFree(self._ew, 0)
            self._ew = None
        self.h = self.out = self.dev = None


class RawParticleStep:
    """The whole force step from UNSORTED particles (SURVEY f1 + f2): per step the host hands over
    positions, masses and softenings (40 bytes per particle) and gets accelerations back in its
    own particle order.  A thin view of the in-library step (changa_b200.step.NativeStep ->
    cb200_step_run): upload, NCCL all-gather, keys, sort, tree, moments, lists, forces, Ewald and the
    copy back all happen inside libchanga_b200.so; no torch on this path."""

    def __init__(self, hc, pos, mass, soft, theta=0.7, n_replicas=0, period=1.0, ewald=None, max_bucket=12,
                 root_lo=(-0.5, -0.5, -0.5), root_hi=(0.5, 0.5, 0.5), comm=None, rung=None, active_rung=0,
                 overlap_ewald=True, cost_cuts=True):
        """rung (one byte per particle, caller order) + active_rung: a multistep force step
         -- only buckets holding a particle with rung >= active_rung get lists and
        forces (Compute.cpp:1278,1574), only particles with rung >= active_rung get the Ewald sum
        (Ewald.cpp:416-437).  Rows of inactive buckets come back zero.

        comm (changa_b200.step.Comm, world > 1): one process per GPU; every rank holds rows
        [rank*chunk, (rank+1)*chunk) of the particle set on its host, ONE all-gather per step
        replicates the 40-byte records, every rank builds the same tree and moments, then walks,
        evaluates and returns only its own contiguous SFC range of buckets.  run() then returns
        (caller indices, rows)."""
        from .step import NativeStep
        self.hc = hc
        self.n = len(pos)
        self.st = NativeStep(hc, self.n, theta=theta, n_replicas=n_replicas, period=period, ewald=ewald,
                             max_bucket=max_bucket, root_lo=root_lo, root_hi=root_hi, comm=comm,
                             active_rung=active_rung if rung is not None else 0, overlap_ewald=overlap_ewald,
                             cost_cuts=cost_cuts)
        self.world, self.rank = self.st.world, self.st.rank
        self.st.set_particles(pos, mass, soft, rung)
        self.info = None
        self.h2d_bytes = self.d2h_bytes = 0

    def update_positions(self, pos):
        """new positions (caller order) for the next run(): an integrator's drift"""
        lo, hi = self.st.my_rows()
        self.st.rec.array[: hi - lo, :3] = np.asarray(pos, dtype=np.float64)[lo:hi]

    def run(self, phases=None, keep_tree=False, count_pairs=False, keep_lists=False):
        """phases: optional dict, the milliseconds of each phase are ADDED to it (CUDA events on the step's
        stream).  keep_tree: leave the downloaded tree arrays in self.kept_tree."""
        res = self.st.run(keep_lists=keep_lists)
        if phases is not None:
            for k, v in self.st.phases().items():
                phases[k] = phases.get(k, 0.0) + v
        self.range = (res.bucketLo, res.bucketHi, res.partLo, res.partHi)
        self.info = {"nodes": res.numNodes, "buckets": res.numBuckets, "levels": res.numLevels, "nCell": int(res.nCell),
                     "nSoft": int(res.nSoft), "nPart": int(res.nPart), "pc_pairs": int(res.pcPairs),
                     "pp_pairs": int(res.ppPairs)}
        if self.st.cfg.activeRung > 0:
            self.active = {"buckets": res.activeBuckets, "particles": res.activeParticles}
        self.h2d_bytes, self.d2h_bytes = int(res.h2dBytes), int(res.d2hBytes)
        if keep_tree:
            self.kept_tree = self.st.tree()
        if self.world > 1:
            return self.st.idx.array[: res.rows], self.st.out.array[: res.rows]
        return self.st.out.array[: self.n]

    def free(self):
        self.st.free()
