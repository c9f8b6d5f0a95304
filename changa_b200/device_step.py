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
    own particle order; keys, sort, tree topology, boxes, moments, interaction lists, forces and
    the Ewald sum all run on the device (cb200_build_tree, cb200_build_moments,
    cb200_walk_device, cb200_*_list_device_ex, cb200_EwaldHost)."""

    def __init__(self, hc, pos, mass, soft, theta=0.7, n_replicas=0, period=1.0, ewald=None, max_bucket=12,
                 root_lo=(-0.5, -0.5, -0.5), root_hi=(0.5, 0.5, 0.5), dist=None, rank=0, world=1,
                 rung=None, active_rung=0):
        """rung (one byte per particle, caller order) + active_rung: a multistep force step
         -- only buckets holding a particle with rung >= active_rung get lists and
        forces (Compute.cpp:1278,1574), only particles with rung >= active_rung get the Ewald sum
        (Ewald.cpp:416-437); the sets are made on the device (cb200_active_sets_device) and the
        walk is cb200_walk_device_active.  Rows of inactive buckets come back zero.

        world > 1 (one process per GPU, torch.distributed `dist`): every rank holds rows
        [rank*chunk, (rank+1)*chunk) of the particle set on its host; ONE all-gather per step
        replicates the 40-byte records, every rank builds the same tree and moments, then walks,
        evaluates and returns only its own contiguous SFC range of buckets (equal particle counts;
        accelerations never leave the owning GPU).  run() then returns (caller indices, rows)."""
        import torch
        assert hc.L.cb200_real_bytes() == 4, "RawParticleStep holds float32 device buffers: use the float build"
        self.torch, self.hc = torch, hc
        self.dist, self.rank, self.world = dist, int(rank), int(world)
        self.theta, self.nrep, self.period = float(theta), int(n_replicas), float(period)
        self.ewald, self.max_bucket = ewald, int(max_bucket)
        self.lo = np.ascontiguousarray(root_lo, dtype=np.float64)
        self.hi = np.ascontiguousarray(root_hi, dtype=np.float64)
        self.ext = torch.cuda.Stream()
        self.stream = self.ext.cuda_stream
        n = len(pos)
        self.n = n
        # {x, y, z, mass, soft}: the 40-byte record that crosses PCIe and NVLink; only this rank's rows
        self.chunk = -(-n // self.world)
        mine = np.zeros((self.chunk, 5))
        mine[:, 4] = 1.0
        # pad rows (only when world does not divide n) never reach the tree: the gathered array is
        # cut back to n rows
        lo_r = self.rank * self.chunk
        hi_r = min(n, lo_r + self.chunk)
        k = max(0, hi_r - lo_r)
        mine[:k, :3] = np.asarray(pos)[lo_r:hi_r]
        mine[:k, 3] = np.broadcast_to(mass, (n,))[lo_r:hi_r]
        mine[:k, 4] = np.broadcast_to(soft, (n,))[lo_r:hi_r]
        self.h = {"rec": torch.from_numpy(mine).pin_memory()}
        self.h2d_bytes = self.h["rec"].numel() * 8
        self.active_rung = int(active_rung)
        if rung is not None:
            r = np.zeros(self.chunk, dtype=np.uint8)
            r[:k] = np.asarray(rung, dtype=np.uint8)[lo_r:hi_r]
            self.h["rung"] = torch.from_numpy(r).pin_memory()
            self.h2d_bytes += self.chunk
        rows = n if self.world == 1 else 2 * self.chunk + 64
        self.out = torch.zeros((rows, 5), dtype=torch.float32).pin_memory()
        self.out_idx = torch.zeros(rows, dtype=torch.int32).pin_memory()
        self.d2h_bytes = self.out.numel() * 4
        self.dev = None
        self.info = None
        self._ew = None

    def update_positions(self, pos):
        """new positions (caller order) for the next run(): an integrator's drift.  world == 1 only."""
        assert self.world == 1
        self.h["rec"][: self.n, :3] = self.torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float64))

    def run(self, phases=None, keep_tree=False, count_pairs=False):
        """count_pairs: also leave {pc_pairs, pp_pairs} (sum over buckets of list length x bucket
        size, as Compute.cpp:1643-1651 counts interactions) in self.info"""
        torch, hc, L, s, n = self.torch, self.hc, self.hc.L, self.stream, self.n
        marks = []

        def mark(name):
            if phases is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.ext)
                marks.append((name, e))

        with torch.cuda.stream(self.ext):
            if self.dev is None:
                self.dev = {"rec": torch.empty_like(self.h["rec"], device="cuda"),
                            "all": torch.empty((self.chunk * self.world, 5), dtype=torch.float64, device="cuda"),
                            "pos": torch.empty((n, 3), dtype=torch.float64, device="cuda"),
                            "mass": torch.empty(n, dtype=torch.float64, device="cuda"),
                            "soft": torch.empty(n, dtype=torch.float64, device="cuda"),
                            "vars": torch.empty((n, 5), dtype=torch.float32, device="cuda")}
                if self.world == 1:
                    self.dev["out"] = torch.empty((n, 5), dtype=torch.float32, device="cuda")
            d = self.dev
            multistep = "rung" in self.h
            if multistep and "rung" not in d:
                d["rung"] = torch.empty_like(self.h["rung"], device="cuda")
                d["rung_all"] = torch.empty(self.chunk * self.world, dtype=torch.uint8, device="cuda")
                d["markers"] = torch.empty(n, dtype=torch.int32, device="cuda")
            mark("start")
            d["rec"].copy_(self.h["rec"], non_blocking=True)
            if multistep:
                d["rung"].copy_(self.h["rung"], non_blocking=True)
            mark("h2d")
            if self.world > 1:
                self.dist.all_gather_into_tensor(d["all"], d["rec"])
                full = d["all"]
                if multistep:
                    self.dist.all_gather_into_tensor(d["rung_all"], d["rung"])
            else:
                full = d["rec"]
                if multistep:
                    d["rung_all"] = d["rung"]
            d["pos"].copy_(full[:n, :3])
            d["mass"].copy_(full[:n, 3])
            d["soft"].copy_(full[:n, 4])
            mark("gather")
            tr = hc.T.DevTree()
            L.cb200_build_tree(d["pos"].data_ptr(), d["mass"].data_ptr(), d["soft"].data_ptr(), n, self.max_bucket,
                               self.lo.ctypes.data, self.hi.ctypes.data, C.byref(tr), s)
            if tr.error:
                raise RuntimeError("device tree build: node capacity exceeded")
            mark("tree")
            nn, nb = tr.numNodes, tr.numBuckets
            key = (nn,)
            if d.get("key") != key:  # node-sized buffers follow the tree
                d["mom32"] = torch.empty((nn, 27), dtype=torch.float32, device="cuda")
                d["mom64"] = torch.empty((nn, 27), dtype=torch.float64, device="cuda")
                d["pk_mom"] = torch.empty(nn * L.cb200_packed_moment_bytes(), dtype=torch.uint8, device="cuda")
                d["key"] = key
            mom32, mom64 = d["mom32"], d["mom64"]
            lvl = C.addressof(tr) + hc.T.DevTree.levelStart.offset
            L.cb200_build_moments(tr.d_pos, tr.d_mass, tr.d_soft, n, tr.d_child0, tr.d_child1, tr.d_first, tr.d_last,
                                  tr.d_geolo, tr.d_geohi, tr.d_boxlo, tr.d_boxhi, lvl, tr.numLevels, nn,
                                  mom32.data_ptr(), mom64.data_ptr(), s)
            mark("moments")
            b0, b1, p0, p1 = 0, nb, 0, n
            active_ptr, n_act = None, n
            if multistep:
                if d.get("nb") != nb:
                    d["bucket_active"] = torch.empty(nb, dtype=torch.uint8, device="cuda")
                    d["nb"] = nb
                counts = (C.c_int * 2)()
                L.cb200_active_sets_device(d["rung_all"].data_ptr(), tr.d_order, n, tr.d_bucketStarts,
                                           tr.d_bucketSizes, nb, self.active_rung, d["bucket_active"].data_ptr(),
                                           d["markers"].data_ptr(), counts, s)
                active_ptr, n_act = d["bucket_active"].data_ptr(), int(counts[1])
                self.active = {"buckets": int(counts[0]), "particles": n_act}
            if self.world > 1:  # my contiguous SFC range of buckets: equal particle counts, never splits a bucket
                starts = torch.empty(nb, dtype=torch.int32, device="cuda")
                L.cb200_copy_device(starts.data_ptr(), tr.d_bucketStarts, nb * 4, s)
                want = torch.tensor([self.rank * n // self.world, (self.rank + 1) * n // self.world],
                                    dtype=torch.int32, device="cuda")
                if multistep and n_act > 0:  # equal ACTIVE particle counts
                    at = [min(n_act - 1, self.rank * n_act // self.world),
                          min(n_act - 1, (self.rank + 1) * n_act // self.world)]
                    want = d["markers"][at]
                cut = torch.searchsorted(starts, want, right=False).tolist()
                b0 = 0 if self.rank == 0 else int(cut[0])
                b1 = nb if self.rank == self.world - 1 else int(cut[1])
                edge = starts[[min(b0, nb - 1), min(b1, nb - 1)]].tolist()
                p0 = int(edge[0]) if b0 < nb else n
                p1 = int(edge[1]) if b1 < nb else n
            self.range = (b0, b1, p0, p1)
            lists = hc.T.Lists()
            L.cb200_walk_device_active(nn, nb, tr.numLevels, lvl, tr.d_child0, tr.d_child1, tr.d_parent, tr.d_first,
                                       tr.d_last, tr.d_bucketFirst, tr.d_bucketCount, tr.d_bucketNode, tr.d_boxlo,
                                       tr.d_boxhi, mom64.data_ptr(), self.theta, self.nrep, self.period, b0, b1,
                                       active_ptr, C.byref(lists), s)
            if lists.error:
                raise RuntimeError(f"device walk: per-node capacity exceeded (error {lists.error})")
            mark("walk")
            L.cb200_pack_moments_device(mom32.data_ptr(), d["pk_mom"].data_ptr(), nn, s)
            vars_ = d["vars"]
            L.cb200_zero_vars_device(vars_.data_ptr(), n, s)
            P, V, M = tr.d_packedParts, vars_.data_ptr(), d["pk_mom"].data_ptr()
            fper = self.period if (self.nrep or self.ewald is not None) else 0.0
            mark("pack")
            if self.ewald is not None:
                from .tree import ewald_tables_fast as ewald_tables
                hc.stream_synchronize(s)
                root = mom64[0].cpu().numpy()
                momc, ewt = ewald_tables(root, self.period, self.ewald.get("dEwhCut", 2.8))
                if self._ew is None:
                    self._ew = hc.EwaldHostMemorySetup(1, len(ewt), 0)
                if p1 > p0:
                    hc.fill_ewald(self._ew, root, momc, ewt, self.period, float(self.ewald.get("dEwCut", 2.6)),
                                  self.nrep, active=None, first=p0, last=p1 - 1)
                    if multistep:  # large-phase form: the markers of my particle range, already on the device
                        mk = d["markers"][:n_act]
                        i0, i1 = torch.searchsorted(mk, torch.tensor([p0, p1], dtype=torch.int32, device="cuda")).tolist()
                        if i1 > i0:
                            L.cb200_ewald_device(P, V, mk.data_ptr() + 4 * i0, i1 - i0, self._ew.cachedData,
                                                 self._ew.ewt, s)
                    else:
                        L.cb200_EwaldHost(P, V, C.byref(self._ew), s, None, 0, 0)
            mark("ewald")
            mx = self.max_bucket
            L.cb200_cell_list_device_ex(P, V, M, lists.d_cell, lists.d_cellMarkers, lists.d_starts, lists.d_sizes,
                                        nb, fper, mx, s)
            L.cb200_part_list_device_ex(P, V, P, lists.d_part, lists.d_partMarkers, lists.d_starts, lists.d_sizes,
                                        nb, fper, mx, s)
            if lists.nSoft:
                L.cb200_part_list_device_ex(P, V, lists.d_nodeParticles, lists.d_soft, lists.d_softMarkers,
                                            lists.d_starts, lists.d_sizes, nb, fper, mx, s)
            mark("forces")
            # back to the caller's particle order: out[order[i]] = vars[i]
            order = torch.empty(n, dtype=torch.int32, device="cuda")
            L.cb200_copy_device(order.data_ptr(), tr.d_order, n * 4, s)
            if self.world == 1:
                d["out"].index_copy_(0, order.long(), vars_)
                self.out.copy_(d["out"], non_blocking=True)
            else:  # my rows only, with the caller indices they belong to
                if p1 - p0 > self.out.shape[0]:
                    raise RuntimeError("bucket range larger than the result buffer")
                self.out[: p1 - p0].copy_(vars_[p0:p1], non_blocking=True)
                self.out_idx[: p1 - p0].copy_(order[p0:p1], non_blocking=True)
            mark("d2h")
            self.info = {"nodes": nn, "buckets": nb, "levels": tr.numLevels, "nCell": int(lists.nCell),
                         "nSoft": int(lists.nSoft), "nPart": int(lists.nPart)}
            if keep_tree:
                self.kept_tree = self._download_tree(tr)
            if count_pairs:
                def dev_i32(ptr, count):
                    t = torch.empty(count, dtype=torch.int32, device="cuda")
                    L.cb200_copy_device(t.data_ptr(), ptr, count * 4, s)
                    return t.long()
                sizes = dev_i32(lists.d_sizes, nb)
                pairs = [int((torch.diff(dev_i32(m, nb + 1)) * sizes).sum().item())
                         for m in (lists.d_cellMarkers, lists.d_partMarkers, lists.d_softMarkers)]
                self.info.update(pc_pairs=pairs[0], pp_pairs=pairs[1] + pairs[2])
            L.cb200_lists_free(C.byref(lists), s)
            L.cb200_tree_free(C.byref(tr), s)
            hc.stream_synchronize(s)
        if phases is not None:
            for (_, a), (name, b) in zip(marks, marks[1:]):
                phases[name] = phases.get(name, 0.0) + a.elapsed_time(b)
        if self.world > 1:
            k = self.range[3] - self.range[2]
            self.d2h_bytes = k * 24  # 5 floats + the caller index per row
            return self.out_idx.numpy()[:k], self.out.numpy()[:k]
        return self.out.numpy()

    def _download_tree(self, tr):
        torch, L, s = self.torch, self.hc.L, self.stream

        def arr(ptr, count, dtype):
            t = torch.empty(count, dtype=dtype, device="cuda")
            if count:
                L.cb200_copy_device(t.data_ptr(), ptr, count * t.element_size(), s)
            self.hc.stream_synchronize(s)
            return t.cpu().numpy()
        n, nn, nb = tr.numParticles, tr.numNodes, tr.numBuckets
        i32, f64 = torch.int32, torch.float64
        out = {"order": arr(tr.d_order, n, i32), "pos": arr(tr.d_pos, 3 * n, f64).reshape(n, 3),
               "mass": arr(tr.d_mass, n, f64), "soft": arr(tr.d_soft, n, f64),
               "level_start": np.array(tr.levelStart[:tr.numLevels + 1], dtype=np.int32)}
        for name, ptr in (("child0", tr.d_child0), ("child1", tr.d_child1), ("parent", tr.d_parent),
                          ("first", tr.d_first), ("last", tr.d_last), ("bucket_first", tr.d_bucketFirst),
                          ("bucket_count", tr.d_bucketCount)):
            out[name] = arr(ptr, nn, i32)
        for name, ptr in (("geolo", tr.d_geolo), ("geohi", tr.d_geohi), ("boxlo", tr.d_boxlo), ("boxhi", tr.d_boxhi)):
            out[name] = arr(ptr, 3 * nn, f64).reshape(nn, 3)
        for name, ptr in (("bucket_node", tr.d_bucketNode), ("bucket_starts", tr.d_bucketStarts),
                          ("bucket_sizes", tr.d_bucketSizes)):
            out[name] = arr(ptr, nb, i32)
        return out

    def free(self):
        if self._ew is not None:
            self.hc.EwaldHostMemoryFree(self._ew, 0)
            self._ew = None
        self.h = self.out = self.dev = None
