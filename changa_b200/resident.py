"""The list kernels with every input already in HBM: the device-pointer entry points
(cb200_pack_*_device, cb200_zero_vars_device, cb200_ewald_device, cb200_cell_list_device_ex,
cb200_part_list_device_ex) driven on one stream -- what a host that builds or stages its lists on the
device calls, and what the kernel-only measurements and the CUDA-graph tests use.  Single GPU; torch
holds the device arrays."""
import numpy as np


class ResidentStep:
    def __init__(self, hc, wl, torch):
        self.hc, self.torch = hc, torch
        L = hc.L
        self.stream = hc.stream_create()
        self.ext = torch.cuda.ExternalStream(self.stream)
        rt = hc.np_real  # float32, or float64 for the CUDA_USE_DOUBLE build
        dev = torch.device("cuda", torch.cuda.current_device())
        self.n, self.nn = len(wl["parts"]), len(wl["moments"])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        with torch.cuda.stream(self.ext):
            self.raw_parts = up(np.ascontiguousarray(wl["parts"], dtype=rt))
            self.raw_mom = up(np.ascontiguousarray(wl["moments"], dtype=rt))
            pb, mb = L.cb200_packed_particle_bytes(), L.cb200_packed_moment_bytes()
            self.pk_parts = torch.empty(self.n * pb, dtype=torch.uint8, device=dev)
            self.pk_mom = torch.empty(self.nn * mb, dtype=torch.uint8, device=dev)
            self.vars = torch.zeros((self.n, 5), dtype=torch.float64 if rt == np.float64 else torch.float32, device=dev)
            self.lists = {}
            for key in ("cell", "part", "softcell"):
                if wl.get(key) and len(wl[key][0]):
                    il, m, st, sz = wl[key][:4]
                    self.lists[key] = (up(il), up(m), up(st), up(sz), len(st), int(sz.max()))
            self.soft_src = None
            if "softcell" in self.lists:
                raw = up(np.ascontiguousarray(wl["softcell"][4], dtype=rt))
                self.soft_src = torch.empty(raw.shape[0] * pb, dtype=torch.uint8, device=dev)
                L.cb200_pack_particles_device(raw.data_ptr(), self.soft_src.data_ptr(), raw.shape[0], self.stream)
            self.ew = None
            ew = wl.get("ewald")
            if ew:
                act = ew["active"] if ew["active"] is not None else np.arange(self.n, dtype=np.int32)
                self.ew_markers = up(np.ascontiguousarray(act, dtype=np.int32))
                e = hc.EwaldHostMemorySetup(len(act), len(ew["ewt"]), 1)
                hc.fill_ewald(e, ew["root"], ew["momc"], ew["ewt"], ew["L"], ew["fEwCut"], ew["nReps"], active=act)
                self.ew, self.ew_n = e, len(act)
            self.fperiod = float(wl.get("fperiod", 0.0))
        torch.cuda.synchronize()

    def step(self):
        L, s = self.hc.L, self.stream
        L.cb200_pack_moments_device(self.raw_mom.data_ptr(), self.pk_mom.data_ptr(), self.nn, s)
        L.cb200_pack_particles_device(self.raw_parts.data_ptr(), self.pk_parts.data_ptr(), self.n, s)
        L.cb200_zero_vars_device(self.vars.data_ptr(), self.n, s)
        P, V, M = self.pk_parts.data_ptr(), self.vars.data_ptr(), self.pk_mom.data_ptr()
        if self.ew is not None:
            L.cb200_ewald_device(P, V, self.ew_markers.data_ptr(), self.ew_n, self.ew.cachedData, self.ew.ewt, s)
        if "cell" in self.lists:
            il, m, st, sz, nb, mx = self.lists["cell"]
            L.cb200_cell_list_device_ex(P, V, M, il.data_ptr(), m.data_ptr(), st.data_ptr(), sz.data_ptr(), nb,
                                        self.fperiod, mx, s)
        if "part" in self.lists:
            il, m, st, sz, nb, mx = self.lists["part"]
            L.cb200_part_list_device_ex(P, V, P, il.data_ptr(), m.data_ptr(), st.data_ptr(), sz.data_ptr(), nb,
                                        self.fperiod, mx, s)
        if "softcell" in self.lists:
            il, m, st, sz, nb, mx = self.lists["softcell"]
            L.cb200_part_list_device_ex(P, V, self.soft_src.data_ptr(), il.data_ptr(), m.data_ptr(), st.data_ptr(),
                                        sz.data_ptr(), nb, self.fperiod, mx, s)

    def capture(self):
        """one step as a CUDA graph (kernels and one memset on one stream, all operands resident)"""
        torch = self.torch
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.ext):
            self.step()
        return self.graph
