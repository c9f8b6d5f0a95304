"""Synthetic inputs for the gravity hot path.

A workload is a dict in the shape DataManager/TreePiece hand to the GPU entry
points after serialisation:
  parts    (Np,5) float64  {mass, soft, x, y, z}        CompactPartData order
  moments  (Nn,27) float64 CudaMultipoleMoments order
  cell     (ilist (Li,2) int32 {index, offsetID}, markers (nb+1), starts (nb), sizes (nb))
  part     same, entries index particles (already expanded per source particle)
  fperiod  scalar period applied on all axes
  softcell None | (ilist, markers, starts, sizes, sources (Ns,5)): cells the reference's host
           evaluates as softened monopoles (Compute.cpp:1683-1699); here they travel as ad-hoc
           source particles {M, soft, cm} of a p-p request
  ewald    None | {root (27), momc (32), ewt (nh,5), L, fEwCut, nReps, active}
Arrays are float64 here; the C-ABI front end converts to cudatype on staging.
"""
import os

import numpy as np

from .hostcuda import encode_offset


def _ragged(rng, nb, lo, hi):
    return rng.integers(lo, hi + 1, nb)


def random_workload(seed=0, n_buckets=64, max_bucket=12, n_cells=400, cell_len=(0, 200), part_len=(0, 120),
                    periodic=True, min_bucket=1):
    """Random buckets, cells and lists with no tree behind them: exercises every
    code path of the list kernels (ragged and empty lists, self pairs, both
    spline branches, replica offsets, multi-pass buckets) at sizes a CPU check
    finishes instantly."""
    rng = np.random.default_rng(seed)
    sizes = _ragged(rng, n_buckets, min_bucket, max_bucket).astype(np.int32)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    n = int(sizes.sum())
    centres = rng.uniform(-0.45, 0.45, (n_buckets, 3))
    pos = np.repeat(centres, sizes, axis=0) + rng.normal(0, 0.01, (n, 3))
    parts = np.column_stack([rng.uniform(0.5, 1.5, n) / n, rng.uniform(0.004, 0.02, n), pos])

    mom = np.zeros((n_cells, 27))
    mom[:, 0] = rng.uniform(0.01, 0.03, n_cells)            # radius
    mom[:, 1] = rng.uniform(0.004, 0.02, n_cells)           # soft
    mom[:, 2] = rng.uniform(0.5, 2.0, n_cells) * 8 / n      # mass
    mom[:, 3:6] = rng.uniform(-0.5, 0.5, (n_cells, 3))      # cm
    mom[:, 6:] = rng.normal(0, 0.15, (n_cells, 21)) * mom[:, 2:3]

    fperiod = 1.0 if periodic else 0.0
    reps = (-1, 0, 1) if periodic else (0,)

    cl, cm_ = [], [0]
    for b in range(n_buckets):
        want = int(rng.integers(cell_len[0], cell_len[1] + 1))
        idx = rng.integers(0, n_cells, want)
        off = rng.choice(reps, (want, 3))
        d = mom[idx, 3:6] + off * fperiod - centres[b]
        keep = np.sqrt((d ** 2).sum(1)) > 0.12                # well-separated: expansion converges
        idx, off = idx[keep], off[keep]
        codes = np.array([encode_offset(int(o[0]), int(o[1]), int(o[2]), b & 0x3fffff) for o in off], dtype=np.int64)
        cl.append(np.column_stack([idx, codes]) if len(idx) else np.zeros((0, 2), dtype=np.int64))
        cm_.append(cm_[-1] + len(idx))
    cell_list = np.concatenate(cl).astype(np.int64)
    cell_list = (cell_list & 0xffffffff).astype(np.uint32).view(np.int32).reshape(-1, 2)

    pl, pm = [], [0]
    for b in range(n_buckets):
        want = int(rng.integers(part_len[0], part_len[1] + 1))
        own = np.arange(starts[b], starts[b] + sizes[b])     # includes self pairs (r = 0 -> skipped)
        others = rng.integers(0, n, max(want - len(own), 0))
        idx = np.concatenate([own, others])[:max(want, 0)] if want else np.zeros(0, dtype=np.int64)
        off = rng.choice(reps, (len(idx), 3))
        off[: min(len(own), len(idx))] = 0
        codes = np.array([encode_offset(int(o[0]), int(o[1]), int(o[2])) for o in off], dtype=np.int64)
        pl.append(np.column_stack([idx, codes]) if len(idx) else np.zeros((0, 2), dtype=np.int64))
        pm.append(pm[-1] + len(idx))
    part_list = np.concatenate(pl).astype(np.int64)
    part_list = (part_list & 0xffffffff).astype(np.uint32).view(np.int32).reshape(-1, 2)

    def filled(ilist, markers):
        """serialize() drops buckets with empty lists (Compute.cpp:1075-1100)"""
        markers = np.asarray(markers)
        ln = np.diff(markers)
        k = ln > 0
        m = np.concatenate([[0], np.cumsum(ln[k])]).astype(np.int32)
        return (np.ascontiguousarray(ilist, dtype=np.int32), m, starts[k].copy(), sizes[k].copy())

    return {
        "parts": parts, "moments": mom, "fperiod": fperiod,
        "cell": filled(cell_list, cm_), "part": filled(part_list, pm),
        "ewald": None, "name": f"random(seed={seed},nb={n_buckets},maxb={max_bucket})",
    }


def interaction_counts(wl):
    """pair interactions the way ChaNGa counts them (Compute.cpp:1643-1651):
    list entries x target particles of the bucket."""
    out = {}
    for key in ("cell", "part", "softcell"):
        if wl.get(key):
            _, m, _, sz = wl[key][:4]
            out[key] = int((np.diff(m).astype(np.int64) * sz.astype(np.int64)).sum())
        else:
            out[key] = 0
    return out


# ---------------------------------------------------------------------------
# particle sets of BASELINE.json's configs and the tree workloads built on them
# ---------------------------------------------------------------------------
def uniform_box(n, seed=1, xmin=-0.5, xmax=0.5):
    """Poisson box exactly as the reference's testdata/ppartt.c:34-65 makes it: glibc
    srand(seed); x,y,z = xmin + rand()/RAND_MAX*(xmax-xmin) per particle in that order;
    m = 1/N; eps = N^(-1/3) (xmax-xmin)/20.  Returns (pos, mass, soft)."""
    import ctypes
    from .tree import lib
    pos = np.zeros((n, 3))
    f = lib().cb200h_uniform_box
    f.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    f(int(seed), int(n), float(xmin), float(xmax), pos.ctypes.data)
    mass = np.full(n, 1.0 / n)
    soft = np.full(n, n ** (-1.0 / 3.0) * (xmax - xmin) / 20.0)
    return pos, mass, soft


def clustered_box(n, seed=2, n_halos=None, frac_halo=0.7, rs_range=(2e-4, 2e-2)):
    """SURVEY 8d config C4 recipe: (1-frac_halo) uniform background + frac_halo in Plummer
    spheres (centres uniform, scale radii log-uniform in rs_range, halo mass ~ r_s), wrapped
    into [-0.5,0.5); m = 1/N, eps = N^(-1/3)/20."""
    rng = np.random.default_rng(seed)
    n_halos = n_halos or max(8, n // 16384)
    nh = int(frac_halo * n)
    nb = n - nh
    bg = rng.uniform(-0.5, 0.5, (nb, 3))
    cen = rng.uniform(-0.5, 0.5, (n_halos, 3))
    rs = np.exp(rng.uniform(np.log(rs_range[0]), np.log(rs_range[1]), n_halos))
    which = rng.choice(n_halos, nh, p=rs / rs.sum())
    # Plummer radius from the cumulative mass profile, truncated at 10 r_s
    u = rng.uniform(0, 0.985, nh)
    r = rs[which] / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    d = rng.normal(size=(nh, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    halo = cen[which] + d * r[:, None]
    pos = np.concatenate([bg, halo])
    pos = (pos + 0.5) % 1.0 - 0.5
    pos = pos[rng.permutation(n)]
    return pos, np.full(n, 1.0 / n), np.full(n, n ** (-1.0 / 3.0) / 20.0)


def clustered_box_rows(n, lo, hi, seed=2, n_halos=None, frac_halo=0.7, rs_range=(2e-4, 2e-2), block=1 << 20):
    """Rows [lo, hi) of the C4-recipe box of n particles, generated block by block with one random
    stream per block of 2^20 particles, so that a rank of a multi-GPU run makes only its own rows (the
    single-stream clustered_box above needs the whole box: 90 s and 10 GB at 512^3).  The same recipe:
    every particle is background (uniform) with probability 1 - frac_halo, otherwise a member of one of
    the Plummer spheres (halo table: one stream shared by all blocks).  Rows do not depend on lo / hi."""
    n_halos = n_halos or max(8, n // 16384)
    hr = np.random.default_rng([seed, 0x68616c6f])
    cen = hr.uniform(-0.5, 0.5, (n_halos, 3))
    rs = np.exp(hr.uniform(np.log(rs_range[0]), np.log(rs_range[1]), n_halos))
    cdf = np.cumsum(rs / rs.sum())
    out = np.empty((hi - lo, 3))
    for b in range(lo // block, (hi + block - 1) // block):
        b0, b1 = b * block, min(n, (b + 1) * block)
        rng = np.random.default_rng([seed, b])
        m = b1 - b0
        pos = rng.uniform(-0.5, 0.5, (m, 3))
        in_halo = rng.random(m) < frac_halo
        k = int(in_halo.sum())
        which = np.minimum(np.searchsorted(cdf, rng.random(k)), n_halos - 1)
        u = rng.uniform(0, 0.985, k)
        r = rs[which] / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        d = rng.normal(size=(k, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pos[in_halo] = cen[which] + d * r[:, None]
        pos = (pos + 0.5) % 1.0 - 0.5
        a, e = max(lo, b0), min(hi, b1)
        out[a - lo:e - lo] = pos[a - b0:e - b0]
    return out, 1.0 / n, n ** (-1.0 / 3.0) / 20.0


def density_rungs(pos, period=1.0, max_rung=6):
    """particle rung = clamp(floor(log2(rho_local / rho_mean) / 2), 0, 6), so that
    force steps at activeRung 2, 4 exercise the multistep path (active-bucket subset + Ewald
    markers).  rho_local: counts on a grid with ~8 particles per cell on average."""
    pos = np.asarray(pos)
    n = len(pos)
    g = max(1, int(round((n / 8.0) ** (1.0 / 3.0))))
    ijk = np.floor((pos / period + 0.5) * g).astype(np.int64) % g
    cell = (ijk[:, 0] * g + ijk[:, 1]) * g + ijk[:, 2]
    cnt = np.bincount(cell, minlength=g ** 3)
    ratio = cnt[cell] * (float(g) ** 3 / n)
    return np.clip(np.floor(np.log2(ratio) / 2.0), 0, max_rung).astype(np.uint8)


def cosmo_box(n_side=48, seed=300, growth=0.035):
    """Stand-in for testcosmo/cube300.tbin (48^3 = 110592 dark particles, periodic unit box,
    the fixture itself stays in the reference tree): a grid displaced by a Gaussian random
    field with a P(k) ~ k^-2 spectrum (Zel'dovich), which gives the filament/void clustering
    of an evolved box.  m = 1/N, eps = N^(-1/3)/20."""
    rng = np.random.default_rng(seed)
    n = n_side
    k = np.fft.fftfreq(n) * n
    kx, ky, kz = np.meshgrid(k, k, np.fft.rfftfreq(n) * n, indexing="ij")
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    k2[0, 0, 0] = 1.0
    amp = k2 ** (-1.0) * np.exp(-k2 / (0.35 * n) ** 2)
    amp[0, 0, 0] = 0.0
    delta = (rng.normal(size=k2.shape) + 1j * rng.normal(size=k2.shape)) * amp
    disp = [np.fft.irfftn(1j * kk / k2 * delta, s=(n, n, n), axes=(0, 1, 2)) for kk in (kx, ky, kz)]
    disp = np.stack(disp, axis=-1)
    disp *= growth / disp.std()
    g = (np.arange(n) + 0.5) / n - 0.5
    q = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1)
    pos = ((q + disp + 0.5) % 1.0 - 0.5).reshape(-1, 3)
    np_ = n ** 3
    return pos, np.full(np_, 1.0 / np_), np.full(np_, np_ ** (-1.0 / 3.0) / 20.0)


def read_tipsy(path):
    """Standard (big-endian XDR) Tipsy snapshot -> (pos, mass, soft) of all particles
    (testdata/tipsydefs.h:4-110: gas 12 floats, dark 9, star 11; header 28 bytes + pad)."""
    import struct
    with open(path, "rb") as f:
        raw = f.read()
    time, nbodies, ndim, nsph, ndark, nstar = struct.unpack(">diiiii", raw[:28])
    off = 32 if len(raw) >= 32 + 4 * (12 * nsph + 9 * ndark + 11 * nstar) else 28
    out = []
    for cnt, width, soft_col in ((nsph, 12, 9), (ndark, 9, 7), (nstar, 11, 9)):
        a = np.frombuffer(raw, dtype=">f4", count=cnt * width, offset=off).reshape(cnt, width).astype(np.float64)
        off += 4 * cnt * width
        out.append((a[:, 1:4], a[:, 0], a[:, soft_col]))
    pos = np.concatenate([o[0] for o in out])
    return pos, np.concatenate([o[1] for o in out]), np.concatenate([o[2] for o in out])


CONFIGS = {
    # name: (generator, kwargs, theta, nReplicas, periodic/Ewald)
    "cube300": dict(gen="cosmo", n_side=48, theta=0.7, n_replicas=1, ewald=True),
    "king": dict(gen="plummer", n=36000, theta=0.7, n_replicas=0, ewald=False),
    "uniform": dict(gen="uniform", n=1 << 20, theta=0.7, n_replicas=1, ewald=True),
    "clustered": dict(gen="clustered", n=1 << 20, theta=0.7, n_replicas=1, ewald=True),
    # config 5 (testcollapse/adiabtophat_glass_28721.bin: a uniform sphere at rest, gravity only,
    # theta = 0.55, isolated; the reference runs it in double)
    "collapse": dict(gen="tophat", n=28721, theta=0.55, n_replicas=0, ewald=False),
}


def tophat_sphere(n, seed=5, radius=0.5):
    """stand-in for testcollapse/adiabtophat_glass_28721.bin: a homogeneous sphere of unit mass
    (the fixture is a glass; a Poisson sphere has the same tree depth and list lengths)"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = radius * rng.uniform(0, 1, n) ** (1.0 / 3.0)
    return d * r[:, None], np.full(n, 1.0 / n), np.full(n, radius * n ** (-1.0 / 3.0) / 2.0)


def plummer_sphere(n, seed=7, rs=2.0, soft=0.2):
    """stand-in for teststep/king_soft.bin (36000 equal-mass particles, isolated cluster)"""
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 0.99, n)
    r = rs / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d * r[:, None], np.full(n, 1.0 / n), np.full(n, soft)


_FIXTURES = None


def fixture_particles(name):
    """(pos, mass, soft) of the reference's Tipsy fixture of a config (testcosmo/cube300.tbin,
    teststep/king_soft.bin, testcollapse/adiabtophat_glass_28721.bin) from the committed
    tests/golden/fixture_positions.npz, or None.  The periodic box is wrapped into [-0.5, 0.5)."""
    global _FIXTURES
    if _FIXTURES is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "fixture_positions.npz")
        _FIXTURES = dict(np.load(path)) if os.path.exists(path) else {}
    if name + "_pos" not in _FIXTURES:
        return None
    pos = _FIXTURES[name + "_pos"].astype(np.float64)
    if name == "cube300":
        pos = (pos + 0.5) % 1.0 - 0.5
    n = len(pos)
    return pos, np.full(n, float(_FIXTURES[name + "_mass"])), np.full(n, float(_FIXTURES[name + "_soft"]))


def config_workload(name="cube300", n=None, seed=None, max_bucket=12, bucket_range_of=None, gen_kwargs=None,
                    **over):
    """Tree workload for one of BASELINE.json's configs.  bucket_range_of=(rank, world):
    lists only for that rank's contiguous SFC share of the buckets (equal particle counts);
    the whole tree (particles + moments) stays in the workload, as after the all-gather."""
    from .tree import Tree, tree_workload
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    fixture = fixture_particles(name) if (n is None and seed is None and not gen_kwargs) else None
    label = name
    if fixture is not None:
        # the reference's own particle set for this config (tests/golden/fixture_positions.npz)
        pos, mass, soft = fixture
        label = {"cube300": "cube300.tbin", "king": "king_soft.bin", "collapse": "adiabtophat_glass_28721.bin"}[name]
    elif cfg["gen"] == "cosmo":
        side = cfg["n_side"] if n is None else int(round(n ** (1.0 / 3.0)))
        pos, mass, soft = cosmo_box(side, seed=seed or 300)
    elif cfg["gen"] == "uniform":
        pos, mass, soft = uniform_box(n or cfg["n"], seed=seed or 1)
    elif cfg["gen"] == "clustered":
        pos, mass, soft = clustered_box(n or cfg["n"], seed=seed or 2)
    elif cfg["gen"] == "tophat":
        pos, mass, soft = tophat_sphere(n or cfg["n"], seed=seed or 5)
    else:
        pos, mass, soft = plummer_sphere(n or cfg["n"], seed=seed or 7, **(gen_kwargs or {}))
    root_lo, root_hi = (-0.5,) * 3, (0.5,) * 3
    if not cfg["ewald"]:
        ext = float(np.abs(pos).max()) * 1.0001
        root_lo, root_hi = (-ext,) * 3, (ext,) * 3
    t = Tree(pos, mass, soft, max_bucket=max_bucket, root_lo=root_lo, root_hi=root_hi)
    rng_b = None
    if bucket_range_of is not None:
        rank, world = bucket_range_of
        from .multigpu import bucket_cuts_by_particles
        cuts = bucket_cuts_by_particles(t.bucket_sizes, world)
        rng_b = (int(cuts[rank]), int(cuts[rank + 1]))
    wl = tree_workload(None, None, None, theta=cfg["theta"], n_replicas=cfg["n_replicas"], period=1.0,
                       ewald={} if cfg["ewald"] else None, bucket_range=rng_b, tree=t,
                       name=f"{label}(N={t.n},theta={cfg['theta']},nReplicas={cfg['n_replicas']},bucket={max_bucket})")
    wl["bucket_range"] = rng_b or (0, t.num_buckets)
    return wl
