"""ctypes loader for tests/emu/libpcpx_emu.so — the host build of the device traversal code.

TEST HARNESS ONLY: lets the CPU-only suite check the search logic against the oracle. It is not
a fallback of the product; libpcpx.so contains none of it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libpcpx_emu.so")
CSRC = os.path.join(ROOT, "point-cloud-processing_b200", "csrc")

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def build_emu(force=False):
    srcs = [os.path.join(EMU_DIR, "emu.cu")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))
    ]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.run(
            ["nvcc", "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-diag-suppress", "20013",
             "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
             "-I", os.path.join(ROOT, "include"), "-I", CSRC,
             os.path.join(EMU_DIR, "emu.cu"), "-o", EMU_SO],
            check=True, capture_output=True)


class Emu:
    def __init__(self):
        build_emu()
        L = C.CDLL(EMU_SO)
        L.emu_index_create.restype = C.c_void_p
        L.emu_index_create.argtypes = [_f32p, C.c_size_t, C.c_int, _f32p, C.c_uint32, C.c_uint32]
        L.emu_index_destroy.argtypes = [C.c_void_p]
        L.emu_index_info.argtypes = [C.c_void_p, _u64p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     _u64p]
        L.emu_plan.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.emu_knn.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_uint32, C.c_double, C.c_double,
                              C.c_int, _u32p, _f32p, _u32p, _u64p, _u32p]
        L.emu_normals.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_uint32, C.c_double,
                                  C.c_double, C.c_int, _f32p, _f32p, _f32p, _u32p]
        L.emu_tile.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_uint32, C.c_double, C.c_int, _u32p, _f32p, _u32p, _f32p, _f32p,
                               _f32p, _u8p, _u64p]
        L.emu_radius.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, C.c_float, _u32p, _u64p,
                                 _u32p]
        L.emu_density_keep.argtypes = [C.c_void_p, C.c_float, C.c_uint32, _u8p]
        L.emu_smallest_eigenvector.argtypes = [_f32p, _f32p, _f32p]
        L.emu_sorted_order.argtypes = [C.c_void_p, _u32p]
        L.emu_bilateral_points_step.argtypes = [C.c_void_p, _f32p, C.c_float, C.c_float, _f32p]
        L.emu_bilateral_normals_step.argtypes = [C.c_void_p, _f32p, C.c_float, C.c_float, _f32p]
        L.emu_wlop_density.argtypes = [C.c_void_p, C.c_float, C.c_float, _f32p]
        L.emu_wlop_step.argtypes = [C.c_void_p, _f32p, C.c_void_p, _f32p, C.c_float, C.c_float,
                                    _f32p]
        self.L = L

    def index(self, xyz, bbox=None, max_level=0, min_occ=0):
        return EmuIndex(self, xyz, bbox, max_level, min_occ)

    # ---- the host loops of smoothing.cu, restated over the emulated index ------------------
    def bilateral_filter_points(self, xyz, normals, sigmaf, sigmag, iterations):
        cur = _f32(xyz).reshape(-1, 3).copy()
        nrm = _f32(normals).reshape(-1, 3)
        for _ in range(iterations):
            ix = self.index(cur)
            nxt = np.zeros_like(cur)
            self.L.emu_bilateral_points_step(ix.h, _ptr(nrm, _f32p), sigmaf, sigmag,
                                             _ptr(nxt, _f32p))
            cur = nxt
        return cur

    def bilateral_filter_normals(self, xyz, normals, sigmaf, sigmag, iterations):
        ix = self.index(_f32(xyz).reshape(-1, 3))
        cur = _f32(normals).reshape(-1, 3).copy()
        for _ in range(iterations):
            nxt = np.zeros_like(cur)
            self.L.emu_bilateral_normals_step(ix.h, _ptr(cur, _f32p), sigmaf, sigmag,
                                              _ptr(nxt, _f32p))
            cur = nxt
        return cur

    def wlop(self, xyz, initial, mu, h, iterations, uniform=True):
        xyz = _f32(xyz).reshape(-1, 3)
        x = xyz[np.asarray(initial, dtype=np.int64)].copy()
        ixp = self.index(xyz)
        vj = wi = None
        if uniform:
            vj = np.zeros(len(xyz), np.float32)
            self.L.emu_wlop_density(ixp.h, h, mu, _ptr(vj, _f32p))
        for _ in range(iterations):
            ixq = self.index(x)
            if uniform:
                wi = np.zeros(len(x), np.float32)
                self.L.emu_wlop_density(ixq.h, h, mu, _ptr(wi, _f32p))
            nxt = np.zeros_like(x)
            self.L.emu_wlop_step(ixp.h, _ptr(vj, _f32p), ixq.h, _ptr(wi, _f32p), h, mu,
                                 _ptr(nxt, _f32p))
            x = nxt
        return x

    def smallest_eigenvector(self, cov6):
        c = _f32(cov6)
        n = np.zeros(3, np.float32)
        g = np.zeros(1, np.float32)
        self.L.emu_smallest_eigenvector(_ptr(c, _f32p), _ptr(n, _f32p), _ptr(g, _f32p))
        return n, float(g[0])


class EmuIndex:
    def __init__(self, emu, xyz, bbox, max_level, min_occ):
        self.L = emu.L
        self.xyz = _f32(xyz).reshape(-1, 3)
        self.n = len(self.xyz)
        bb = _f32(bbox)
        self.h = self.L.emu_index_create(_ptr(self.xyz, _f32p), self.n, 0 if bb is None else 1,
                                         _ptr(bb, _f32p), max_level, min_occ)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_index_destroy(self.h)
            self.h = None

    def info(self):
        n = C.c_uint64()
        lcap, lfine = C.c_int(), C.c_int()
        slots = C.c_uint64()
        self.L.emu_index_info(self.h, C.byref(n), C.byref(lcap), C.byref(lfine), C.byref(slots))
        return dict(n_indexed=n.value, lcap=lcap.value, lfine=lfine.value, slots=slots.value)

    def sorted_order(self):
        out = np.zeros(self.n, np.uint32)
        self.L.emu_sorted_order(C.c_void_p(self.h), _ptr(out, _u32p))
        return out

    def plan(self, k, margin=1.15):
        lv, rg, occ = C.c_int(), C.c_int(), C.c_double()
        self.L.emu_plan(self.h, k, margin, C.byref(lv), C.byref(rg), C.byref(occ))
        return dict(level=lv.value, rings=rg.value, occupancy=occ.value)

    def knn(self, queries, k, eps=1e-5, level_factor=1.15, exact_only=False):
        """exact_only=False mirrors the product kernel (two-pass + exact fallback); True forces
        the 64-bit (distance, index) search for every query.  st = candidates, lookups,
        attempts, fallbacks."""
        q = _f32(queries)
        nq = self.n if q is None else len(q)
        idx = np.full((nq, k), 0xFFFFFFFF, np.uint32)
        d2 = np.full((nq, k), np.inf, np.float32)
        cnt = np.zeros(nq, np.uint32)
        st = np.zeros(4, np.uint64)
        self.per_query_candidates = np.zeros(nq, np.uint32)
        rc = self.L.emu_knn(self.h, _ptr(q, _f32p), nq, k, eps, level_factor,
                            1 if exact_only else 0, _ptr(idx, _u32p),
                            _ptr(d2, _f32p), _ptr(cnt, _u32p), _ptr(st, _u64p),
                            _ptr(self.per_query_candidates, _u32p))
        assert rc == 0
        return idx, d2, cnt, st

    def normals(self, queries, k, eps=1e-5, level_factor=1.15, want_means=False,
                exact_only=False):
        q = _f32(queries)
        nq = self.n if q is None else len(q)
        ctr = np.zeros((nq, 3), np.float32)
        nrm = np.zeros((nq, 3), np.float32)
        means = np.zeros(nq, np.float32) if want_means else None
        ties = np.zeros(1, np.uint32)
        rc = self.L.emu_normals(self.h, _ptr(q, _f32p), nq, k, eps, level_factor,
                                1 if exact_only else 0, _ptr(ctr, _f32p), _ptr(nrm, _f32p), _ptr(means, _f32p),
                                _ptr(ties, _u32p))
        assert rc == 0
        return nrm, ctr, means, int(ties[0])

    def tile(self, k, mode, level, sub=2, eps=1e-5, max_points=1024, scan_cap=1.0, nthreads=128,
             alg=2):
        """The tile path (tile_core.cuh) over the indexed points themselves; mode 0 = kNN rows,
        1 = mean distance, 2 = normals.  Returns a dict with the outputs, `done` (rows the tile
        pass finished; the product sends the rest to the retry queue) and `stats`."""
        n = self.n
        idx = np.full((n, k), 0xFFFFFFFF, np.uint32)
        d2 = np.full((n, k), np.inf, np.float32)
        cnt = np.zeros(n, np.uint32)
        nrm = np.zeros((n, 3), np.float32)
        ctr = np.zeros((n, 3), np.float32)
        means = np.zeros(n, np.float32)
        done = np.zeros(n, np.uint8)
        st = np.zeros(6, np.uint64)
        rc = self.L.emu_tile(self.h, k, eps, mode, sub, alg, level, max_points, scan_cap, nthreads,
                             _ptr(idx, _u32p), _ptr(d2, _f32p), _ptr(cnt, _u32p),
                             _ptr(nrm, _f32p), _ptr(ctr, _f32p), _ptr(means, _f32p),
                             _ptr(done, _u8p), _ptr(st, _u64p))
        assert rc == 0, rc
        return dict(idx=idx, d2=d2, cnt=cnt, normals=nrm, centroids=ctr, means=means,
                    done=done.astype(bool),
                    stats=dict(tiles=int(st[0]), fallback_tiles=int(st[1]), not_final=int(st[2]),
                               candidates=int(st[3]), max_region=int(st[4]),
                               order_ambiguous=int(st[5])))

    def radius_count(self, queries, r, radii=None):
        q = _f32(queries)
        radii = _f32(radii)
        nq = self.n if q is None else len(q)
        cnt = np.zeros(nq, np.uint32)
        self.L.emu_radius(self.h, _ptr(q, _f32p), nq, _ptr(radii, _f32p), np.float32(r),
                          _ptr(cnt, _u32p), None, None)
        return cnt

    def radius_search(self, queries, r, radii=None):
        q = _f32(queries)
        radii = _f32(radii)
        nq = self.n if q is None else len(q)
        cnt = self.radius_count(queries, r, radii)
        off = np.zeros(nq + 1, np.uint64)
        np.cumsum(cnt, out=off[1:])
        idx = np.zeros(int(off[-1]), np.uint32)
        self.L.emu_radius(self.h, _ptr(q, _f32p), nq, _ptr(radii, _f32p), np.float32(r), None,
                          _ptr(off, _u64p), _ptr(idx, _u32p))
        return off, idx

    def density_keep(self, r, threshold):
        keep = np.zeros(self.n, np.uint8)
        self.L.emu_density_keep(self.h, np.float32(r), threshold, _ptr(keep, _u8p))
        return keep
