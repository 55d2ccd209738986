"""ctypes binding of oracle/libpumipic_oracle.so (test infrastructure; see oracle/pumipic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_fp = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_ubyte)


class SearchStats(C.Structure):
    _fields_ = [("loops", C.c_int), ("not_in_elem", C.c_int), ("not_found", C.c_int),
                ("aborted", C.c_int)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libpumipic_oracle.so")
    src = os.path.join(ORACLE_DIR, "pumipic_oracle.c")
    hdr = os.path.join(ORACLE_DIR, "pumipic_oracle.h")
    if (not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
            or os.path.getmtime(so) < os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libpumipic_oracle.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_oracle())
        _LIB.orc_mesh_create.restype = C.c_void_p
        _LIB.orc_compute_tolerance.restype = C.c_double
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _u8(a):
    return a.ctypes.data_as(c_u8p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleMesh:
    """Owns an orc_mesh built from a tests.meshes.Mesh."""

    def __init__(self, mesh):
        self.mesh = mesh
        L = lib()
        self._keep = (mesh.coords, mesh.elem2verts, mesh.elem2sides, mesh.side2verts)
        self.h = C.c_void_p(L.orc_mesh_create(
            C.c_int(mesh.dim), C.c_int(mesh.nverts), _dp(mesh.coords), C.c_int(mesh.nelems),
            _ip(mesh.elem2verts), C.c_int(mesh.nsides), _ip(mesh.elem2sides),
            _ip(mesh.side2verts)))

    def __del__(self):
        try:
            lib().orc_mesh_destroy(self.h)
        except Exception:
            pass

    @property
    def tol(self):
        return lib().orc_compute_tolerance(self.h)

    def _derived(self, idx_ptr, n, dtype):
        # struct layout: 4 ints, 4 pointers, then derived pointers in declaration order
        base = C.cast(self.h, C.POINTER(C.c_void_p))
        addr = base[2 + 4 + idx_ptr]   # 4 ints = 2 pointer slots
        ct = {np.int32: C.c_int, np.float64: C.c_double, np.int8: C.c_byte}[dtype]
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ct)), shape=(n,)).copy()

    def side2elem_off(self):
        return self._derived(0, self.mesh.nsides + 1, np.int32)

    def side2elem(self):
        off = self.side2elem_off()
        return self._derived(1, int(off[-1]), np.int32)

    def dual_off(self):
        return self._derived(2, self.mesh.nelems + 1, np.int32)

    def dual(self):
        off = self.dual_off()
        return self._derived(3, int(off[-1]), np.int32)

    def exposed(self):
        return self._derived(4, self.mesh.nsides, np.int8)

    def vol(self):
        return self._derived(5, self.mesh.nelems, np.float64)

    def vert2elem_off(self):
        return self._derived(6, self.mesh.nverts + 1, np.int32)

    def vert2elem(self):
        off = self.vert2elem_off()
        return self._derived(7, int(off[-1]), np.int32)

    # ---- searches: all particle arrays are [3, stride] float64 (component-major) ----
    def search_mesh(self, slot_elem, mask, x, xtgt, elem_ids=None, require_intersection=False,
                    looplimit=0):
        cap = mask.shape[0]
        empty = elem_ids is None
        ids = np.full(cap, -1, np.int32) if empty else np.ascontiguousarray(elem_ids, np.int32).copy()
        dim = self.mesh.dim
        faces = np.full(cap, -1, np.int32)
        pts = np.zeros(dim * cap, np.float64)
        st = SearchStats()
        x = _f64(x); xtgt = _f64(xtgt)
        found = lib().orc_search_mesh(
            self.h, C.c_int(cap), _ip(np.ascontiguousarray(slot_elem, np.int32)),
            _u8(np.ascontiguousarray(mask, np.uint8)), _dp(x), _dp(xtgt),
            C.c_long(x.shape[1]), _ip(ids), C.c_int(empty), C.c_int(bool(require_intersection)),
            _ip(faces), _dp(pts), C.c_int(1), C.c_int(looplimit), C.byref(st))
        return bool(found), ids, faces, pts.reshape(cap, dim), st

    def search_mesh_2d(self, slot_elem, mask, xtgt, elem_ids, looplimit=0):
        cap = mask.shape[0]
        ids = np.ascontiguousarray(elem_ids, np.int32).copy()
        st = SearchStats()
        xtgt = _f64(xtgt)
        found = lib().orc_search_mesh_2d(
            self.h, C.c_int(cap), _ip(np.ascontiguousarray(slot_elem, np.int32)),
            _u8(np.ascontiguousarray(mask, np.uint8)), _dp(xtgt), C.c_long(xtgt.shape[1]),
            _ip(ids), C.c_int(looplimit), C.byref(st))
        return bool(found), ids, st

    def search_mesh_legacy3d(self, slot_elem, mask, x, xtgt, elem_ids=None, looplimit=0):
        cap = mask.shape[0]
        empty = elem_ids is None
        ids = np.full(cap, -1, np.int32) if empty else np.ascontiguousarray(elem_ids, np.int32).copy()
        xpts = np.zeros(3 * cap, np.float64)
        xface = np.full(cap, -1, np.int32)
        st = SearchStats()
        x = _f64(x); xtgt = _f64(xtgt)
        found = lib().orc_search_mesh_legacy3d(
            self.h, C.c_int(cap), _ip(np.ascontiguousarray(slot_elem, np.int32)),
            _u8(np.ascontiguousarray(mask, np.uint8)), _dp(x), _dp(xtgt), C.c_long(x.shape[1]),
            _ip(ids), C.c_int(empty), _dp(xpts), _ip(xface), C.c_int(looplimit), C.byref(st))
        return bool(found), ids, xpts.reshape(cap, 3), xface, st

    def search_mesh_3d(self, slot_elem, mask, x, xtgt, elem_ids=None, looplimit=0):
        cap = mask.shape[0]
        empty = elem_ids is None
        ids = np.full(cap, -1, np.int32) if empty else np.ascontiguousarray(elem_ids, np.int32).copy()
        xpts = np.zeros(3 * cap, np.float64)
        xface = np.full(cap, -1, np.int32)
        st = SearchStats()
        x = _f64(x); xtgt = _f64(xtgt)
        found = lib().orc_search_mesh_3d(
            self.h, C.c_int(cap), _ip(np.ascontiguousarray(slot_elem, np.int32)),
            _u8(np.ascontiguousarray(mask, np.uint8)), _dp(x), _dp(xtgt), C.c_long(x.shape[1]),
            _ip(ids), C.c_int(empty), _dp(xpts), _ip(xface), C.c_int(looplimit), C.byref(st))
        return bool(found), ids, xpts.reshape(cap, 3), xface, st

    def gyro_scatter(self, slot_elem, mask, v2v, rmax, nrings, ppr):
        out = np.zeros(self.mesh.nverts, np.float64)
        lib().orc_gyro_scatter(self.h, C.c_int(mask.shape[0]),
                               _ip(np.ascontiguousarray(slot_elem, np.int32)),
                               _u8(np.ascontiguousarray(mask, np.uint8)),
                               _ip(np.ascontiguousarray(v2v, np.int32)), C.c_double(rmax),
                               C.c_int(nrings), C.c_int(ppr), _dp(out))
        return out

    def gyro_ring_map(self, rmax, nrings, ppr, theta_deg):
        n = 3 * self.mesh.nverts * nrings * ppr
        out = np.empty(n, np.int32)
        found = lib().orc_gyro_ring_map(self.h, C.c_double(rmax), C.c_int(nrings), C.c_int(ppr),
                                        C.c_double(theta_deg), _ip(out))
        return bool(found), out


# ---- geometry primitives ----
def barycentric_tet(vol, M, p):
    bcc = np.zeros(4)
    ok = lib().orc_barycentric_tet(C.c_double(vol), _dp(_f64(M).ravel()), _dp(_f64(p)), _dp(bcc))
    return ok, bcc


def find_barycentric_tet(M, p):
    bcc = np.zeros(4)
    ok = lib().orc_find_barycentric_tet(_dp(_f64(M).ravel()), _dp(_f64(p)), _dp(bcc))
    return ok, bcc


def barycentric_tri(area, M, p):
    bcc = np.zeros(3)
    lib().orc_barycentric_tri(C.c_double(area), _dp(_f64(M).ravel()), _dp(_f64(p)), _dp(bcc))
    return bcc


def ray_intersects_triangle(face, orig, dest, tol, flip, segment=False):
    xp = np.zeros(3)
    dproj, close, par = C.c_double(), C.c_double(), C.c_double()
    fn = lib().orc_line_segment_intersects_triangle if segment else lib().orc_ray_intersects_triangle
    hit = fn(_dp(_f64(face).ravel()), _dp(_f64(orig)), _dp(_f64(dest)), _dp(xp), C.c_double(tol),
             C.c_int(flip), C.byref(dproj), C.byref(close), C.byref(par))
    return bool(hit), xp, dproj.value, close.value, par.value


def is_face_flipped_3d(fi, fv, tv):
    return lib().orc_is_face_flipped_3d(C.c_int(fi), _ip(np.asarray(fv, np.int32)),
                                        _ip(np.asarray(tv, np.int32)))


def push_constant(mask, x, xtgt, distance, d):
    x = _f64(x)
    lib().orc_push_constant(C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)),
                            _dp(x), _dp(xtgt), C.c_long(x.shape[1]), C.c_double(distance),
                            C.c_double(d[0]), C.c_double(d[1]), C.c_double(d[2]))


def push_direction(mask, tgt, direction, distance):
    lib().orc_push_direction(C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)),
                             _dp(tgt), _dp(_f64(direction)), C.c_long(tgt.shape[1]),
                             C.c_double(distance))


def push_boris(pos, pos_prev, vel, efield, bfield, dt):
    """in place on [3][n] float64 arrays (src/pumipic_push.hpp:17-74)"""
    n = pos.shape[1]
    lib().orc_push_boris(C.c_int(n), _dp(pos), _dp(pos_prev), _dp(vel), _dp(_f64(efield)),
                         _dp(_f64(bfield)), C.c_double(dt))


def update_positions(x, xtgt):
    lib().orc_update_positions(C.c_int(x.shape[1]), _dp(x), _dp(xtgt), C.c_long(x.shape[1]))


def elliptical_setup(mask, x, b, phi, h, k, d):
    lib().orc_elliptical_setup(C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)),
                               _dp(x), C.c_long(x.shape[1]), b.ctypes.data_as(c_fp),
                               phi.ctypes.data_as(c_fp), C.c_double(h), C.c_double(k), C.c_double(d))


def elliptical_push(slot_elem, mask, xtgt, b, phi, class_ids, h, k, d, deg):
    lib().orc_elliptical_push(C.c_int(mask.shape[0]), _ip(np.ascontiguousarray(slot_elem, np.int32)),
                              _u8(np.ascontiguousarray(mask, np.uint8)), _dp(xtgt),
                              C.c_long(xtgt.shape[1]), b.ctypes.data_as(c_fp),
                              phi.ctypes.data_as(c_fp), _ip(np.ascontiguousarray(class_ids, np.int32)),
                              C.c_double(h), C.c_double(k), C.c_double(d), C.c_double(deg))


def set_unsafe_procs(mask, elems, safe, owner, self_rank):
    cap = mask.shape[0]
    ne = np.empty(cap, np.int32)
    npr = np.empty(cap, np.int32)
    lib().orc_set_unsafe_procs(C.c_int(cap), _u8(np.ascontiguousarray(mask, np.uint8)),
                               _ip(np.ascontiguousarray(elems, np.int32)),
                               _ip(np.ascontiguousarray(safe, np.int32)),
                               _ip(np.ascontiguousarray(owner, np.int32)), C.c_int(self_rank),
                               _ip(ne), _ip(npr))
    return ne, npr


# ---- gather (field interpolation) ----
def gather_tet_field(om, mask, x, elem_ids, field, dof):
    x = _f64(x); field = _f64(field)
    out = np.zeros((dof, x.shape[1]), np.float64)
    bad = lib().orc_gather_tet_field(om.h, C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)),
                                     _dp(x), C.c_long(x.shape[1]), _ip(np.ascontiguousarray(elem_ids, np.int32)),
                                     _dp(field), C.c_int(dof), _dp(out))
    return out, int(bad)


def interpolate2d_field(data, gridx0, gridz0, dx, dz, nx, nz, pos, cyl, ncomp=1, comp=0):
    f = lib().orc_interpolate2d_field
    f.restype = C.c_double
    return f(_dp(_f64(data)), C.c_double(gridx0), C.c_double(gridz0), C.c_double(dx), C.c_double(dz),
             C.c_int(nx), C.c_int(nz), _dp(_f64(pos)), C.c_int(int(cyl)), C.c_int(ncomp), C.c_int(comp))


def gather_grid2d_vector(mask, x, data3, gridx0, gridz0, dx, dz, nx, nz, cyl):
    x = _f64(x)
    out = np.zeros((3, x.shape[1]), np.float64)
    lib().orc_gather_grid2d_vector(C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)), _dp(x),
                                   C.c_long(x.shape[1]), _dp(_f64(data3)), C.c_double(gridx0),
                                   C.c_double(gridz0), C.c_double(dx), C.c_double(dz), C.c_int(nx),
                                   C.c_int(nz), C.c_int(int(cyl)), _dp(out))
    return out


def interpolate3d_field(x, y, z, gridx, gridy, gridz, data):
    f = lib().orc_interpolate3d_field
    f.restype = C.c_double
    return f(C.c_double(x), C.c_double(y), C.c_double(z), C.c_int(len(gridx)), C.c_int(len(gridy)),
             C.c_int(len(gridz)), _dp(_f64(gridx)), _dp(_f64(gridy)), _dp(_f64(gridz)), _dp(_f64(data)))


def gather_grid3d(mask, x, data, gridx, gridy, gridz):
    x = _f64(x)
    out = np.zeros(x.shape[1], np.float64)
    lib().orc_gather_grid3d(C.c_int(mask.shape[0]), _u8(np.ascontiguousarray(mask, np.uint8)), _dp(x),
                            C.c_long(x.shape[1]), _dp(_f64(data)), _dp(_f64(gridx)), _dp(_f64(gridy)),
                            _dp(_f64(gridz)), C.c_int(len(gridx)), C.c_int(len(gridy)), C.c_int(len(gridz)),
                            _dp(out))
    return out
