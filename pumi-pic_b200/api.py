"""Python handles over the C ABI (tests / bench plumbing).  Device arrays are torch tensors."""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib


def _torch():
    import torch
    return torch


def _stream():
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


# ------------------------------------------------------------------ host utilities
def host_derive_sides(dim, elem2verts):
    ev = np.ascontiguousarray(elem2verts, np.int32)
    ns = C.c_int32()
    e2s = capi.c_i32p()
    s2v = capi.c_i32p()
    check(lib().pp_host_derive_sides(dim, ev.shape[0], ev.ctypes.data_as(capi.c_i32p),
                                     C.byref(ns), C.byref(e2s), C.byref(s2v)))
    a = np.ctypeslib.as_array(e2s, shape=(ev.shape[0], dim + 1)).copy()
    b = np.ctypeslib.as_array(s2v, shape=(ns.value, dim)).copy()
    lib().pp_host_free(e2s)
    lib().pp_host_free(s2v)
    return a, b


def _host_gen(fn, n, length, dim):
    nv, ne = C.c_int32(), C.c_int32()
    co = capi.c_dp()
    ev = capi.c_i32p()
    check(fn(n, length, C.byref(nv), C.byref(co), C.byref(ne), C.byref(ev)))
    coords = np.ctypeslib.as_array(co, shape=(nv.value, dim)).copy()
    elems = np.ctypeslib.as_array(ev, shape=(ne.value, dim + 1)).copy()
    lib().pp_host_free(co)
    lib().pp_host_free(ev)
    return coords, elems


def host_kuhn_cube(n, length=1.0):
    return _host_gen(lib().pp_host_kuhn_cube, n, length, 3)


def host_plate(n, length=1.0):
    return _host_gen(lib().pp_host_plate, n, length, 2)


FULL, BFS, MINIMUM, NONE = 0, 1, 2, 3


def host_picpart_tags(dim, nverts, elem2verts, owner, nranks, rank, buffer_method=FULL,
                      safe_method=BFS, buffer_layers=3, safe_layers=1):
    ev = np.ascontiguousarray(elem2verts, np.int32)
    ow = np.ascontiguousarray(owner, np.int32)
    safe = np.empty(ev.shape[0], np.int32)
    part = np.empty(nranks, np.int32)
    check(lib().pp_host_picpart_tags(dim, nverts, ev.shape[0], ev.ctypes.data_as(capi.c_i32p),
                                     ow.ctypes.data_as(capi.c_i32p), nranks, rank, buffer_method,
                                     safe_method, buffer_layers, safe_layers,
                                     safe.ctypes.data_as(capi.c_i32p), part.ctypes.data_as(capi.c_i32p)))
    return safe, part


def host_picpart_tags_bridged(nbridges, elem2bridges, owner, nranks, rank, buffer_method=FULL,
                              safe_method=BFS, buffer_layers=3, safe_layers=1):
    """Input::bridge_dim: elem2bridges [nelems, k] = the elements' entities of the bridge dimension."""
    eb = np.ascontiguousarray(elem2bridges, np.int32)
    ow = np.ascontiguousarray(owner, np.int32)
    safe = np.empty(eb.shape[0], np.int32)
    part = np.empty(nranks, np.int32)
    check(lib().pp_host_picpart_tags_bridged(nbridges, eb.shape[0], eb.shape[1], eb.ctypes.data_as(capi.c_i32p),
                                             ow.ctypes.data_as(capi.c_i32p), nranks, rank, buffer_method,
                                             safe_method, buffer_layers, safe_layers,
                                             safe.ctypes.data_as(capi.c_i32p), part.ctypes.data_as(capi.c_i32p)))
    return safe, part


def host_entity_owners(nents, elem2ents, elem_owner, nranks):
    ee = np.ascontiguousarray(elem2ents, np.int32)
    ow = np.ascontiguousarray(elem_owner, np.int32)
    out = np.empty(nents, np.int32)
    check(lib().pp_host_entity_owners(nents, ee.shape[0], ee.shape[1], ee.ctypes.data_as(capi.c_i32p),
                                      ow.ctypes.data_as(capi.c_i32p), nranks,
                                      out.ctypes.data_as(capi.c_i32p)))
    return out


def host_picpart_extract(dim, coords, elem2verts, owner, nranks, has_part):
    """(elem_l2g, vert_l2g, elem2verts_local, coords_local) of the PICpart that buffers `has_part`."""
    co = np.ascontiguousarray(coords, np.float64)
    ev = np.ascontiguousarray(elem2verts, np.int32)
    ow = np.ascontiguousarray(owner, np.int32)
    hp = np.ascontiguousarray(has_part, np.int32)
    ne, nv = C.c_int32(), C.c_int32()
    el2g, vl2g, evl = capi.c_i32p(), capi.c_i32p(), capi.c_i32p()
    col = capi.c_dp()
    check(lib().pp_host_picpart_extract(dim, co.shape[0], ev.shape[0], co.ctypes.data_as(capi.c_dp),
                                        ev.ctypes.data_as(capi.c_i32p), ow.ctypes.data_as(capi.c_i32p),
                                        nranks, hp.ctypes.data_as(capi.c_i32p), C.byref(ne), C.byref(el2g),
                                        C.byref(nv), C.byref(vl2g), C.byref(evl), C.byref(col)))
    out = (np.ctypeslib.as_array(el2g, shape=(max(ne.value, 1),))[:ne.value].copy(),
           np.ctypeslib.as_array(vl2g, shape=(max(nv.value, 1),))[:nv.value].copy(),
           np.ctypeslib.as_array(evl, shape=(max(ne.value, 1), dim + 1))[:ne.value].copy(),
           np.ctypeslib.as_array(col, shape=(max(nv.value, 1), dim))[:nv.value].copy())
    for p in (el2g, vl2g, evl, col):
        lib().pp_host_free(p)
    return out


_TAG_DTYPES = {capi.PP_TAG_I8: np.int8, capi.PP_TAG_I32: np.int32, capi.PP_TAG_I64: np.int64,
               capi.PP_TAG_F64: np.float64}
_TAG_CODES = {np.dtype(v): k for k, v in _TAG_DTYPES.items()}


def _as_np(ptr, n, dtype):
    """Copy of n values behind a ctypes pointer (empty when n == 0 or the pointer is NULL)."""
    if n <= 0 or not ptr:
        return np.empty(0, dtype)
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    return np.frombuffer((C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr),
                         dtype=dtype).copy()


class HostMesh:
    """Host-side mesh with every entity dimension and its tags (what the reference keeps in an
    Omega_h::Mesh on the set-up side); reads and writes Omega_h `.osh` directories."""

    def __init__(self, handle, owned=True, keepalive=None):
        self.h = C.c_void_p(handle) if isinstance(handle, int) else handle
        self._owned = owned
        self._keepalive = keepalive

    @classmethod
    def read_osh(cls, path):
        h = C.c_void_p()
        check(lib().pp_host_mesh_read_osh(str(path).encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def from_elems(cls, dim, coords, elem2verts):
        co = np.ascontiguousarray(coords, np.float64)
        ev = np.ascontiguousarray(elem2verts, np.int32)
        h = C.c_void_p()
        check(lib().pp_host_mesh_from_elems(dim, co.shape[0], co.ctypes.data_as(capi.c_dp),
                                            ev.shape[0], ev.ctypes.data_as(capi.c_i32p), C.byref(h)))
        return cls(h)

    def write_osh(self, path):
        check(lib().pp_host_mesh_write_osh(self.h, str(path).encode()))

    @property
    def dim(self):
        return lib().pp_host_mesh_dim(self.h)

    def nents(self, d):
        return lib().pp_host_mesh_nents(self.h, d)

    def down(self, d):
        return _as_np(lib().pp_host_mesh_down(self.h, d), self.nents(d) * (d + 1), np.int32).reshape(-1, d + 1)

    def codes(self, d):
        return _as_np(lib().pp_host_mesh_codes(self.h, d), self.nents(d) * (d + 1), np.int8).reshape(-1, d + 1)

    def ent2verts(self, d):
        return _as_np(lib().pp_host_mesh_ent2verts(self.h, d), self.nents(d) * (d + 1), np.int32).reshape(-1, d + 1)

    def coords(self):
        return _as_np(lib().pp_host_mesh_coords(self.h), self.nents(0) * self.dim, np.float64).reshape(-1, self.dim)

    def tag_names(self, d):
        out = []
        t = capi.HostTag()
        for i in range(lib().pp_host_mesh_ntags(self.h, d)):
            check(lib().pp_host_mesh_tag_at(self.h, d, i, C.byref(t)))
            out.append(t.name.decode())
        return out

    def has_tag(self, d, name):
        return name in self.tag_names(d)

    def tag(self, d, name):
        t = capi.HostTag()
        check(lib().pp_host_mesh_find_tag(self.h, d, name.encode(), C.byref(t)))
        a = _as_np(t.data, t.nvalues, _TAG_DTYPES[t.type])
        return a.reshape(-1, t.ncomps) if t.ncomps > 1 else a

    def set_tag(self, d, name, values):
        a = np.ascontiguousarray(values)
        ncomps = 1 if a.ndim == 1 else a.shape[1]
        assert a.shape[0] == self.nents(d), "one row per entity"
        check(lib().pp_host_mesh_set_tag(self.h, d, name.encode(), ncomps, _TAG_CODES[a.dtype],
                                         _np_ptr(a)))

    def sides(self):
        """(elem2verts, elem2sides, side2verts) as pp_mesh_create wants them."""
        d = self.dim
        return self.ent2verts(d), self.down(d), self.ent2verts(d - 1) if d > 1 else None

    def __del__(self):
        try:
            if self._owned and self.h:
                lib().pp_host_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass


def host_read_partition(path, nelems, elem_class=None):
    """Owner per element from a `.ptn` / `.cpn` file (pumipic::Input, pumipic_input.cpp:44-89)."""
    out = np.empty(nelems, np.int32)
    cls = None if elem_class is None else np.ascontiguousarray(elem_class, np.int32)
    check(lib().pp_host_read_partition(str(path).encode(), nelems,
                                       None if cls is None else cls.ctypes.data_as(capi.c_i32p),
                                       out.ctypes.data_as(capi.c_i32p)))
    return out


class Picpart:
    """pumipic::Mesh on the host: the PICpart's mesh, tags and communication record."""

    _FIELDS = ("buffered_parts", "offset_ents_per_rank", "ent_to_comm_arr_index", "is_complete_part",
               "boundary_parts", "offset_bounded", "bounded_ent_ids", "ent_l2g")

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def build(cls, full, elem_owner, nranks, rank, buffer_method=BFS, safe_method=BFS,
              buffer_layers=-1, safe_layers=-1, bridge_dim=0):
        ow = np.ascontiguousarray(elem_owner, np.int32)
        assert ow.shape[0] == full.nents(full.dim)
        h = C.c_void_p()
        check(lib().pp_host_picpart_build_bridged(full.h, ow.ctypes.data_as(capi.c_i32p), nranks, rank,
                                                  buffer_method, safe_method, buffer_layers, safe_layers,
                                                  bridge_dim, C.byref(h)))
        return cls(h)

    @classmethod
    def read(cls, prefix, nranks, rank):
        h = C.c_void_p()
        check(lib().pp_host_picpart_read(str(prefix).encode(), nranks, rank, C.byref(h)))
        return cls(h)

    def write(self, prefix):
        check(lib().pp_host_picpart_write(self.h, str(prefix).encode()))

    @property
    def nranks(self):
        return lib().pp_host_picpart_nranks(self.h)

    @property
    def rank(self):
        return lib().pp_host_picpart_rank(self.h)

    @property
    def is_full_mesh(self):
        return bool(lib().pp_host_picpart_is_full_mesh(self.h))

    def mesh(self):
        return HostMesh(lib().pp_host_picpart_mesh(self.h), owned=False, keepalive=self)

    def dim_info(self, d):
        """dict of the per-dimension members pumipic::Mesh keeps (pumipic_mesh.hpp:118-143)."""
        s = capi.PicpartDim()
        check(lib().pp_host_picpart_get(self.h, d, C.byref(s)))
        nr = self.nranks
        has = s.nents > 0 or s.num_entities > 0
        sizes = {"buffered_parts": max(s.num_cores, 0), "offset_ents_per_rank": nr + 1 if has else 0,
                 "ent_to_comm_arr_index": s.nents, "is_complete_part": nr if has else 0,
                 "boundary_parts": s.num_boundaries, "offset_bounded": s.n_offset_bounded,
                 "bounded_ent_ids": s.n_bounded_ent_ids, "ent_l2g": s.nents}
        out = {"num_entities": s.num_entities, "nents": s.nents, "num_cores": s.num_cores,
               "num_bounds": s.num_bounds, "num_boundaries": s.num_boundaries}
        for f in self._FIELDS:
            out[f] = _as_np(getattr(s, f), sizes[f], np.int32)
        return out

    def sbars(self):
        """{global sbar id: tuple of parts} for the safe-zone overlaps this part belongs to, and max_sbar."""
        n, mx = C.c_int32(), C.c_int32()
        ids, off, parts = capi.c_i32p(), capi.c_i32p(), capi.c_i32p()
        check(lib().pp_host_picpart_sbars(self.h, C.byref(n), C.byref(ids), C.byref(off), C.byref(parts),
                                          C.byref(mx)))
        ids_a = _as_np(ids, n.value, np.int32)
        off_a = _as_np(off, n.value + 1, np.int32)
        parts_a = _as_np(parts, int(off_a[-1]) if n.value else 0, np.int32)
        return {int(ids_a[i]): tuple(int(x) for x in parts_a[off_a[i]:off_a[i + 1]])
                for i in range(n.value)}, mx.value

    def __del__(self):
        try:
            if self.h:
                lib().pp_host_picpart_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------ mesh
class Mesh:
    """pumipic::Mesh / o::Mesh stand-in: owns a pp_mesh built from host numpy arrays."""

    def __init__(self, dim, coords, elem2verts, elem2sides, side2verts, elem_class=None):
        coords = np.ascontiguousarray(coords, np.float64)
        ev = np.ascontiguousarray(elem2verts, np.int32)
        es = np.ascontiguousarray(elem2sides, np.int32)
        sv = np.ascontiguousarray(side2verts, np.int32)
        cls = None if elem_class is None else np.ascontiguousarray(elem_class, np.int32)
        d = capi.MeshDesc(dim, coords.shape[0], ev.shape[0], sv.shape[0], coords.ctypes.data,
                          ev.ctypes.data, es.ctypes.data, sv.ctypes.data,
                          0 if cls is None else cls.ctypes.data, capi.PP_HOST)
        self.h = C.c_void_p()
        check(lib().pp_mesh_create(C.byref(d), _stream(), C.byref(self.h)))
        self.dim = dim
        self.nelems = ev.shape[0]
        self.nverts = coords.shape[0]
        self.nsides = sv.shape[0]

    def info(self):
        i = capi.MeshInfo()
        check(lib().pp_mesh_get_info(self.h, C.byref(i)))
        return i

    def arrays(self):
        """Derived device arrays copied back to host numpy (test helper)."""
        torch = _torch()
        a = capi.MeshArrays()
        check(lib().pp_mesh_get_arrays(self.h, C.byref(a)))
        torch.cuda.synchronize()

        def fetch(ptr, n, dtype):
            tdt = {np.float64: torch.float64, np.int8: torch.int8, np.int32: torch.int32}[dtype]
            if n == 0:
                return np.zeros(0, dtype)
            return _tensor_from_ptr(ptr, (n,), tdt, self).cpu().numpy()
        res = {
            "measure": fetch(a.measure, self.nelems, np.float64),
            "exposed": fetch(a.exposed, self.nsides, np.int8),
            "side2elem": fetch(a.side2elem, 2 * self.nsides, np.int32).reshape(-1, 2),
            "dual_off": fetch(a.dual_off, self.nelems + 1, np.int32),
        }
        res["dual"] = fetch(a.dual, int(res["dual_off"][-1]), np.int32)
        return res

    def set_picpart(self, safe, owner, self_rank):
        s = np.ascontiguousarray(safe, np.int32)
        o = np.ascontiguousarray(owner, np.int32)
        check(lib().pp_mesh_set_picpart(self.h, _np_ptr(s), _np_ptr(o), self_rank, capi.PP_HOST,
                                        _stream()))
        _torch().cuda.synchronize()

    def __del__(self):
        try:
            lib().pp_mesh_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------ particle structure
_NP_OF = {(8, "f"): np.float64, (4, "f"): np.float32, (4, "i"): np.int32, (8, "i"): np.int64}


class ParticleStructure:
    """ps::ParticleStructure<MemberTypes<...>> stand-in.

    members: list of (numpy dtype, ncomp), e.g. [(np.float64, 3), (np.float64, 3), (np.int32, 1)]
    """

    def __init__(self, kind, members, ppe, elem_gids=None, particle_elements=None,
                 particle_info=None, team_size=32, sigma=0x7fffffff, V=1024, config=None):
        self.members = [(np.dtype(dt), int(nc)) for dt, nc in members]
        ppe = np.ascontiguousarray(ppe, np.int32)
        cfg = capi.PsConfig()
        lib().pp_ps_config_default(C.byref(cfg), kind)
        cfg.team_size, cfg.sigma, cfg.V = team_size, sigma, V
        for k, v in (config or {}).items():
            setattr(cfg, k, v)
        md = (capi.MemberDesc * len(members))(*[capi.MemberDesc(dt.itemsize, nc)
                                                for dt, nc in self.members])
        np_ = int(ppe.sum()) if particle_elements is None else len(particle_elements)
        gids = None if elem_gids is None else np.ascontiguousarray(elem_gids, np.int64)
        pel = None if particle_elements is None else np.ascontiguousarray(particle_elements, np.int32)
        info = None
        keep = []
        if particle_info is not None:
            info = (C.c_void_p * len(members))()
            for i, a in enumerate(particle_info):
                a = np.ascontiguousarray(a, self.members[i][0])
                keep.append(a)
                info[i] = a.ctypes.data
        self.h = C.c_void_p()
        check(lib().pp_ps_create(C.byref(cfg), len(members), md, ppe.shape[0], np_, _np_ptr(ppe),
                                 None if gids is None else _np_ptr(gids),
                                 None if pel is None else _np_ptr(pel), info, capi.PP_HOST,
                                 _stream(), C.byref(self.h)))

    nelems = property(lambda self: lib().pp_ps_nelems(self.h))
    nptcls = property(lambda self: lib().pp_ps_nptcls(self.h))
    capacity = property(lambda self: lib().pp_ps_capacity(self.h))
    numrows = property(lambda self: lib().pp_ps_numrows(self.h))

    def get(self, i):
        """Segment of member i as a torch view [ncomp, stride] over the structure's own memory."""
        torch = _torch()
        base = C.c_void_p()
        stride = C.c_int64()
        check(lib().pp_ps_member(self.h, i, C.byref(base), C.byref(stride)))
        dt, nc = self.members[i]
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
               np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[dt]
        n = nc * stride.value
        if n == 0:
            return torch.empty((nc, 0), dtype=tdt, device="cuda")
        return _tensor_from_ptr(base.value, (nc, stride.value), tdt, self)

    def set_rebuild_remap(self, src_member):
        """one-shot: after the next rebuild / migrate member i holds what member src_member[i] held
        (-1: zeros); e.g. [1, -1, 2, 3] = updatePtclPositions folded into the record move"""
        a = np.ascontiguousarray(src_member, np.int32)
        check(lib().pp_ps_set_rebuild_remap(self.h, _np_ptr(a), int(a.shape[0])))

    def rebuild(self, new_element, new_particle_elements=None, new_particle_info=None):
        """ParticleStructure::rebuild: all arguments are device tensors ([ncomp, n_new] members)."""
        n_new = 0 if new_particle_elements is None else int(new_particle_elements.shape[0])
        info = None
        if n_new:
            info = (C.c_void_p * len(self.members))(*[t.data_ptr() for t in new_particle_info])
        check(lib().pp_ps_rebuild(self.h, _ptr(new_element), n_new, _ptr(new_particle_elements),
                                  info, _stream()))

    def layout(self):
        lay = capi.PsLayout()
        check(lib().pp_ps_get_layout(self.h, _stream(), C.byref(lay)))
        return lay

    def get_pids(self):
        """ParticleStructure::getPIDs: (pids[nptcls], offsets[nelems+1]) as cuda int32 tensors."""
        torch = _torch()
        pids = torch.empty(self.nptcls, dtype=torch.int32, device="cuda")
        offsets = torch.empty(self.nelems + 1, dtype=torch.int32, device="cuda")
        check(lib().pp_ps_get_pids(self.h, _ptr(pids), _ptr(offsets), _stream()))
        return pids, offsets

    def slot_elem_and_mask(self):
        """(slot_elem[cap] int32, mask[cap] uint8) on the host (test helper)."""
        torch = _torch()
        lay = self.layout()
        torch.cuda.synchronize()
        cap = lay.capacity
        if cap == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.uint8)
        se = _tensor_from_ptr(lay.slot_elem, (cap,), torch.int32, self).cpu().numpy()
        nw = (cap + 31) // 32
        mb = _tensor_from_ptr(lay.mask_bits, (nw,), torch.int32, self).cpu().numpy().view(np.uint32)
        bits = ((mb[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)
        return se, bits.ravel()[:cap].copy()

    def __del__(self):
        try:
            lib().pp_ps_destroy(self.h)
        except Exception:
            pass


class _CudaArrayView:
    """Exposes foreign device memory through __cuda_array_interface__ so torch can wrap it."""

    def __init__(self, ptr, shape, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": tuple(shape),
                                         "typestr": typestr, "version": 2, "strides": None}


def _tensor_from_ptr(ptr, shape, tdt, owner):
    torch = _torch()
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4",
               torch.int64: "<i8", torch.int8: "|i1"}[tdt]
    return torch.as_tensor(_CudaArrayView(ptr, shape, typestr, owner), device="cuda")


# ------------------------------------------------------------------ hot path calls
class SearchResult:
    def __init__(self, stats):
        for f, _ in capi.SearchStats._fields_:
            setattr(self, f, getattr(stats, f))

    def __repr__(self):
        return "SearchResult(" + ", ".join("%s=%s" % (f, getattr(self, f))
                                           for f, _ in capi.SearchStats._fields_) + ")"


def _search_args(x_orig, x_tgt, elem_ids, elem_ids_empty, variant, require_intersection,
                 inter_faces, inter_points, looplimit):
    return capi.SearchArgs(variant, _ptr(x_orig).value, _ptr(x_tgt).value, x_tgt.shape[1],
                           _ptr(elem_ids).value, int(bool(elem_ids_empty)),
                           int(bool(require_intersection)), _ptr(inter_faces).value,
                           _ptr(inter_points).value, looplimit)


def search_mesh(mesh, ps, x_orig, x_tgt, elem_ids, elem_ids_empty=False,
                variant=capi.PP_SEARCH_NEW, require_intersection=False, inter_faces=None,
                inter_points=None, looplimit=0, sync=True):
    a = _search_args(x_orig, x_tgt, elem_ids, elem_ids_empty, variant, require_intersection,
                     inter_faces, inter_points, looplimit)
    st = capi.SearchStats()
    check(lib().pp_search_mesh(mesh.h, ps.h, C.byref(a), C.byref(st) if sync else None, _stream()))
    return SearchResult(st) if sync else None


def trace_particle_through_mesh(mesh, ps, x_orig, x_tgt, elem_ids, elem_ids_empty=False,
                                require_intersection=False, inter_faces=None, inter_points=None,
                                looplimit=0, handler=None):
    """adjacency.tpp:461-640 phase by phase (pp_trace_*).  handler(elem_ids, inter_faces, last_exit,
    inter_points, ptcl_done) stands where the reference's `Func` stands; None = the stock
    RemoveParticleOnGeometricModelExit.  Returns (found, loops, not_in_elem, not_found)."""
    torch = _torch()
    a = _search_args(x_orig, x_tgt, elem_ids, elem_ids_empty, capi.PP_SEARCH_NEW,
                     require_intersection, inter_faces, inter_points, looplimit)
    cap = ps.capacity
    done = torch.empty(max(cap, 1), dtype=torch.int32, device="cuda")
    last_exit = torch.empty(max(cap, 1), dtype=torch.int32, device="cuda")
    L, st = lib(), _stream()
    n = C.c_int32()
    check(L.pp_trace_begin(mesh.h, ps.h, C.byref(a), _ptr(done), _ptr(last_exit), C.byref(n), st))
    not_in, loops, lost, found = n.value, 0, 0, False
    while not found:
        check(L.pp_trace_find_exit_face(mesh.h, ps.h, C.byref(a), _ptr(done), _ptr(last_exit), st))
        if handler is None:
            check(L.pp_trace_check_model_intersection(mesh.h, ps.h, C.byref(a), _ptr(done),
                                                      _ptr(last_exit), st))
        else:
            handler(elem_ids, inter_faces, last_exit, inter_points, done)
        check(L.pp_trace_set_new_element(mesh.h, ps.h, C.byref(a), _ptr(done), _ptr(last_exit), st))
        check(L.pp_trace_pending(mesh.h, ps.h, C.byref(a), _ptr(done), _ptr(last_exit), 0, C.byref(n), st))
        found = n.value == 0
        loops += 1
        if loops > 1000000:      # a handler that never finishes its particles must not hang the caller
            raise capi.PumipicError("trace_particle_through_mesh: no progress after %d iterations" % loops)
        if looplimit and loops >= looplimit:
            check(L.pp_trace_pending(mesh.h, ps.h, C.byref(a), _ptr(done), _ptr(last_exit), 1,
                                     C.byref(n), st))
            lost = n.value
            break
    return found, loops, not_in, lost


def push_direction_search(mesh, ps, direction, distance, x_orig, x_tgt, elem_ids,
                          elem_ids_empty=False, require_intersection=False, inter_faces=None,
                          inter_points=None, looplimit=0, sync=True, from_orig=False):
    a = _search_args(x_orig, x_tgt, elem_ids, elem_ids_empty, capi.PP_SEARCH_NEW,
                     require_intersection, inter_faces, inter_points, looplimit)
    st = capi.SearchStats()
    check(lib().pp_push_direction_search(mesh.h, ps.h, _ptr(direction), distance,
                                         int(bool(from_orig)), C.byref(a),
                                         C.byref(st) if sync else None, _stream()))
    return SearchResult(st) if sync else None


def push_direction_search_host(mesh, ps, h_x, h_dir, h_xtgt, h_ids, distance, looplimit=0, nparts=0,
                               sync=True):
    """Host-buffer step: h_* are CPU tensors (pinned for overlap) of shape [3, stride] / [capacity]."""
    st = capi.SearchStats()
    check(lib().pp_push_direction_search_host(mesh.h, ps.h, _ptr(h_x), _ptr(h_dir), _ptr(h_xtgt),
                                              _ptr(h_ids), h_x.shape[1], distance, looplimit, nparts,
                                              C.byref(st) if sync else None, _stream()))
    return SearchResult(st) if sync else None


def push_boris(pos, pos_prev, vel, efield, bfield, dt):
    check(lib().pp_push_boris(pos.shape[1], pos.shape[1], _ptr(pos), _ptr(pos_prev), _ptr(vel),
                              _ptr(efield), _ptr(bfield), dt, _stream()))


def push_constant(ps, x, xtgt, distance, d):
    check(lib().pp_push_constant(ps.h, _ptr(x), _ptr(xtgt), x.shape[1], distance, d[0], d[1], d[2],
                                 _stream()))


def push_direction(ps, tgt, direction, distance):
    check(lib().pp_push_direction(ps.h, _ptr(tgt), _ptr(direction), tgt.shape[1], distance,
                                  _stream()))


def update_positions(ps, x, xtgt):
    check(lib().pp_update_positions(ps.h, _ptr(x), _ptr(xtgt), x.shape[1], _stream()))


def push_from(ps, x, xtgt, direction, distance):
    check(lib().pp_push_from(ps.h, _ptr(x), _ptr(xtgt), _ptr(direction), x.shape[1], distance,
                             _stream()))


def elliptical_setup(ps, x, b, phi, h, k, d):
    check(lib().pp_push_elliptical_setup(ps.h, _ptr(x), x.shape[1], _ptr(b), _ptr(phi), h, k, d,
                                         _stream()))


def elliptical_push(mesh, ps, xtgt, b, phi, h, k, d, deg):
    check(lib().pp_push_elliptical(mesh.h, ps.h, _ptr(xtgt), xtgt.shape[1], _ptr(b), _ptr(phi),
                                   h, k, d, deg, _stream()))


def set_unsafe_procs(mesh, ps, elems):
    torch = _torch()
    ne = torch.empty(ps.capacity, dtype=torch.int32, device="cuda")
    npr = torch.empty(ps.capacity, dtype=torch.int32, device="cuda")
    check(lib().pp_set_unsafe_procs(mesh.h, ps.h, _ptr(elems), _ptr(ne), _ptr(npr), _stream()))
    return ne, npr


def gather_tet_field(mesh, ps, x, elem_ids, field, dof):
    """interpolate3dFieldTet for every masked particle; returns (out [dof, stride], n_outside)."""
    torch = _torch()
    out = torch.zeros(dof, x.shape[1], dtype=torch.float64, device="cuda")
    bad = C.c_int32(0)
    check(lib().pp_gather_tet_field(mesh.h, ps.h, _ptr(x), x.shape[1], _ptr(elem_ids), _ptr(field), dof,
                                    _ptr(out), C.byref(bad), _stream()))
    return out, bad.value


def gather_grid2d(ps, x, data, gridx0, gridz0, dx, dz, nx, nz, cyl, ncomp=1, comp=0):
    torch = _torch()
    out = torch.zeros(x.shape[1], dtype=torch.float64, device="cuda")
    check(lib().pp_gather_grid2d(ps.h, _ptr(x), x.shape[1], _ptr(data), gridx0, gridz0, dx, dz, nx, nz,
                                 int(bool(cyl)), ncomp, comp, _ptr(out), _stream()))
    return out


def gather_grid2d_vector(ps, x, data3, gridx0, gridz0, dx, dz, nx, nz, cyl):
    torch = _torch()
    out = torch.zeros(3, x.shape[1], dtype=torch.float64, device="cuda")
    check(lib().pp_gather_grid2d_vector(ps.h, _ptr(x), x.shape[1], _ptr(data3), gridx0, gridz0, dx, dz,
                                        nx, nz, int(bool(cyl)), _ptr(out), _stream()))
    return out


def gather_grid3d(ps, x, data, gridx, gridy, gridz):
    torch = _torch()
    out = torch.zeros(x.shape[1], dtype=torch.float64, device="cuda")
    check(lib().pp_gather_grid3d(ps.h, _ptr(x), x.shape[1], _ptr(data), _ptr(gridx), _ptr(gridy),
                                 _ptr(gridz), gridx.numel(), gridy.numel(), gridz.numel(), _ptr(out),
                                 _stream()))
    return out


def gyro_ring_map(mesh, rmax, nrings, ppr, theta_deg):
    torch = _torch()
    out = torch.empty(3 * mesh.nverts * nrings * ppr, dtype=torch.int32, device="cuda")
    st = capi.SearchStats()
    check(lib().pp_gyro_ring_map(mesh.h, rmax, nrings, ppr, theta_deg, _ptr(out), C.byref(st),
                                 _stream()))
    return out, SearchResult(st)


def gyro_scatter(mesh, ps, v2v, rmax, nrings, ppr):
    torch = _torch()
    out = torch.empty(mesh.nverts, dtype=torch.float64, device="cuda")
    check(lib().pp_gyro_scatter(mesh.h, ps.h, _ptr(v2v), rmax, nrings, ppr, _ptr(out), _stream()))
    return out


def gyro_interleave(fwd, bkwd):
    torch = _torch()
    out = torch.empty(2 * fwd.shape[0], dtype=torch.float64, device="cuda")
    check(lib().pp_gyro_interleave(_ptr(fwd), _ptr(bkwd), fwd.shape[0], _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------ communication
# ------------------------------------------------------------------ phase timers (ppTiming.hpp)
def timing_enable(on=True, rank=0, verbosity=0):
    lib().pp_timing_set_rank(rank)
    lib().pp_timing_set_verbosity(verbosity)
    lib().pp_timing_enable(1 if on else 0)


def timing_reset():
    lib().pp_timing_reset()


def timing_table():
    """{label: {"total_s", "min_s", "max_s", "calls", "avg_ms"}} of the phases recorded so far"""
    out = {}
    for i in range(lib().pp_timing_count()):
        name = C.create_string_buffer(256)
        tot, mn, mx, sq = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        n = C.c_int64()
        check(lib().pp_timing_get(i, name, 256, C.byref(tot), C.byref(mn), C.byref(mx), C.byref(sq), C.byref(n)))
        out[name.value.decode()] = {"total_s": tot.value, "min_s": mn.value, "max_s": mx.value, "calls": n.value,
                                    "avg_ms": 1e3 * tot.value / max(1, n.value)}
    return out


class Comm:
    """Communicator of the C ABI.
    Default: NCCL; the unique id travels over torch.distributed (plumbing).
    hosted=True: pp_comm_create_hosted with torch.distributed's all-gather as the application's
    bootstrap callback (what an MPI code does with MPI_Allgather); nccl=False then builds a
    communicator without NCCL (peer-memory windows only), which also works between processes that
    share one GPU."""

    _DT = None

    def __init__(self, nranks=None, rank=None, hosted=False, nccl=True):
        torch = _torch()
        import torch.distributed as dist
        if nranks is None:
            nranks = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.nranks, self.rank = nranks, rank
        self.h = C.c_void_p()
        self._cb = None
        if hosted:
            def allgather(_ctx, send, recv, nbytes):
                try:
                    parts = [None] * nranks
                    dist.all_gather_object(parts, C.string_at(send, nbytes))
                    C.memmove(recv, b"".join(parts), nbytes * nranks)
                    return 0
                except Exception:          # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            self._cb = capi.HOST_ALLGATHER_FN(allgather)   # kept alive with the communicator
            check(lib().pp_comm_create_hosted(nranks, rank, C.cast(self._cb, C.c_void_p), None,
                                              1 if nccl else 0, C.byref(self.h)))
            return
        uid = None
        if nranks > 1:
            buf = (C.c_uint8 * 128)()
            if rank == 0:
                check(lib().pp_comm_unique_id(buf))
            obj = [bytes(buf)]
            dist.broadcast_object_list(obj, src=0)
            uid = (C.c_uint8 * 128).from_buffer_copy(obj[0])
        check(lib().pp_comm_create(nranks, rank, uid, C.byref(self.h)))

    @staticmethod
    def _dtype(t):
        torch = _torch()
        return {torch.int32: capi.PP_INT32, torch.int64: capi.PP_INT64,
                torch.float32: capi.PP_FLOAT32, torch.float64: capi.PP_FLOAT64}[t.dtype]

    def allreduce(self, t, op=capi.PP_SUM):
        check(lib().pp_comm_allreduce(self.h, _ptr(t), _ptr(t), t.numel(), self._dtype(t), op, _stream()))
        return t

    def alltoall(self, send, recv):
        check(lib().pp_comm_alltoall(self.h, _ptr(send), _ptr(recv), send.numel() // self.nranks,
                                     self._dtype(send), _stream()))
        return recv

    def array_reduce(self, arr, nents, nvals, op, ent_owner=None):
        check(lib().pp_comm_array_reduce(self.h, _ptr(arr), nents, nvals, self._dtype(arr), op,
                                         _ptr(ent_owner), _stream()))
        return arr

    def set_p2p_window(self, bytes_per_peer):
        """size of the peer-memory window's segments; before the first migration"""
        check(lib().pp_comm_set_p2p_window(self.h, int(bytes_per_peer)))

    @property
    def p2p_active(self):
        """True once the first migration has mapped the peer-memory windows (NVLink path)"""
        return bool(lib().pp_comm_p2p_active(self.h))

    def plan(self, ent_gids, ent_owner):
        """Owner fan-in / fan-out plan for comm arrays of a partially buffered PICpart."""
        return CommPlan(self, ent_gids, ent_owner)

    def __del__(self):
        try:
            lib().pp_comm_destroy(self.h)
        except Exception:
            pass


def migrate(ps, comm, new_element, new_process, new_particle_elements=None, new_particle_info=None):
    n_new = 0 if new_particle_elements is None else int(new_particle_elements.shape[0])
    info = None
    if n_new:
        info = (C.c_void_p * len(ps.members))(*[t.data_ptr() for t in new_particle_info])
    st = capi.MigrateStats()
    check(lib().pp_ps_migrate(ps.h, comm.h, _ptr(new_element), _ptr(new_process), n_new,
                              _ptr(new_particle_elements), info, C.byref(st), _stream()))
    return st.sent, st.received


class CommPlan:
    """Mesh::setupComm + reduceCommArray for one entity dimension (pp_comm_plan_*)."""

    def __init__(self, comm, ent_gids, ent_owner):
        g = np.ascontiguousarray(ent_gids, np.int64)
        o = np.ascontiguousarray(ent_owner, np.int32)
        self.comm = comm
        self.nents = g.shape[0]
        self.h = C.c_void_p()
        check(lib().pp_comm_plan_create(comm.h, g.shape[0], _np_ptr(g), _np_ptr(o), capi.PP_HOST, _stream(),
                                        C.byref(self.h)))

    def counts(self):
        a, b = C.c_int64(), C.c_int64()
        check(lib().pp_comm_plan_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reduce(self, arr, nvals, op):
        check(lib().pp_comm_plan_reduce(self.h, _ptr(arr), nvals, Comm._dtype(arr), op, _stream()))
        return arr

    def __del__(self):
        try:
            lib().pp_comm_plan_destroy(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------ particle load balancing
def _sbar_arrays(table):
    ids = np.asarray(sorted(table), np.int32)
    off = np.zeros(len(ids) + 1, np.int32)
    parts = []
    for i, g in enumerate(ids):
        parts.extend(table[int(g)])
        off[i + 1] = len(parts)
    return ids, off, np.asarray(parts, np.int32)


def host_lb_plan(nranks, table, vert_weight, forced=None, tol=1.05, step_factor=0.3):
    """pp_host_lb_plan: `table` = {sbar id: sorted parts}; returns ([(vertex, part, weight)],
    (imbalance before, imbalance planned))."""
    ids, off, parts = _sbar_arrays(table)
    w = np.ascontiguousarray(vert_weight, np.float64)
    f = None if forced is None else np.ascontiguousarray(forced, np.float64)
    n = C.c_int32()
    sv, sp, sw = capi.c_i32p(), capi.c_i32p(), capi.c_dp()
    imb = (C.c_double * 2)()
    check(lib().pp_host_lb_plan(nranks, len(ids), ids.ctypes.data_as(capi.c_i32p),
                                off.ctypes.data_as(capi.c_i32p), parts.ctypes.data_as(capi.c_i32p),
                                w.shape[0], w.ctypes.data_as(capi.c_dp),
                                None if f is None else f.ctypes.data_as(capi.c_dp), tol, step_factor,
                                C.byref(n), C.byref(sv), C.byref(sp), C.byref(sw), imb))
    out = [(int(sv[i]), int(sp[i]), float(sw[i])) for i in range(n.value)]
    for p in (sv, sp, sw):
        lib().pp_host_free(p)
    return out, (imb[0], imb[1])


class Balancer:
    """pumipic::ParticleBalancer (pumipic_lb.hpp:32-115) for one part (pp_balancer_*)."""

    def __init__(self, nranks, rank, table, elem_sbar, elem_owner, comm=None):
        """table = {sbar id: sorted parts}: the global table, or with a multi-rank `comm` the
        regions this part knows (Picpart.sbars()[0]); elem_sbar / elem_owner: numpy or cuda int32."""
        ids, off, parts = _sbar_arrays(table)
        host = isinstance(elem_sbar, np.ndarray)
        if host:
            es = np.ascontiguousarray(elem_sbar, np.int32)
            eo = np.ascontiguousarray(elem_owner, np.int32)
            pes, peo, ne = _np_ptr(es), _np_ptr(eo), es.shape[0]
        else:
            pes, peo, ne = _ptr(elem_sbar), _ptr(elem_owner), elem_sbar.shape[0]
        self.nranks, self.rank, self.nelems = nranks, rank, ne
        self.h = C.c_void_p()
        check(lib().pp_balancer_create(nranks, rank, len(ids), ids.ctypes.data_as(capi.c_i32p),
                                       off.ctypes.data_as(capi.c_i32p),
                                       parts.ctypes.data_as(capi.c_i32p), ne, pes, peo,
                                       capi.PP_HOST if host else capi.PP_DEVICE,
                                       None if comm is None else comm.h, _stream(), C.byref(self.h)))

    def info(self):
        """(graph vertices of all parts, this part's global vertex ids, their sbar ids)"""
        nv, nl = C.c_int32(), C.c_int32()
        lv, ls = capi.c_i32p(), capi.c_i32p()
        check(lib().pp_balancer_info(self.h, C.byref(nv), C.byref(nl), C.byref(lv), C.byref(ls)))
        return nv.value, _as_np(lv, nl.value, np.int32), _as_np(ls, nl.value, np.int32)

    def weights(self):
        """the global weight vector as a cuda float64 tensor view: [nverts] then [nranks] forced"""
        torch = _torch()
        p, n = C.c_void_p(), C.c_int64()
        check(lib().pp_balancer_weights(self.h, C.byref(p), C.byref(n)))
        return _tensor_from_ptr(p.value, (n.value,), torch.float64, self)

    def add_weights(self, ps, new_elems, new_procs):
        check(lib().pp_balancer_add_weights_ps(self.h, ps.h, _ptr(new_elems), _ptr(new_procs), _stream()))

    def add_weights_array(self, ptcls_per_elem):
        check(lib().pp_balancer_add_weights_array(self.h, _ptr(ptcls_per_elem), _stream()))

    def balance(self, comm=None, tol=1.05, step_factor=0.3):
        check(lib().pp_balancer_balance(self.h, None if comm is None else comm.h, tol, step_factor,
                                        _stream()))
        return self.plan()

    def plan(self):
        """([(sbar id, target part, weight)] of this part, (imbalance before, planned))"""
        n = C.c_int32()
        sb, pt, wt = capi.c_i32p(), capi.c_i32p(), capi.c_dp()
        imb = (C.c_double * 2)()
        check(lib().pp_balancer_plan(self.h, C.byref(n), C.byref(sb), C.byref(pt), C.byref(wt), imb))
        return [(int(sb[i]), int(pt[i]), float(wt[i])) for i in range(n.value)], (imb[0], imb[1])

    def select(self, ps, new_elems, new_procs):
        check(lib().pp_balancer_select_ps(self.h, ps.h, _ptr(new_elems), _ptr(new_procs), _stream()))
        return new_procs

    def select_array(self, ptcls_per_elem, nptcls):
        torch = _torch()
        out = torch.empty(nptcls, dtype=torch.int32, device="cuda")
        check(lib().pp_balancer_select_array(self.h, _ptr(ptcls_per_elem), nptcls, _ptr(out), _stream()))
        return out

    def repartition(self, comm, ps, tol, new_elems, new_procs, step_factor=0.3):
        check(lib().pp_balancer_repartition(self.h, None if comm is None else comm.h, ps.h, tol,
                                            _ptr(new_elems), _ptr(new_procs), step_factor, _stream()))
        return new_procs

    def partition(self, comm, ptcls_per_elem, tol, step_factor=0.3):
        torch = _torch()
        nptcls = int(ptcls_per_elem.sum().item())
        out = torch.empty(nptcls, dtype=torch.int32, device="cuda")
        check(lib().pp_balancer_partition(self.h, None if comm is None else comm.h,
                                          _ptr(ptcls_per_elem), nptcls, tol, step_factor, _ptr(out),
                                          _stream()))
        return out

    def __del__(self):
        try:
            if self.h:
                lib().pp_balancer_destroy(self.h)
                self.h = None
        except Exception:
            pass
