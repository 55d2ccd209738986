"""ctypes view of include/pumipic_b200.h (the C ABI of libpumipic_b200.so).

Thin plumbing only: structs, prototypes and a status check.  PyTorch supplies device memory
(tensor.data_ptr()) and streams; no compute happens in Python.  The library must exist:
there is no CPU fallback and importing this module without the built .so raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# pp_host_allgather_fn of include/pumipic_b200.h
HOST_ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)

LIB_PATH = os.environ.get("PUMIPIC_B200_LIB", os.path.join(HERE, "libpumipic_b200.so"))

PP_OK = 0
PP_HOST, PP_DEVICE = 0, 1
PP_PS_SCS, PP_PS_CSR, PP_PS_CABM, PP_PS_DPS = 0, 1, 2, 3
PP_SEARCH_NEW, PP_SEARCH_2D_LEGACY, PP_SEARCH_3D_LEGACY, PP_SEARCH_3D = 0, 1, 2, 3

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_dp = C.POINTER(C.c_double)


class MeshDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nverts", C.c_int32), ("nelems", C.c_int32),
                ("nsides", C.c_int32), ("coords", C.c_void_p), ("elem2verts", C.c_void_p),
                ("elem2sides", C.c_void_p), ("side2verts", C.c_void_p),
                ("elem_class", C.c_void_p), ("memspace", C.c_int32)]


class MeshInfo(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nverts", C.c_int32), ("nelems", C.c_int32),
                ("nsides", C.c_int32), ("tol", C.c_double), ("min_measure", C.c_double),
                ("n_exposed_sides", C.c_int32), ("walk_table_bytes", C.c_int64)]


class MeshArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("coords", "elem2verts", "elem2sides", "side2verts", "elem_class", "measure",
                 "exposed", "side2elem", "dual_off", "dual")]


class MemberDesc(C.Structure):
    _fields_ = [("scalar_bytes", C.c_int32), ("ncomp", C.c_int32)]


class PsConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("team_size", C.c_int32), ("sigma", C.c_int32),
                ("V", C.c_int32), ("shuffle_padding", C.c_double), ("extra_padding", C.c_double),
                ("minimize_size", C.c_double), ("padding_strat", C.c_int32),
                ("always_realloc", C.c_int32)]


class PsLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "C", "V", "nchunks", "nslices", "nrows", "capacity", "nelems", "nptcls")] + \
               [(n, C.c_void_p) for n in
                ("offsets", "slice_to_chunk", "row_to_element", "element_to_row", "mask_bits",
                 "slot_elem")]


class SearchArgs(C.Structure):
    _fields_ = [("variant", C.c_int32), ("x_orig", C.c_void_p), ("x_tgt", C.c_void_p),
                ("stride", C.c_int64), ("elem_ids", C.c_void_p), ("elem_ids_empty", C.c_int32),
                ("require_intersection", C.c_int32), ("inter_faces", C.c_void_p),
                ("inter_points", C.c_void_p), ("looplimit", C.c_int32)]


class SearchStats(C.Structure):
    _fields_ = [("found", C.c_int32), ("loops", C.c_int32), ("not_in_elem", C.c_int32),
                ("not_found", C.c_int32), ("aborted", C.c_int32), ("active", C.c_int32),
                ("hops", C.c_int64)]


class HostTag(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ncomps", C.c_int32), ("type", C.c_int32),
                ("nvalues", C.c_int64), ("data", C.c_void_p)]


class PicpartDim(C.Structure):
    _fields_ = [("num_entities", C.c_int64), ("nents", C.c_int32), ("num_cores", C.c_int32),
                ("buffered_parts", c_i32p), ("offset_ents_per_rank", c_i32p),
                ("ent_to_comm_arr_index", c_i32p), ("is_complete_part", c_i32p),
                ("num_bounds", C.c_int32), ("num_boundaries", C.c_int32),
                ("boundary_parts", c_i32p), ("offset_bounded", c_i32p),
                ("n_offset_bounded", C.c_int32), ("bounded_ent_ids", c_i32p),
                ("n_bounded_ent_ids", C.c_int32), ("ent_l2g", c_i32p)]


class MigrateStats(C.Structure):
    _fields_ = [("sent", C.c_int64), ("received", C.c_int64), ("deferred", C.c_int64)]


PP_INT32, PP_INT64, PP_FLOAT32, PP_FLOAT64 = 0, 1, 2, 3
PP_TAG_I8, PP_TAG_I32, PP_TAG_I64, PP_TAG_F64 = 0, 2, 3, 5
PP_SUM, PP_MAX, PP_MIN, PP_BCAST = 0, 1, 2, 3

# every symbol include/pumipic_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "pp_last_error": (C.c_char_p, []),
    "pp_version": (C.c_char_p, []),
    "pp_build_arch": (C.c_char_p, []),
    "pp_mesh_create": (C.c_int, [C.POINTER(MeshDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "pp_mesh_destroy": (C.c_int, [C.c_void_p]),
    "pp_mesh_get_info": (C.c_int, [C.c_void_p, C.POINTER(MeshInfo)]),
    "pp_mesh_get_arrays": (C.c_int, [C.c_void_p, C.POINTER(MeshArrays)]),
    "pp_mesh_set_picpart": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "pp_host_derive_sides": (C.c_int, [C.c_int32, C.c_int32, c_i32p, c_i32p, C.POINTER(c_i32p),
                                       C.POINTER(c_i32p)]),
    "pp_host_kuhn_cube": (C.c_int, [C.c_int32, C.c_double, c_i32p, C.POINTER(c_dp), c_i32p,
                                    C.POINTER(c_i32p)]),
    "pp_host_plate": (C.c_int, [C.c_int32, C.c_double, c_i32p, C.POINTER(c_dp), c_i32p,
                                C.POINTER(c_i32p)]),
    "pp_host_free": (None, [C.c_void_p]),
    "pp_host_picpart_tags": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       c_i32p, c_i32p]),
    "pp_host_picpart_tags_bridged": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               c_i32p, c_i32p]),
    "pp_host_entity_owners": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32,
                                        c_i32p]),
    "pp_ps_config_default": (None, [C.POINTER(PsConfig), C.c_int32]),
    "pp_ps_create": (C.c_int, [C.POINTER(PsConfig), C.c_int32, C.POINTER(MemberDesc), C.c_int32,
                               C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                               C.POINTER(C.c_void_p)]),
    "pp_ps_destroy": (C.c_int, [C.c_void_p]),
    "pp_ps_nelems": (C.c_int32, [C.c_void_p]),
    "pp_ps_nptcls": (C.c_int32, [C.c_void_p]),
    "pp_ps_capacity": (C.c_int32, [C.c_void_p]),
    "pp_ps_numrows": (C.c_int32, [C.c_void_p]),
    "pp_ps_kind_of": (C.c_int32, [C.c_void_p]),
    "pp_ps_member": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), c_i64p]),
    "pp_ps_get_layout": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PsLayout)]),
    "pp_ps_get_pids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_ps_rebuild": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                C.POINTER(C.c_void_p), C.c_void_p]),
    "pp_ps_set_rebuild_remap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "pp_ps_set_staged_rebuild": (None, [C.c_int32]),
    "pp_ps_set_rebuild_chunk_order": (None, [C.c_int32]),
    "pp_ps_set_rebuild_split_rows": (None, [C.c_int32]),
    "pp_ps_set_rebuild_block_histogram": (None, [C.c_int32]),
    "pp_ps_set_rebuild_tuning": (None, [C.c_int32, C.c_int32]),
    "pp_ps_set_shuffling": (None, [C.c_int32]),
    "pp_ps_set_rank_sort_threshold": (None, [C.c_int32]),
    "pp_push_constant": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "pp_push_direction": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                    C.c_void_p]),
    "pp_update_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_push_elliptical_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "pp_push_elliptical": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_void_p]),
    "pp_set_unsafe_procs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "pp_gyro_ring_map": (C.c_int, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_double,
                                   C.c_void_p, C.POINTER(SearchStats), C.c_void_p]),
    "pp_gyro_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_gyro_interleave": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_search_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs),
                                 C.POINTER(SearchStats), C.c_void_p]),
    "pp_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pp_comm_create": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "pp_comm_create_hosted": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.POINTER(C.c_void_p)]),
    "pp_comm_destroy": (C.c_int, [C.c_void_p]),
    "pp_timing_enable": (None, [C.c_int32]),
    "pp_timing_set_verbosity": (None, [C.c_int32]),
    "pp_timing_set_rank": (None, [C.c_int32]),
    "pp_timing_record": (None, [C.c_char_p, C.c_double]),
    "pp_timing_reset": (None, []),
    "pp_timing_count": (C.c_int32, []),
    "pp_timing_get": (C.c_int, [C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pp_timing_summarize": (None, [C.c_int32]),
    "pp_comm_set_p2p": (None, [C.c_int32]),
    "pp_comm_set_p2p_window": (C.c_int, [C.c_void_p, C.c_int64]),
    "pp_comm_p2p_active": (C.c_int32, [C.c_void_p]),
    "pp_comm_size": (C.c_int32, [C.c_void_p]),
    "pp_comm_rank": (C.c_int32, [C.c_void_p]),
    "pp_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "pp_comm_alltoall": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p]),
    "pp_comm_send": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pp_comm_recv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pp_comm_group_start": (C.c_int, []),
    "pp_comm_group_end": (C.c_int, []),
    "pp_comm_array_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_comm_plan_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.POINTER(C.c_void_p)]),
    "pp_comm_plan_destroy": (C.c_int, [C.c_void_p]),
    "pp_comm_plan_counts": (C.c_int, [C.c_void_p, c_i64p, c_i64p]),
    "pp_comm_plan_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "pp_host_picpart_extract": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_dp, c_i32p, c_i32p,
                                          C.c_int32, c_i32p, C.POINTER(C.c_int32),
                                          C.POINTER(c_i32p), C.POINTER(C.c_int32), C.POINTER(c_i32p),
                                          C.POINTER(c_i32p), C.POINTER(c_dp)]),
    "pp_ps_migrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(MigrateStats),
                                C.c_void_p]),
    "pp_trace_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs), C.c_void_p, C.c_void_p,
                                 c_i32p, C.c_void_p]),
    "pp_trace_find_exit_face": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_trace_check_model_intersection": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_trace_set_new_element": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_trace_pending": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SearchArgs), C.c_void_p, C.c_void_p,
                                   C.c_int32, c_i32p, C.c_void_p]),
    "pp_search_set_staged": (None, [C.c_int32]),
    "pp_search_set_l2_window": (None, [C.c_double]),
    "pp_search_last_stats": (C.c_int, [C.c_void_p, C.POINTER(SearchStats), C.c_void_p]),
    "pp_push_direction_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                           C.c_int32, C.POINTER(SearchArgs),
                                           C.POINTER(SearchStats), C.c_void_p]),
    "pp_push_direction_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                                C.c_int32, C.c_int32, C.POINTER(SearchStats),
                                                C.c_void_p]),
    "pp_push_boris": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_double, C.c_void_p]),
    "pp_gather_tet_field": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32),
                                      C.c_void_p]),
    "pp_gather_grid2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_gather_grid2d_vector": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                                          C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_gather_grid3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p]),
    "pp_host_mesh_read_osh": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "pp_host_mesh_write_osh": (C.c_int, [C.c_void_p, C.c_char_p]),
    "pp_host_mesh_from_elems": (C.c_int, [C.c_int32, C.c_int32, c_dp, C.c_int32, c_i32p,
                                          C.POINTER(C.c_void_p)]),
    "pp_host_mesh_destroy": (None, [C.c_void_p]),
    "pp_host_mesh_dim": (C.c_int32, [C.c_void_p]),
    "pp_host_mesh_nents": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pp_host_mesh_down": (c_i32p, [C.c_void_p, C.c_int32]),
    "pp_host_mesh_codes": (C.POINTER(C.c_int8), [C.c_void_p, C.c_int32]),
    "pp_host_mesh_ent2verts": (c_i32p, [C.c_void_p, C.c_int32]),
    "pp_host_mesh_coords": (c_dp, [C.c_void_p]),
    "pp_host_mesh_ntags": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pp_host_mesh_tag_at": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(HostTag)]),
    "pp_host_mesh_find_tag": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.POINTER(HostTag)]),
    "pp_host_mesh_set_tag": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                       C.c_void_p]),
    "pp_host_read_partition": (C.c_int, [C.c_char_p, C.c_int32, c_i32p, c_i32p]),
    "pp_host_picpart_build": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "pp_host_picpart_build_bridged": (C.c_int, [C.c_void_p, c_i32p, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                C.POINTER(C.c_void_p)]),
    "pp_host_picpart_destroy": (None, [C.c_void_p]),
    "pp_host_picpart_mesh": (C.c_void_p, [C.c_void_p]),
    "pp_host_picpart_get": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(PicpartDim)]),
    "pp_host_picpart_is_full_mesh": (C.c_int32, [C.c_void_p]),
    "pp_host_picpart_nranks": (C.c_int32, [C.c_void_p]),
    "pp_host_picpart_rank": (C.c_int32, [C.c_void_p]),
    "pp_host_picpart_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "pp_host_picpart_read": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "pp_host_ppm_set_compression": (None, [C.c_int32]),
    "pp_host_picpart_sbars": (C.c_int, [C.c_void_p, c_i32p, C.POINTER(c_i32p), C.POINTER(c_i32p),
                                        C.POINTER(c_i32p), c_i32p]),
    "pp_push_from": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_double, C.c_void_p]),
    "pp_host_lb_plan": (C.c_int, [C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, C.c_int32, c_dp, c_dp,
                                  C.c_double, C.c_double, c_i32p, C.POINTER(c_i32p),
                                  C.POINTER(c_i32p), C.POINTER(c_dp), c_dp]),
    "pp_balancer_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_void_p)]),
    "pp_balancer_destroy": (C.c_int, [C.c_void_p]),
    "pp_balancer_info": (C.c_int, [C.c_void_p, c_i32p, c_i32p, C.POINTER(c_i32p), C.POINTER(c_i32p)]),
    "pp_balancer_add_weights_ps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_balancer_add_weights_array": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_balancer_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_i64p]),
    "pp_balancer_balance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    "pp_balancer_plan": (C.c_int, [C.c_void_p, c_i32p, C.POINTER(c_i32p), C.POINTER(c_i32p),
                                   C.POINTER(c_dp), c_dp]),
    "pp_balancer_select_ps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_balancer_select_array": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pp_balancer_repartition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                          C.c_void_p, C.c_double, C.c_void_p]),
    "pp_balancer_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double,
                                        C.c_double, C.c_void_p, C.c_void_p]),
}

_lib = None


class PumipicError(RuntimeError):
    pass


def lib():
    """Load libpumipic_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PumipicError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != PP_OK:
        raise PumipicError("pumipic_b200 status %d: %s" % (status, lib().pp_last_error().decode()))
