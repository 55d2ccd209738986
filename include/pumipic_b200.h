/*
 * pumipic_b200.h -- C ABI of libpumipic_b200.so: the B200-native (sm_100a CUDA + NCCL)
 * implementation of PUMI-PIC's per-timestep particle hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / opaque handles and
 * returns a pp_status; pp_last_error() gives the message of the last failure on the calling
 * thread.  Unless a parameter says "host", array pointers are DEVICE pointers (the reference's
 * Kokkos views and Omega_h arrays are device-resident in a CUDA build, and a binding passes
 * `view.data()`).  Each entry cites the reference interface it replaces
 * (paths relative to SCOREC/pumi-pic @ c09ad045).  The reference-side binding is shown in
 * INTEGRATION.md; the header-only C++ mirror of the reference API lives in
 * pumi-pic_b200/cpp/.
 *
 * Layout conventions (identical to the reference):
 *   - a particle member with N components is component-major SoA ("LayoutLeft"):
 *     value(slot, i) lives at base[i * stride + slot]
 *   - per-slot user arrays (elem_ids, inter_faces, ...) have capacity() entries, indexed by slot
 *   - inter_points is AoS: inter_points[dim * slot + i]
 *   - lid_t = int32, gid_t = int64, fp_t = double
 *
 * Handles are not thread-safe; all work of a call is enqueued on the `stream` argument
 * (a cudaStream_t; NULL = the legacy default stream).  Calls that return values to the host
 * synchronise that stream, the others are asynchronous.
 */
#ifndef PUMIPIC_B200_H
#define PUMIPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pp_mesh pp_mesh;
typedef struct pp_ps pp_ps;
typedef struct pp_comm pp_comm;
typedef void* pp_stream; /* cudaStream_t */

typedef enum pp_status {
  PP_OK = 0,
  PP_ERR_INVALID = 1,     /* bad argument */
  PP_ERR_CUDA = 2,        /* CUDA runtime failure (no device, launch error, ...) */
  PP_ERR_NCCL = 3,
  PP_ERR_NOMEM = 4,
  PP_ERR_UNSUPPORTED = 5
} pp_status;

typedef enum pp_memspace { PP_HOST = 0, PP_DEVICE = 1 } pp_memspace;
typedef enum pp_dtype { PP_INT32 = 0, PP_INT64 = 1, PP_FLOAT32 = 2, PP_FLOAT64 = 3 } pp_dtype;
/* Mesh::Op (pumipic_mesh.hpp:57-62) */
typedef enum pp_op { PP_SUM = 0, PP_MAX = 1, PP_MIN = 2, PP_BCAST = 3 } pp_op;

const char* pp_last_error(void);
/* library version "major.minor.patch" and the GPU architecture it was built for ("sm_100a") */
const char* pp_version(void);
const char* pp_build_arch(void);

/* ============================== mesh ==================================================== */

/* The arrays an Omega_h mesh hands to the reference's searches:
 *   coords      = mesh.coords()                 (adjacency.tpp:79,238-241)
 *   elem2verts  = mesh.ask_elem_verts()         (adjacency.tpp:78)
 *   elem2sides  = mesh.ask_down(dim,dim-1).ab2b (adjacency.tpp:240-241)
 *   side2verts  = mesh.ask_verts_of(dim-1)      (adjacency.tpp:282,501)
 *   elem_class  = mesh.get_array<ClassId>(dim,"class_id") (ellipticalPush.hpp:43), optional
 * Everything else the reference recomputes per search call (measure_elements_real,
 * mark_exposed_sides, ask_up(dim-1,dim), ask_dual, compute_tolerance_from_area;
 * adjacency.tpp:489-501,621) is derived ONCE here, on the device, and cached in the handle. */
typedef struct pp_mesh_desc {
  int32_t dim;       /* 2 (triangles) or 3 (tets) */
  int32_t nverts, nelems, nsides;
  const double* coords;       /* [nverts*dim]         */
  const int32_t* elem2verts;  /* [nelems*(dim+1)]     */
  const int32_t* elem2sides;  /* [nelems*(dim+1)]     */
  const int32_t* side2verts;  /* [nsides*dim]         */
  const int32_t* elem_class;  /* [nelems] or NULL     */
  int32_t memspace;           /* pp_memspace of the five arrays above */
} pp_mesh_desc;

pp_status pp_mesh_create(const pp_mesh_desc* desc, pp_stream stream, pp_mesh** out);
pp_status pp_mesh_destroy(pp_mesh* mesh);

typedef struct pp_mesh_info {
  int32_t dim, nverts, nelems, nsides;
  double tol;               /* compute_tolerance_from_area, adjacency.tpp:419-428 */
  double min_measure;       /* min element area / volume */
  int32_t n_exposed_sides;
  int64_t walk_table_bytes; /* size of the packed per-element walk records in HBM */
} pp_mesh_info;
pp_status pp_mesh_get_info(const pp_mesh* mesh, pp_mesh_info* out);

/* Device pointers to the derived arrays (valid until pp_mesh_destroy):
 *   measure      double[nelems]   == measure_elements_real(&mesh)
 *   exposed      int8[nsides]     == mark_exposed_sides(&mesh)
 *   side2elem    int32[2*nsides]  == ask_up(dim-1,dim): (lower id, higher id or -1) per side
 *   dual_off/dual                 == ask_dual().a2ab / .ab2b                                  */
typedef struct pp_mesh_arrays {
  const double* coords;
  const int32_t* elem2verts;
  const int32_t* elem2sides;
  const int32_t* side2verts;
  const int32_t* elem_class;
  const double* measure;
  const int8_t* exposed;
  const int32_t* side2elem;
  const int32_t* dual_off;
  const int32_t* dual;
} pp_mesh_arrays;
pp_status pp_mesh_get_arrays(const pp_mesh* mesh, pp_mesh_arrays* out);

/* PICpart tags used by setUnsafeProcs (pumipic_ptcl_ops.hpp:33-53): safe = Mesh::safeTag(),
 * owner = Mesh::entOwners(dim).  Copied into the walk records so the fused search tail needs
 * no extra gather.  memspace as above. */
pp_status pp_mesh_set_picpart(pp_mesh* mesh, const int32_t* safe, const int32_t* owner,
                              int32_t self_rank, int32_t memspace, pp_stream stream);

/* ---- host-side mesh utilities (no GPU needed; caller owns the outputs, free with pp_host_free)
 * Derive sides of a simplicial mesh from element->vertex connectivity.  Sides are numbered by
 * the lexicographic order of their sorted vertex tuple; a side's own vertex order is the
 * template order seen from its lowest-numbered adjacent element. */
pp_status pp_host_derive_sides(int32_t dim, int32_t nelems, const int32_t* elem2verts,
                               int32_t* nsides_out, int32_t** elem2sides_out,
                               int32_t** side2verts_out);
/* n^3 cubes on [0,length]^3, six positively oriented tets per cube (Kuhn split). */
pp_status pp_host_kuhn_cube(int32_t n, double length, int32_t* nverts_out, double** coords_out,
                            int32_t* nelems_out, int32_t** elem2verts_out);
/* n^2 squares on [0,length]^2, two counter-clockwise triangles per square. */
pp_status pp_host_plate(int32_t n, double length, int32_t* nverts_out, double** coords_out,
                        int32_t* nelems_out, int32_t** elem2verts_out);
void pp_host_free(void* p);
/* PICpart tags of rank `rank` for an element->owner partition (pumipic::Input semantics,
 * src/pumipic_part_construct.cpp:73-114,409-468; methods: 0 FULL, 1 BFS, 2 MINIMUM, 3 NONE as
 * pumipic_input.hpp:33-39; defaults buffer_layers 3, safe_layers 1, vertex-bridged BFS).
 * safe_out[nelems] = Mesh::safeTag(), has_part_out[nranks] = which cores are buffered here. */
pp_status pp_host_picpart_tags(int32_t dim, int32_t nverts, int32_t nelems,
                               const int32_t* elem2verts, const int32_t* owner, int32_t nranks,
                               int32_t rank, int32_t buffer_method, int32_t safe_method,
                               int32_t buffer_layers, int32_t safe_layers, int32_t* safe_out,
                               int32_t* has_part_out);
/* The same with the BFS bridged through any entity dimension (pumipic::Input::bridge_dim,
 * pumipic_input.hpp; part_construct.cpp:424 ask_up(bridge_dim, dim)): the caller hands
 * elem2bridges[nelems*bridges_per_elem] = the elements' entities of that dimension (vertices,
 * edges or sides; nbridges of them in the mesh). */
pp_status pp_host_picpart_tags_bridged(int32_t nbridges, int32_t nelems, int32_t bridges_per_elem,
                                       const int32_t* elem2bridges, const int32_t* owner,
                                       int32_t nranks, int32_t rank, int32_t buffer_method,
                                       int32_t safe_method, int32_t buffer_layers,
                                       int32_t safe_layers, int32_t* safe_out,
                                       int32_t* has_part_out);
/* The sub-mesh of a partially buffered PICpart (constructPICPart, part_construct.cpp:116-262):
 * elements whose owner's core is buffered here (has_part[owner] != 0, from pp_host_picpart_tags)
 * and their vertices, both keeping their relative order in the full mesh (:182-195).
 *   elem_l2g[nelems_l], vert_l2g[nverts_l]  full-mesh index of every local element / vertex
 *   elem2verts_l[nelems_l*(dim+1)]          connectivity in local vertex numbers
 *   coords_l[nverts_l*dim]
 * Outputs are malloc'd; free with pp_host_free.  Tags (owner, safe, class) follow by indexing
 * the full-mesh arrays with elem_l2g / vert_l2g (convertTag, :597-617). */
pp_status pp_host_picpart_extract(int32_t dim, int32_t nverts, int32_t nelems, const double* coords,
                                  const int32_t* elem2verts, const int32_t* owner, int32_t nranks,
                                  const int32_t* has_part, int32_t* nelems_out,
                                  int32_t** elem_l2g_out, int32_t* nverts_out,
                                  int32_t** vert_l2g_out, int32_t** elem2verts_out,
                                  double** coords_out);
/* Owner of lower-dimensional entities = minimum owner of the adjacent elements
 * (defineOwners, part_construct.cpp:304-323).  elem2ents: [nelems*ents_per_elem]. */
pp_status pp_host_entity_owners(int32_t nents, int32_t nelems, int32_t ents_per_elem,
                                const int32_t* elem2ents, const int32_t* elem_owner,
                                int32_t nranks, int32_t* ent_owner_out);

/* ---- host mesh with every entity dimension, and the Omega_h `.osh` format ----
 * What the reference holds in an Omega_h::Mesh on the set-up side: vertices, edges, faces,
 * elements, the d -> d-1 adjacency with alignment codes, and named tags per dimension.
 * Entities are numbered exactly as in the file (Omega_h's numbering). */
typedef struct pp_host_mesh pp_host_mesh;
typedef enum pp_host_tag_type { PP_TAG_I8 = 0, PP_TAG_I32 = 2, PP_TAG_I64 = 3, PP_TAG_F64 = 5 } pp_host_tag_type;
typedef struct pp_host_tag {
  const char* name;   /* owned by the mesh */
  int32_t ncomps;
  int32_t type;       /* pp_host_tag_type (Omega_h's type codes) */
  int64_t nvalues;    /* nents * ncomps */
  const void* data;   /* owned by the mesh */
} pp_host_tag;
/* Omega_h::binary::read(path, comm) of a serial `.osh` directory (test/test_file.cpp:24,
 * src/pumipic_file.cpp:135). */
pp_status pp_host_mesh_read_osh(const char* path, pp_host_mesh** out);
/* Omega_h::binary::write(path, mesh) (src/pumipic_file.cpp:69). */
pp_status pp_host_mesh_write_osh(const pp_host_mesh* mesh, const char* path);
/* Omega_h::build_from_elems2verts: faces / edges derived from the elements (numbered by sorted
 * vertex tuple, vertex order of the last use), "global" and "coordinates" tags added. */
pp_status pp_host_mesh_from_elems(int32_t dim, int32_t nverts, const double* coords,
                                  int32_t nelems, const int32_t* elem2verts, pp_host_mesh** out);
void pp_host_mesh_destroy(pp_host_mesh* mesh);
int32_t pp_host_mesh_dim(const pp_host_mesh* mesh);
int32_t pp_host_mesh_nents(const pp_host_mesh* mesh, int32_t d);          /* Mesh::nents(d) */
const int32_t* pp_host_mesh_down(const pp_host_mesh* mesh, int32_t d);    /* ask_down(d,d-1).ab2b */
const int8_t* pp_host_mesh_codes(const pp_host_mesh* mesh, int32_t d);    /* ask_down(d,d-1).codes */
const int32_t* pp_host_mesh_ent2verts(const pp_host_mesh* mesh, int32_t d); /* ask_verts_of(d) */
const double* pp_host_mesh_coords(const pp_host_mesh* mesh);              /* Mesh::coords() */
int32_t pp_host_mesh_ntags(const pp_host_mesh* mesh, int32_t d);
pp_status pp_host_mesh_tag_at(const pp_host_mesh* mesh, int32_t d, int32_t i, pp_host_tag* out);
pp_status pp_host_mesh_find_tag(const pp_host_mesh* mesh, int32_t d, const char* name,
                                pp_host_tag* out);
/* Mesh::add_tag / set_tag; data holds nents(d) * ncomps values and is copied. */
pp_status pp_host_mesh_set_tag(pp_host_mesh* mesh, int32_t d, const char* name, int32_t ncomps,
                               int32_t type, const void* data);
/* Partition files of pumipic::Input (src/pumipic_input.cpp:44-89): `.ptn` (owner per element)
 * or `.cpn` (owner per class id; elem_class = the elements' class_id tag, else NULL). */
pp_status pp_host_read_partition(const char* path, int32_t nelems, const int32_t* elem_class,
                                 int32_t* owner_out);

/* ---- PICpart construction and `.ppm` files --------------------------------
 * pumipic::Mesh(Input&) (src/pumipic_part_construct.cpp:73-262) + Mesh::setupComm
 * (src/pumipic_comm.cpp:12-184) + the safe-zone overlap regions of ParticleBalancer
 * (src/pumipic_lb.cpp:23-82) for rank `rank` of `nranks`.  Like the reference, every rank starts
 * from the full mesh loaded in serial; unlike it, nothing is exchanged: what the peers would
 * send (boundary entity lists, safe flags, sbar tables) is a function of the same full mesh and
 * partition and is evaluated locally.  Entities owned by a part of which only a boundary is
 * held get their rank-local ids in ascending entity order (the reference takes them from
 * atomics, any order being valid, pumipic_comm.cpp:66-75).
 * Methods: 0 FULL, 1 BFS, 2 MINIMUM, 3 NONE (pumipic_input.hpp:33-39); layers < 0 = the defaults
 * (buffer 3, safe 1, pumipic_input.cpp:103-110). */
typedef struct pp_host_picpart pp_host_picpart;
pp_status pp_host_picpart_build(const pp_host_mesh* full, const int32_t* elem_owner,
                                int32_t nranks, int32_t rank, int32_t buffer_method,
                                int32_t safe_method, int32_t buffer_layers, int32_t safe_layers,
                                pp_host_picpart** out);
/* The same with Input::bridge_dim (0 <= bridge_dim < dim; 0 = vertices is Input's default and what
 * pp_host_picpart_build uses; test/test_revClass.cpp sets dim-1). */
pp_status pp_host_picpart_build_bridged(const pp_host_mesh* full, const int32_t* elem_owner,
                                        int32_t nranks, int32_t rank, int32_t buffer_method,
                                        int32_t safe_method, int32_t buffer_layers,
                                        int32_t safe_layers, int32_t bridge_dim,
                                        pp_host_picpart** out);
void pp_host_picpart_destroy(pp_host_picpart* pp);
/* The PICpart's own mesh (Mesh::mesh()), with the tags the reference puts on it: "ownership"
 * (entOwners), "gids" (globalIds), "rank_lids" (rankLocalIndex), "global" (index in the full
 * mesh) on every dimension, "safe" (safeTag) and "sbar_id" (ParticleBalancer::getSbarIDs) on
 * elements.  Owned by the PICpart. */
const pp_host_mesh* pp_host_picpart_mesh(const pp_host_picpart* pp);
typedef struct pp_host_picpart_dim {
  int64_t num_entities;                 /* Mesh::nents of the full mesh                      */
  int32_t nents;                        /* entities of this dimension in the PICpart         */
  int32_t num_cores;                    /* numBuffers(d) - 1                                 */
  const int32_t* buffered_parts;        /* bufferedRanks(d) [num_cores]                      */
  const int32_t* offset_ents_per_rank;  /* nentsOffsets(d) [nranks+1]                        */
  const int32_t* ent_to_comm_arr_index; /* commArrayIndex(d) [nents]                         */
  const int32_t* is_complete_part;      /* [nranks]: 0 absent, 1 boundary only, 2 complete   */
  int32_t num_bounds, num_boundaries;
  const int32_t* boundary_parts;        /* [num_boundaries]                                  */
  const int32_t* offset_bounded;        /* [nranks+1] or empty (element dimension)           */
  int32_t n_offset_bounded;
  const int32_t* bounded_ent_ids;       /* [n_bounded_ent_ids]                               */
  int32_t n_bounded_ent_ids;
  const int32_t* ent_l2g;               /* full-mesh index per local entity, NULL after a read */
} pp_host_picpart_dim;
pp_status pp_host_picpart_get(const pp_host_picpart* pp, int32_t d, pp_host_picpart_dim* out);
int32_t pp_host_picpart_is_full_mesh(const pp_host_picpart* pp);   /* Mesh::isFullMesh() */
int32_t pp_host_picpart_nranks(const pp_host_picpart* pp);
int32_t pp_host_picpart_rank(const pp_host_picpart* pp);
/* pumipic::write(picparts, prefix) (src/pumipic_file.cpp:45-116): <prefix>_<nranks>.ppm/
 * <name>_<rank>.osh + <name>_<rank>.ppm (format version 2). */
pp_status pp_host_picpart_write(const pp_host_picpart* pp, const char* prefix);
/* pumipic::read(lib, comm, prefix, &mesh) (src/pumipic_file.cpp:118-205), versions 1 and 2. */
pp_status pp_host_picpart_read(const char* prefix, int32_t nranks, int32_t rank,
                               pp_host_picpart** out);
/* The reference compresses the .ppm arrays only when Omega_h was built with zlib (OMEGA_H_USE_ZLIB,
 * src/pumipic_file.cpp:76-80) and the file carries no flag.  on = 1 (default) writes compressed
 * arrays, 0 raw ones (also PUMIPIC_PPM_ZLIB=0); the reader tries the configured form first and the
 * other one when the arrays do not decode, so files of either build are read. */
void pp_host_ppm_set_compression(int32_t on);
/* The safe-zone overlap regions ("sbars") this part belongs to (ParticleBalancer,
 * src/pumipic_lb.hpp:97-101): global sbar id and the sorted parts sharing it.  parts of sbar i
 * are parts[off[i] .. off[i+1]).  Pointers are owned by the PICpart. */
pp_status pp_host_picpart_sbars(const pp_host_picpart* pp, int32_t* nsbars,
                                const int32_t** sbar_ids, const int32_t** parts_off,
                                const int32_t** parts, int32_t* max_sbar);

/* ============================== particle structure ====================================== */

typedef enum pp_ps_kind {
  PP_PS_SCS = 0,  /* particle_structs/src/scs/SellCSigma.h  */
  PP_PS_CSR = 1,  /* particle_structs/src/csr/CSR.hpp       */
  PP_PS_CABM = 2, /* particle_structs/src/cabm/cabm.hpp (API-compatible, SCS storage engine) */
  PP_PS_DPS = 3   /* particle_structs/src/dps/dps.hpp  (flat: parent array + mask)           */
} pp_ps_kind;

typedef enum pp_padding { PP_PAD_EVENLY = 0, PP_PAD_PROPORTIONALLY = 1, PP_PAD_INVERSELY = 2 } pp_padding;

/* One entry per member of MemberTypes<...> (support/MemberTypes.h:21-60): scalar size in bytes
 * (BaseType<T>::type) and number of scalars (BaseType<T>::size). */
typedef struct pp_member_desc {
  int32_t scalar_bytes;
  int32_t ncomp;
} pp_member_desc;

/* Mirrors SCS_Input (scs/scs_input.hpp:4-36) and the constructor arguments
 * (SellCSigma.h:66-71, CSR.hpp:37-44, dps.hpp:41-48). */
typedef struct pp_ps_config {
  int32_t kind;            /* pp_ps_kind */
  int32_t team_size;       /* policy.team_size(): maximum chunk height C (SCS) */
  int32_t sigma;           /* sorting window; INT_MAX = full sort */
  int32_t V;               /* vertical slice width */
  double shuffle_padding;  /* default 0.1  */
  double extra_padding;    /* default 0.05 */
  double minimize_size;    /* default 0.8  */
  int32_t padding_strat;   /* pp_padding, default PP_PAD_EVENLY */
  int32_t always_realloc;  /* default 0 */
} pp_ps_config;
void pp_ps_config_default(pp_ps_config* cfg, int32_t kind);

/* Build a structure with `ne` elements and `np` particles.
 *   ppe[ne]                particles per element
 *   elem_gids[ne]          element global ids, or NULL
 *   particle_elements[np]  parent element of each initial particle, or NULL
 *   particle_info[nmembers] one array per member, each [ncomp][np] component-major, or NULL
 * memspace tells where ppe / elem_gids / particle_elements / particle_info live. */
pp_status pp_ps_create(const pp_ps_config* cfg, int32_t nmembers, const pp_member_desc* members,
                       int32_t ne, int32_t np, const int32_t* ppe, const int64_t* elem_gids,
                       const int32_t* particle_elements, const void* const* particle_info,
                       int32_t memspace, pp_stream stream, pp_ps** out);
pp_status pp_ps_destroy(pp_ps* ps);

/* particle_structure.hpp:71-75 nElems / nPtcls / capacity / numRows */
int32_t pp_ps_nelems(const pp_ps* ps);
int32_t pp_ps_nptcls(const pp_ps* ps);
int32_t pp_ps_capacity(const pp_ps* ps);
int32_t pp_ps_numrows(const pp_ps* ps);
int32_t pp_ps_kind_of(const pp_ps* ps);

/* get<N>() (particle_structure.hpp:83-97): device base pointer and slot stride of member N.
 * Invalidated by rebuild / migrate, exactly like the reference's Segment handles. */
pp_status pp_ps_member(const pp_ps* ps, int32_t member, void** base_out, int64_t* stride_out);

/* Slot geometry for user kernels (what SellCSigma::parallel_for iterates, SellCSigma.h:528-558).
 * mask_bits: bit (slot & 31) of word (slot >> 5) is the particle_mask of the slot.
 * slot_elem: int32[capacity] element of the row that owns each slot (materialised on demand). */
typedef struct pp_ps_layout {
  int32_t kind, C, V, nchunks, nslices, nrows, capacity, nelems, nptcls;
  const int32_t* offsets;        /* [nslices+1]  (SCS) / [nelems+1] (CSR) */
  const int32_t* slice_to_chunk; /* [nslices]    (SCS) */
  const int32_t* row_to_element; /* [nrows]      (SCS) */
  const int32_t* element_to_row; /* [nrows]      (SCS) */
  const uint32_t* mask_bits;     /* [(capacity+31)/32] */
  const int32_t* slot_elem;      /* [capacity] */
} pp_ps_layout;
pp_status pp_ps_get_layout(pp_ps* ps, pp_stream stream, pp_ps_layout* out);
/* ParticleStructure::getPIDs (particle_structs/src/ps_for.hpp:57-88): pids[nptcls] = the slots of
 * the masked particles grouped by element (ascending slot inside a group), offsets[nelems+1] = the
 * start of each element's group; both device int32. */
pp_status pp_ps_get_pids(pp_ps* ps, int32_t* pids, int32_t* offsets, pp_stream stream);

/* rebuild (particle_structure.hpp:99-100; SCS_rebuild.h:123, CSR_rebuild.hpp:18, dps_rebuild.hpp):
 *   new_element[capacity]       new parent element per slot, -1 deletes the particle
 *   new_particle_elements[n_new], new_particle_info[nmembers] (each [ncomp][n_new]) particles to add
 * All pointers are device pointers (new_particle_info itself is a host array of device pointers).
 * Slot numbering, capacity and member pointers change; re-fetch them afterwards.  Returns
 * PP_ERR_INVALID if a new particle has element -1 (the reference exits the process). */
pp_status pp_ps_rebuild(pp_ps* ps, const int32_t* new_element, int32_t n_new,
                        const int32_t* new_particle_elements,
                        const void* const* new_particle_info, pp_stream stream);

/* One-shot member remap of the NEXT pp_ps_rebuild / pp_ps_migrate of this structure: in the rebuilt
 * structure member i holds what member src_member[i] held before (-1: zeros); src_member[i] must have
 * member i's type and feed at most one destination; n = number of members (0 clears).  Particles
 * received or added by that call are remapped the same way (they travel in the old member order).
 * This folds the reference drivers' updatePtclPositions (x <- xtgt, xtgt <- 0,
 * test/pseudoPushAndSearch.cpp:142-154) into the record move the rebuild does anyway: one pass over the
 * particles less, and the zero-filled members are not read at all.  Same result for every particle
 * that survives the rebuild. */
pp_status pp_ps_set_rebuild_remap(pp_ps* ps, const int32_t* src_member, int32_t n);
/* Record move of a full re-layout.  2 (default): device-side layout with one host read per rebuild;
 * narrow structures (<= 14 columns per chunk on average) then gather the records in one pass in
 * destination order, the destination chunks handed out in element order so that gathered source
 * sectors stay in L2, wider ones go through the record stage (Sell-C-sigma with C = 32 and
 * sparse rows; other cases use mode 1).  1: through an array-of-records stage behind the host-synchronous
 * layout code (every structure kind; does not depend on the locality of the element numbering).
 * 0: direct scatter (A/B).  Same results as sets of particles in every mode. */
void pp_ps_set_staged_rebuild(int32_t mode);
/* Mode 2: hand destination chunks to the blocks in ascending order of their first row's element
 * (default) or in slot order (0, A/B: the gather then re-fetches every source sector from DRAM). */
void pp_ps_set_rebuild_chunk_order(int32_t on);
/* Layout of a mostly empty structure (one sort window, fewer than 40 % of the rows non-empty after the
 * previous rebuild -- a PICpart that buffers the whole mesh): 1 (default) sorts only the non-empty rows and
 * places the empty ones by a prefix sum, 0 sorts all rows.  Same layout either way; A/B switch. */
void pp_ps_set_rebuild_split_rows(int32_t on);
/* Destination histogram of the rebuild (the particle's rank in its new element): 1 (default) counts a block's
 * particles per destination in shared memory and reserves their ranks with one global atomic per
 * destination, 0 uses one global atomic per particle.  Same ranks as sets; A/B switch. */
void pp_ps_set_rebuild_block_histogram(int32_t on);
/* Mode 2 A/B knobs.  gather_blocks_per_sm: resident blocks per SM of the gather; <= 0 (default) sizes
 * the grid so that the source footprint of the chunks in flight fits L2.  gather_max_cols: average
 * columns per chunk up to which the records are gathered; wider structures go through the record
 * stage (default 14; 0 = always stage, < 0 keeps the setting). */
void pp_ps_set_rebuild_tuning(int32_t gather_blocks_per_sm, int32_t gather_max_cols);
/* SellCSigma::setShuffling (scs/SellCSigma.h:92): try the in-place reshuffle (SCS_rebuild.h:4-120)
 * before a full re-layout.  Default on. */
void pp_ps_set_shuffling(int32_t on);
/* Average particles per element from which a rebuild derives counts and in-row ranks from a sort
 * of the particles by destination element instead of per-element atomics (default 128). */
void pp_ps_set_rank_sort_threshold(int32_t particles_per_element);

/* ============================== push ===================================================== */

/* test/pseudoPushAndSearch.cpp:87-118: xtgt = x + distance*(dx,dy,dz) for masked slots */
pp_status pp_push_constant(pp_ps* ps, const double* x, double* xtgt, int64_t stride,
                           double distance, double dx, double dy, double dz, pp_stream stream);
/* test/test_adj.cpp:550-562: tgt += distance*dir for masked slots */
pp_status pp_push_direction(pp_ps* ps, double* tgt, const double* dir, int64_t stride,
                            double distance, pp_stream stream);
/* test/pseudoPushAndSearch.cpp:142-154 updatePtclPositions: x = xtgt; xtgt = 0 (all slots) */
pp_status pp_update_positions(pp_ps* ps, double* x, double* xtgt, int64_t stride,
                              pp_stream stream);

/* test/ellipticalPush.hpp:10-34 setup: phi = atan2(d(z-k), w-h), b = (z-k)/sin(phi), stored as
 * float like the reference's particle members; :36-70 push: the particle advances `deg` degrees
 * (scaled by 1/class_id of its row element, 0.01x for class 1) along its ellipse. */
pp_status pp_push_elliptical_setup(pp_ps* ps, const double* x, int64_t stride, float* b,
                                   float* phi, double h, double k, double d, pp_stream stream);
pp_status pp_push_elliptical(pp_mesh* mesh, pp_ps* ps, double* xtgt, int64_t stride,
                             const float* b, float* phi, double h, double k, double d, double deg,
                             pp_stream stream);
/* src/pumipic_push.hpp:17-74 pushBoris on flat component-major arrays [3][stride] of n particles:
 * v- = v - q'E, v' = v- + q'(v- x B), v = v- + c (v' x B) + q'E, x = x_prev + v dt, x_prev = x
 * (charge 1, amu 10 as in the reference). */
pp_status pp_push_boris(int64_t n, int64_t stride, double* pos, double* pos_prev, double* vel,
                        const double* efield, const double* bfield, double dt, pp_stream stream);
/* src/pumipic_ptcl_ops.hpp:33-53 setUnsafeProcs: new_elems = elems, new_procs = owner of the
 * element when it is not safe on this PICpart (tags from pp_mesh_set_picpart), else this rank. */
pp_status pp_set_unsafe_procs(pp_mesh* mesh, pp_ps* ps, const int32_t* elems, int32_t* new_elems,
                              int32_t* new_procs, pp_stream stream);

/* ============================== search =================================================== */

typedef enum pp_search_variant {
  PP_SEARCH_NEW = 0,       /* adjacency.tpp:642 search_mesh (BCC or ray intersection, 2D/3D) */
  PP_SEARCH_2D_LEGACY = 1, /* adjacency.hpp:1013 search_mesh_2d */
  PP_SEARCH_3D_LEGACY = 2, /* adjacency.hpp:559 search_mesh (line-triangle + dual graph) */
  PP_SEARCH_3D = 3         /* adjacency.hpp:316 search_mesh_3d (barycentric_coords_tet, tol 1e-20,
                            * checkCurrentElm / findIntersection / processUndetected) */
} pp_search_variant;

typedef struct pp_search_args {
  int32_t variant;               /* pp_search_variant */
  const double* x_orig;          /* [3][stride] start positions  (x_ps_orig) */
  const double* x_tgt;           /* [3][stride] target positions (x_ps_tgt)  */
  int64_t stride;
  int32_t* elem_ids;             /* [capacity] in/out parent element per slot */
  int32_t elem_ids_empty;        /* !=0: behave as if elem_ids.size()==0: every particle starts in its row
                                  * element and the array's old contents are never read (all variants,
                                  * all structure kinds); every slot of elem_ids is written */
  int32_t require_intersection;  /* PP_SEARCH_NEW only */
  int32_t* inter_faces;          /* [capacity] or NULL (required when require_intersection / legacy 3D and search_mesh_3d xface) */
  double* inter_points;          /* [dim*capacity] AoS or NULL (legacy 3D, search_mesh_3d: xpoints [3*capacity]) */
  int32_t looplimit;             /* 0 = unlimited */
} pp_search_args;

typedef struct pp_search_stats {
  int32_t found;        /* the reference's bool return value */
  int32_t loops;        /* walk iterations the reference would have executed */
  int32_t not_in_elem;  /* deleted by check_initial_parents (adjacency.tpp:73-145) */
  int32_t not_found;    /* deleted by the loop limit */
  int32_t aborted;      /* legacy 3D, search_mesh_3d: particles whose origin is outside their element (reference aborts) */
  int32_t active;       /* particles that entered the walk */
  int64_t hops;         /* total element hops */
} pp_search_stats;

/* Replaces search_mesh / search_mesh_2d / legacy search_mesh: ONE fused kernel (setup, origin
 * check, walk to completion, boundary handling), no per-iteration host round trip.
 * stats_host may be NULL (fully asynchronous); otherwise the stream is synchronised and the
 * counters copied out. */
pp_status pp_search_mesh(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                         pp_search_stats* stats_host, pp_stream stream);
/* The phases of trace_particle_through_mesh (adjacency.tpp:461-640) as separate calls, for
 * applications that pass their own handler `Func` (called on the host once per walk iteration,
 * between find_exit_face and set_new_element, with the device arrays elem_ids, inter_faces,
 * lastExit, inter_points and ptcl_done).  args as for pp_search_mesh (variant PP_SEARCH_NEW);
 * ptcl_done and last_exit are caller-owned device int32[capacity].  The loop of the reference:
 *   pp_trace_begin                      setInitial, finishUnmoved, initializeIntersection,
 *                                       check_initial_parents (:484-552); *not_in_elem = particles
 *                                       deleted because their origin is outside their element
 *   repeat { pp_trace_find_exit_face    (:232-364; BCC when !require_intersection, else ray / edge)
 *            handler                    stock: pp_trace_check_model_intersection (:366-387)
 *            pp_trace_set_new_element   (:390-416)
 *            pp_trace_pending(0)        (:568-573) *count == 0 <=> every slot is done
 *   } until done or the loop limit, then pp_trace_pending(1) removes what is left (:584-606).
 * With the stock handler the arrays end identical to pp_search_mesh's, which does the same in
 * one kernel and is the path to use whenever the handler is the stock one. */
pp_status pp_trace_begin(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args, int32_t* ptcl_done,
                         int32_t* last_exit, int32_t* not_in_elem, pp_stream stream);
pp_status pp_trace_find_exit_face(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                  int32_t* ptcl_done, int32_t* last_exit, pp_stream stream);
pp_status pp_trace_check_model_intersection(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                            int32_t* ptcl_done, int32_t* last_exit,
                                            pp_stream stream);
pp_status pp_trace_set_new_element(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args,
                                   int32_t* ptcl_done, int32_t* last_exit, pp_stream stream);
pp_status pp_trace_pending(pp_mesh* mesh, pp_ps* ps, const pp_search_args* args, int32_t* ptcl_done,
                           int32_t* last_exit, int32_t remove_pending, int32_t* count,
                           pp_stream stream);
/* Kernel selection for the barycentric walks: 2 (default) = Sell-C-sigma chunk walk where it
 * applies (C = 32, elem_ids seeded from the rows), 1 = block-staged kernel, 0 = the simple
 * thread-per-slot kernel.  All give identical results; the switch exists for A/B measurements. */
void pp_search_set_staged(int32_t on);
/* L2 access-policy window of the chunk walk over the mesh's walk table: `fraction` (0..1) of the
 * device's persisting L2 carve-out is given to the table's lines (hit property "persisting"), the
 * rest of the kernel's accesses stream.  0 (default) = no window; also PUMIPIC_L2_WINDOW in the
 * environment.  Results do not change.  Measured on the headline workload: a loss (0.252 -> 0.464 ms per
 * 10 M particles, profiles/r2F_l2_window_ab.txt) -- the records are served from L2 without it (DRAM traffic
 * is at its floor) and the carve-out takes L2 away from the particle columns; off unless asked for. */
void pp_search_set_l2_window(double fraction);
/* Counters of the most recent search on this mesh handle (synchronises the stream). */
pp_status pp_search_last_stats(pp_mesh* mesh, pp_search_stats* stats_host, pp_stream stream);

/* Fused push + search: the push immediately followed by the new-API walk on the freshly pushed
 * target, positions kept in registers (x_tgt is written once, never re-read).
 *   push_from_orig == 0: x_tgt += distance*dir           (test/test_adj.cpp:550-562)
 *   push_from_orig != 0: x_tgt  = x_orig + distance*dir  (the PIC form of
 *                        test/pseudoPushAndSearch.cpp:104-114 with a per-particle direction)
 * Same results as the separate push followed by pp_search_mesh. */
pp_status pp_push_direction_search(pp_mesh* mesh, pp_ps* ps, const double* dir, double distance,
                                   int32_t push_from_orig, const pp_search_args* args,
                                   pp_search_stats* stats_host, pp_stream stream);
/* The same fused step for callers whose particle columns live in HOST memory (pinned memory
 * recommended): xtgt = x + distance*dir, then search_mesh seeded from the structure rows
 * (elem_ids passed empty, adjacency.tpp:504-515), results written back to host arrays.
 *   h_x_orig, h_dir  host [3][stride] in;  h_x_tgt host [3][stride] out;  h_elem_ids host [capacity] out
 * h_dir may be NULL when the previous call on this mesh uploaded the directions of the same structure
 * (same capacity): the column stays resident on the device and is not copied again.
 * The slot range is cut into `nparts` pieces at chunk boundaries (Sell-C-sigma, C = 32; other
 * structures run as one piece); host->device copies, the kernel and device->host copies of
 * successive pieces overlap on internal streams.  nparts <= 0 picks a default.  All work is
 * ordered after prior work on `stream`, and `stream` waits for the last copy. */
pp_status pp_push_direction_search_host(pp_mesh* mesh, pp_ps* ps, const double* h_x_orig,
                                        const double* h_dir, double* h_x_tgt, int32_t* h_elem_ids,
                                        int64_t stride, double distance, int32_t looplimit,
                                        int32_t nparts, pp_search_stats* stats_host,
                                        pp_stream stream);
/* Unfused PIC-form push: xtgt = x + distance*dir for masked slots. */
pp_status pp_push_from(pp_ps* ps, const double* x, double* xtgt, const double* dir,
                       int64_t stride, double distance, pp_stream stream);

/* ============================== gather (field -> particle) =============================== */
/* The interpolation helpers an application's push calls per particle (GITRm's Boris push feeds
 * pp_push_boris with them).  All outputs are component-major [ncomp][stride] like the particle
 * members; only masked slots are written.
 *
 * src/pumipic_adjacency.hpp:801-809 findBCCoordsInTet + :772-799 interpolateTetVtx /
 * interpolate3dFieldTet: barycentric coordinates of x in tet elem_ids[slot] (find_barycentric_tet,
 * :97-133), then out[c] = sum_f bcc[f] * field[vertex_opposite(f)*dof + c].  Slots with
 * elem_ids < 0 are skipped.  n_outside_host (may be NULL; non-NULL synchronises the stream)
 * receives the number of particles the reference's OMEGA_H_CHECKs would abort on (point outside
 * its element by more than 1e-10); their outputs are left untouched.
 *   field: device double[nverts*dof], vertex-major (an Omega_h vertex tag with dof components) */
pp_status pp_gather_tet_field(pp_mesh* mesh, pp_ps* ps, const double* x, int64_t stride,
                              const int32_t* elem_ids, const double* field, int32_t dof, double* out,
                              int32_t* n_outside_host, pp_stream stream);
/* src/pumipic_utils.hpp:298-321 interpolate2d_field: bilinear interpolation of component `comp` of
 * an ncomp-component table data[(i + j*nx)*ncomp + comp] on the uniform grid
 * (gridx0 + i*dx, gridz0 + j*dz) at (x or sqrt(x^2+y^2) when cyl_symm, z).  out: [stride]. */
pp_status pp_gather_grid2d(pp_ps* ps, const double* x, int64_t stride, const double* data,
                           double gridx0, double gridz0, double dx, double dz, int32_t nx, int32_t nz,
                           int32_t cyl_symm, int32_t ncomp, int32_t comp, double* out,
                           pp_stream stream);
/* src/pumipic_utils.hpp:439-456 interp2dVector: the three components of a 3-component table, the
 * first two rotated by atan2(y, x) when cyl_symm.  out: [3][stride]. */
pp_status pp_gather_grid2d_vector(pp_ps* ps, const double* x, int64_t stride, const double* data3,
                                  double gridx0, double gridz0, double dx, double dz, int32_t nx,
                                  int32_t nz, int32_t cyl_symm, double* out, pp_stream stream);
/* src/pumipic_utils.hpp:377-420 interpolate3d_field: trilinear interpolation of
 * data[i + j*nx + k*nx*ny] on the grid lines gridx[nx], gridy[ny], gridz[nz] (device arrays;
 * ny or nz may be 1).  out: [stride]. */
pp_status pp_gather_grid3d(pp_ps* ps, const double* x, int64_t stride, const double* data,
                           const double* gridx, const double* gridy, const double* gridz, int32_t nx,
                           int32_t ny, int32_t nz, double* out, pp_stream stream);

/* ============================== gyro scatter (2D) ======================================== */

/* test/gyroScatter.hpp:96-166 createGyroRingMappings (+ :25-90 searchAndBuildMap): for every
 * vertex x ring x point the 3 vertices of the triangle containing the ring point, or -1.
 * map_out: device int32[3*nverts*nrings*points_per_ring].  Forward and backward maps are
 * identical in the reference (:126-131), build one and use it twice. */
pp_status pp_gyro_ring_map(pp_mesh* mesh, double rmax, int32_t nrings, int32_t points_per_ring,
                           double theta_deg, int32_t* map_out, pp_search_stats* stats_host,
                           pp_stream stream);
/* test/gyroScatter.hpp:168-229 gyroScatter: scatter_w[nverts] (device) is zeroed and filled. */
pp_status pp_gyro_scatter(pp_mesh* mesh, pp_ps* ps, const int32_t* v2v, double rmax,
                          int32_t nrings, int32_t points_per_ring, double* scatter_w,
                          pp_stream stream);
/* test/gyroScatter.hpp:244-248 setSyncArray: sync[2v] = fwd[v], sync[2v+1] = bkwd[v] */
pp_status pp_gyro_interleave(const double* fwd, const double* bkwd, int32_t nverts,
                             double* sync_array, pp_stream stream);

/* ============================== phase timers (support/ppTiming.hpp) ====================== */

/* The library's phases carry the reference's labels ("pumipic search_mesh", "SCS rebuild", "SCS
 * particle migration", "gyro scatter" ...): always as NVTX ranges, and with pp_timing_enable(1)
 * also as CUDA-event timings on the phase's stream, accumulated per label like
 * pumipic::RecordTime (ppTiming.cpp:67-100).  Reading the table (count / get / summarize) waits for
 * the pending phases.  pp_timing_record adds a caller-measured time under a label (RecordTime);
 * pp_timing_summarize prints the reference's table to stderr (SummarizeTime :168-213; sort: 0
 * alphabetical, 1 order of first occurrence, 2 longest first, 3 shortest first). */
void pp_timing_enable(int32_t on);
void pp_timing_set_verbosity(int32_t verbosity);   /* -1 none, 0 summary (default), 1 every record */
void pp_timing_set_rank(int32_t rank);             /* the rank printed in front of the lines */
void pp_timing_record(const char* label, double seconds);
void pp_timing_reset(void);
int32_t pp_timing_count(void);
pp_status pp_timing_get(int32_t i, char* name, int32_t name_cap, double* total_s, double* min_s,
                        double* max_s, double* sum_sq, int64_t* calls);
void pp_timing_summarize(int32_t sort);

/* ============================== communication (NCCL over NVLink) ========================= */

/* One process per GPU.  Rank 0 calls pp_comm_unique_id and distributes the 128 bytes by any
 * means (torch.distributed, MPI, a file); every rank then calls pp_comm_create.  nranks == 1
 * needs no id and never loads NCCL.  Replaces the MPI_Comm of support/ViewComm.h. */
pp_status pp_comm_unique_id(uint8_t id_out[128]);
pp_status pp_comm_create(int32_t nranks, int32_t rank, const uint8_t id[128], pp_comm** out);
/* The application's own bootstrap instead of a hand-carried NCCL id -- the reference's world is MPI
 * (support/ViewComm.h takes an MPI_Comm; an MPI code passes a wrapper of MPI_Allgather on that
 * communicator).  `allgather(ctx, send, recv, bytes)`: every rank contributes `bytes` bytes of host
 * memory, recv receives nranks * bytes in rank order; returns 0 on success.  It is called collectively,
 * on the host, while the communicator or one of its peer-memory windows is being set up -- never in a
 * step.  use_nccl != 0: the NCCL id travels through it; the communicator then equals pp_comm_create's.
 * use_nccl == 0: no NCCL is loaded at all; migration (pp_ps_migrate), pp_comm_array_reduce and
 * pp_comm_allreduce run over the peer-memory windows (CUDA IPC -- also between several processes that
 * share ONE GPU, which is how the multi-rank tests run on a single-GPU box); alltoall / send / recv and
 * the comm plans need NCCL and fail with PP_ERR_INVALID on such a communicator. */
typedef int32_t (*pp_host_allgather_fn)(void* ctx, const void* send, void* recv, int64_t bytes);
pp_status pp_comm_create_hosted(int32_t nranks, int32_t rank, pp_host_allgather_fn allgather, void* ctx,
                                int32_t use_nccl, pp_comm** out);
pp_status pp_comm_destroy(pp_comm* comm);
/* Migration transport.  Default: a peer-memory window per rank (cudaMalloc + CUDA IPC, mapped into every
 * peer by the first pp_ps_migrate of the communicator, collectively): the pack kernel stores the leaving
 * particles straight into the destination GPU's memory over NVLink, counts and a step flag follow, the
 * receiver's kernels wait on the flag -- no NCCL call, no host round trip in the step.  pp_comm_set_p2p(0)
 * (process-wide, before the first migration) or a failed IPC mapping selects the NCCL path (AllGather of
 * counts + grouped Send/Recv).  pp_comm_set_p2p_window sizes the window's segment per (sender, step
 * parity) in bytes (default: room for 1/16 of the first migrated structure's slots, at least 24 MiB, the
 * largest request of any rank; a rank's window is 2 * nranks segments); particles that do not fit a
 * segment stay where they are for that step (pp_migrate_stats.deferred).  pp_comm_p2p_active tells
 * which path the communicator ended up with. */
void pp_comm_set_p2p(int32_t enable);
pp_status pp_comm_set_p2p_window(pp_comm* comm, int64_t bytes_per_peer);
int32_t pp_comm_p2p_active(const pp_comm* comm);
int32_t pp_comm_size(const pp_comm* comm);
int32_t pp_comm_rank(const pp_comm* comm);

/* PS_Comm_Allreduce / PS_Comm_Alltoall / PS_Comm_Send / PS_Comm_Recv (support/ViewComm.h:51-291)
 * on device buffers; dtype is a pp_dtype, op a pp_op (SUM/MAX/MIN).  Sends and receives that
 * must progress together go between pp_comm_group_start/end (as PS_Comm_Isend/Irecv + Waitall). */
pp_status pp_comm_allreduce(pp_comm* comm, const void* send, void* recv, int64_t count,
                            int32_t dtype, int32_t op, pp_stream stream);
pp_status pp_comm_alltoall(pp_comm* comm, const void* send, void* recv, int64_t count_per_peer,
                           int32_t dtype, pp_stream stream);
pp_status pp_comm_send(pp_comm* comm, const void* buf, int64_t count, int32_t dtype, int32_t peer,
                       pp_stream stream);
pp_status pp_comm_recv(pp_comm* comm, void* buf, int64_t count, int32_t dtype, int32_t peer,
                       pp_stream stream);
pp_status pp_comm_group_start(void);
pp_status pp_comm_group_end(void);

/* Mesh::reduceCommArray (src/pumipic_comm.cpp:223-440) for full-mesh PICparts: every rank holds
 * a copy of all nents entities; after the call every copy holds the SUM/MAX/MIN over ranks, or
 * for PP_BCAST the owner's value (ent_owner = Mesh::entOwners(edim), device int32[nents]).
 * comm_array: device [nents*nvals], entity-major (createCommArray, pumipic_comm.cpp:187-192).
 * Between GPUs with peer access (the default; PUMIPIC_P2P=0 / pp_comm_set_p2p(0) select ncclAllReduce) the
 * reduction runs over a peer-memory window: every rank copies its array into its window, reduces ITS
 * slice of all ranks' copies with direct NVLink loads in ascending rank order and stores the result
 * into every rank's window -- the same bits on every rank and in every run (the reference's
 * MPI_Allreduce / atomic merges are order-dependent).  The first call (and a later call with a larger
 * array) maps the windows, collectively; successive calls on one communicator must use one stream. */
pp_status pp_comm_array_reduce(pp_comm* comm, void* comm_array, int64_t nents, int32_t nvals,
                               int32_t dtype, int32_t op, const int32_t* ent_owner,
                               pp_stream stream);

/* Mesh::reduceCommArray for partially buffered PICparts (src/pumipic_comm.cpp:249-439) with the
 * set-up of Mesh::setupComm (:12-184).  A plan is built once per entity dimension from
 *   ent_gids[nents]   global id of every entity of this PICpart (Mesh::globalIds(edim))
 *   ent_owner[nents]  owning rank (Mesh::entOwners(edim))
 * (memspace says where the two arrays live; collective: every rank calls it).  The owner of an
 * entity must hold a copy of it.  pp_comm_plan_reduce then leaves, in every copy of every entity,
 * the SUM / MAX / MIN over all copies, or for PP_BCAST the owner's value: copies are sent to the
 * owner (fan in), merged there in ascending rank order (deterministic, unlike the reference's
 * atomics) and sent back (fan out), all on device buffers over NCCL.
 * comm_array: device [nents*nvals], entity-major (createCommArray, pumipic_comm.cpp:187-192). */
typedef struct pp_comm_plan pp_comm_plan;
pp_status pp_comm_plan_create(pp_comm* comm, int64_t nents, const int64_t* ent_gids,
                              const int32_t* ent_owner, int32_t memspace, pp_stream stream,
                              pp_comm_plan** out);
pp_status pp_comm_plan_destroy(pp_comm_plan* plan);
/* entities this rank sends to owners / receives as an owner per reduction */
pp_status pp_comm_plan_counts(const pp_comm_plan* plan, int64_t* n_send, int64_t* n_recv);
pp_status pp_comm_plan_reduce(pp_comm_plan* plan, void* comm_array, int32_t nvals, int32_t dtype,
                              int32_t op, pp_stream stream);

/* migrate (particle_structure.hpp:101-104; SCS_migrate.h:5-221): particles whose new_process is
 * another rank are packed (element sent as global id), exchanged all-to-all-v, mapped back to
 * local element ids and inserted by the rebuild that ends the call.  new_element is modified in
 * place exactly like the reference (sent particles become -1).  World distributor only. */
typedef struct pp_migrate_stats {
  int64_t sent, received;
  int64_t deferred;   /* peer-memory path: particles that did not fit the destination's window segment and
                       * stay on this rank (still unsafe) until the next migration; 0 in normal operation */
} pp_migrate_stats;
pp_status pp_ps_migrate(pp_ps* ps, pp_comm* comm, int32_t* new_element,
                        const int32_t* new_process, int32_t n_new,
                        const int32_t* new_particle_elements,
                        const void* const* new_particle_info, pp_migrate_stats* stats_host,
                        pp_stream stream);

/* ============================== particle load balancing ==================== */

/* The plan of ParticleBalancer::balance (src/pumipic_lb.cpp:478-511), i.e. what
 * engpar::balanceWeights returns for the N-graph of buildNgraph (:395-462), evaluated for ALL parts
 * from the global weight vector.  EnGPar is a third-party dependency that is not in the reference
 * tree and its parity is UNPINNED: instead of its neighbour-to-neighbour diffusion the plan is
 * solved directly as a transportation problem (csrc/pp_host_lb.cpp) -- surplus of the overloaded
 * parts -> the regions they share with underloaded parts -> their deficits, a maximum flow.  The
 * reference's own acceptance bounds (test/test_lb.cpp:126,176) are what the tests hold it to.
 *   sbars: nsbars regions, global id sbar_ids[i], sorted parts parts[parts_off[i]..parts_off[i+1]);
 *          graph vertex of part parts[parts_off[i]+j] in sbar i = sbar_ids[i] + j (< nverts)
 *   vert_weight[nverts]: particles per vertex;  forced[nranks] (or NULL): particles each part is
 *          already receiving from others (pumipic_lb.hpp:196-200)
 *   tol: target max/avg (1.05 = 5 %): at or below it nothing is planned;  step_factor in (0,1]:
 *          EnGPar's diffusion rate, validated but without effect on a direct solve
 * Output (pp_host_free each): nsends transfers "send_weight[i] particles of vertex send_vert[i] to
 * part send_part[i]", ascending (vertex, part); imbalance[0] before, imbalance[1] planned (or NULL). */
pp_status pp_host_lb_plan(int32_t nranks, int32_t nsbars, const int32_t* sbar_ids,
                          const int32_t* parts_off, const int32_t* parts, int32_t nverts,
                          const double* vert_weight, const double* forced, double tol,
                          double step_factor, int32_t* nsends, int32_t** send_vert,
                          int32_t** send_part, double** send_weight, double* imbalance);

/* pumipic::ParticleBalancer (src/pumipic_lb.hpp:32-115) for one part.
 *   sbar table: the regions this part knows (pp_host_picpart_sbars); with a communicator of more
 *   than one rank the tables of all ranks are merged (collective), without one the table given
 *   must already be the global one.
 *   elem_sbar[nelems] = ParticleBalancer::getSbarIDs (the "sbar_id" tag), elem_owner[nelems] =
 *   Mesh::entOwners(dim); memspace says where the two arrays live. */
typedef struct pp_balancer pp_balancer;
pp_status pp_balancer_create(int32_t nranks, int32_t rank, int32_t nsbars, const int32_t* sbar_ids,
                             const int32_t* parts_off, const int32_t* parts, int32_t nelems,
                             const int32_t* elem_sbar, const int32_t* elem_owner, int32_t memspace,
                             pp_comm* comm, pp_stream stream, pp_balancer** out);
pp_status pp_balancer_destroy(pp_balancer* b);
/* graph vertices of all parts / of this part (global vertex id and sbar id, ascending; host
 * pointers owned by the balancer) */
pp_status pp_balancer_info(const pp_balancer* b, int32_t* nverts, int32_t* nlocal,
                           const int32_t** local_verts, const int32_t** local_sbars);
/* addWeights (pumipic_lb.hpp:133-208): counts, per own vertex, the masked particles that stay on
 * this rank (new_procs == rank, new_elems != -1) and, per destination, the ones already leaving;
 * new_elems / new_procs: device int32[capacity] indexed by slot.  The result lands in this part's
 * entries of the global weight vector. */
pp_status pp_balancer_add_weights_ps(pp_balancer* b, pp_ps* ps, const int32_t* new_elems,
                                     const int32_t* new_procs, pp_stream stream);
/* addWeights (pumipic_lb.hpp:211-229): ptcls_per_elem = device int32[nelems] */
pp_status pp_balancer_add_weights_array(pp_balancer* b, const int32_t* ptcls_per_elem,
                                        pp_stream stream);
/* The global weight vector on the device: [nverts] vertex weights then [nranks] forced weights. */
pp_status pp_balancer_weights(pp_balancer* b, double** weights_dev, int64_t* n);
/* balance (pumipic_lb.cpp:478-511): sums the weight vector over the communicator (skipped when
 * comm is NULL or has one rank: the vector is then taken as already global), runs
 * pp_host_lb_plan and keeps this part's sends.  One rank: the empty plan. */
pp_status pp_balancer_balance(pp_balancer* b, pp_comm* comm, double tol, double step_factor,
                              pp_stream stream);
/* this part's plan: nsends x (sbar id, target part, weight); host pointers owned by the balancer */
pp_status pp_balancer_plan(const pp_balancer* b, int32_t* nsends, const int32_t** sbar,
                           const int32_t** part, const double** weight, double imbalance[2]);
/* selectParticles (pumipic_lb.hpp:231-289): re-targets new_procs of ceil(weight) particles per
 * planned send, particles in elements of other parts' cores first, then any.  Consumes the plan. */
pp_status pp_balancer_select_ps(pp_balancer* b, pp_ps* ps, const int32_t* new_elems,
                                int32_t* new_procs, pp_stream stream);
/* selectParticles (pumipic_lb.hpp:291-353): new_procs = device int32[nptcls], nptcls = sum of
 * ptcls_per_elem; the particles of element e are entries [scan(e), scan(e) + ptcls_per_elem[e]). */
pp_status pp_balancer_select_array(pp_balancer* b, const int32_t* ptcls_per_elem, int64_t nptcls,
                                   int32_t* new_procs, pp_stream stream);
/* repartition (pumipic_lb.hpp:355-366) = add_weights_ps + balance + select_ps;
 * partition (:368-381) = add_weights_array + balance + select_array. */
pp_status pp_balancer_repartition(pp_balancer* b, pp_comm* comm, pp_ps* ps, double tol,
                                  const int32_t* new_elems, int32_t* new_procs, double step_factor,
                                  pp_stream stream);
pp_status pp_balancer_partition(pp_balancer* b, pp_comm* comm, const int32_t* ptcls_per_elem,
                                int64_t nptcls, double tol, double step_factor, int32_t* new_procs,
                                pp_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PUMIPIC_B200_H */
