/*
 * pumipic_oracle.h -- CPU restatement of PUMI-PIC's per-timestep particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a plain-C restatement of the
 * reference algorithm (SCOREC/pumi-pic @ c09ad045), one loop per reference kernel, written
 * from the reference's semantics with every function citing the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libpumipic_b200.so) never
 * links, loads or calls anything in this directory.
 *
 * PARITY PINNING: the real reference cannot be compiled as a whole here (needs Kokkos, Omega_h,
 * EnGPar, MPI -- none present, no network).  The oracle is pinned
 *   (1) against the reference's own golden vectors: test/search2d.cpp (15 element-id cases),
 *       test/moller_trumbore_line_tri_test.cpp, src/unit_tests.hpp barycentric values and
 *       test/pseudoXGCm_scatter.cpp vertex values (tests/test_oracle_golden.py);
 *   (2) against the reference's own source, compiled unmodified from /root/reference over
 *       stand-ins for the Omega_h / Kokkos types (oracle/_ref/, built by
 *       oracle/build_ref_primitives.py): every geometric primitive, and the four searches
 *       themselves (search_mesh -> trace_particle_through_mesh, search_mesh_2d, legacy 3D
 *       search_mesh, search_mesh_3d), the gyro ring mapping and scatter, the elliptical push
 *       and setUnsafeProcs give bit-identical results to the functions below on
 *       random and degenerate inputs (tests/test_oracle_vs_reference_source.py);
 *   (3) on geometry, with test/test_adj.cpp's property checks at a strict tolerance
 *       (tests/test_oracle_properties.py).
 * Not pinned (no reference test or source fixes them): see DESIGN.md section 2 "Unpinned".
 *
 * Omega_h small-vector arithmetic (cross, inner_product, norm, ...) lives in a third-party
 * dependency that is not vendored in the reference tree (SCOREC/omega_h, scorec-v10.8.4 in the
 * reference CI); its published formulas are restated in the static helpers of the .c file.
 *
 * Layout conventions (the reference's own):
 *   - particle members are component-major SoA ("LayoutLeft"): value(slot, i) = a[i*stride+slot]
 *   - every per-slot loop runs over ALL `cap` slots of the structure; `slot_elem[s]` is the
 *     element of the structure row that owns slot s, `mask[s]` says whether it holds a particle
 *   - mesh arrays use Omega_h numbering: coords[nverts*dim], elem2verts[nelems*(dim+1)],
 *     elem2sides[nelems*(dim+1)], side2verts[nsides*dim]
 */
#ifndef PUMIPIC_ORACLE_H
#define PUMIPIC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh {
  int dim, nverts, nelems, nsides;
  const double* coords;
  const int* elem2verts;
  const int* elem2sides;
  const int* side2verts;
  /* derived (owned) */
  int* side2elem_off;   /* ask_up(dim-1, dim).a2ab */
  int* side2elem;       /* ask_up(dim-1, dim).ab2b, ascending element id */
  int* dual_off;        /* ask_dual().a2ab */
  int* dual;            /* ask_dual().ab2b : neighbour across each NON-exposed side, local order */
  signed char* exposed; /* mark_exposed_sides */
  double* vol;          /* measure_elements_real */
  int* vert2elem_off;   /* ask_up(0, dim).a2ab */
  int* vert2elem;       /* ask_up(0, dim).ab2b, ascending element id */
} orc_mesh;

orc_mesh* orc_mesh_create(int dim, int nverts, const double* coords, int nelems,
                          const int* elem2verts, int nsides, const int* elem2sides,
                          const int* side2verts);
void orc_mesh_destroy(orc_mesh* m);
double orc_compute_tolerance(const orc_mesh* m);
void orc_set_num_threads(int n);
int orc_get_max_threads(void);

/* --- geometry primitives (exported for the known-answer tests) --- */
int orc_barycentric_tet(double vol, const double M[12], const double p[3], double bcc[4]);
void orc_barycentric_tri(double area, const double M[6], const double p[2], double bcc[3]);
int orc_find_barycentric_tet(const double M[12], const double p[3], double bcc[4]);
int orc_all_positive(const double* v, int n, double tol);
int orc_min_index(const double* v, int n);
int orc_max_index(const double* v, int n);
int orc_min3(const double v[3]);
int orc_is_face_flipped_3d(int fi, const int fv[3], const int tv[4]);
int orc_is_face_flipped_2d(const int ev[2], const int tv[3]);
int orc_ray_intersects_triangle(const double face[9], const double orig[3], const double dest[3],
                                double xpoint[3], double tol, int flip, double* dproj,
                                double* closeness, double* param);
int orc_line_segment_intersects_triangle(const double face[9], const double orig[3],
                                         const double dest[3], double xpoint[3], double tol,
                                         int flip, double* dproj, double* closeness,
                                         double* param);
int orc_line_edge_2d(const double edge[4], const double orig[2], const double dest[2],
                     double xpoint[2], double tol, int flip);
int orc_line_triangle_intx_simple(const double abc[9], const double origin[3],
                                  const double dest[3], double xpoint[3], double* dproj,
                                  int reverse, double tol);

/* --- searches --- */
typedef struct orc_search_stats {
  int loops;          /* walk iterations executed (reference `loops`) */
  int not_in_elem;    /* particles deleted by check_initial_parents */
  int not_found;      /* particles deleted by the loop limit */
  int aborted;        /* legacy search: OMEGA_H_CHECK(origin in element) would have fired */
} orc_search_stats;

/* new API: src/pumipic_adjacency.tpp:642 search_mesh -> :461 trace_particle_through_mesh.
 * elem_ids_empty != 0 reproduces "elem_ids.size()==0" (array is then fully overwritten).
 * inter_empty != 0 reproduces empty inter_faces/inter_points (fresh 0 / -1 arrays). */
int orc_search_mesh(const orc_mesh* m, int cap, const int* slot_elem, const unsigned char* mask,
                    const double* x_orig, const double* x_tgt, long stride,
                    int* elem_ids, int elem_ids_empty, int require_intersection,
                    int* inter_faces, double* inter_points, int inter_empty, int looplimit,
                    orc_search_stats* stats);

/* src/pumipic_adjacency.hpp:1013 search_mesh_2d */
int orc_search_mesh_2d(const orc_mesh* m, int cap, const int* slot_elem,
                       const unsigned char* mask, const double* x_tgt, long stride,
                       int* elem_ids, int looplimit, orc_search_stats* stats);

/* src/pumipic_adjacency.hpp:559 legacy 3D search_mesh (line-triangle + dual graph) */
int orc_search_mesh_legacy3d(const orc_mesh* m, int cap, const int* slot_elem,
                             const unsigned char* mask, const double* x_orig,
                             const double* x_tgt, long stride, int* elem_ids,
                             int elem_ids_empty, double* xpoints, int* xface, int looplimit,
                             orc_search_stats* stats);

/* src/pumipic_adjacency.hpp:316 search_mesh_3d (3 kernels per iteration, tol 1e-20) */
int orc_barycentric_coords_tet(const double M[12], const double p[3], double bcc[4], double tol);
int orc_search_mesh_3d(const orc_mesh* m, int cap, const int* slot_elem,
                       const unsigned char* mask, const double* x_orig, const double* x_tgt,
                       long stride, int* elem_ids, int elem_ids_empty, double* xpoints,
                       int* xface, int looplimit, orc_search_stats* stats);

/* --- gather: src/pumipic_adjacency.hpp:770-809, src/pumipic_utils.hpp:186-456 --- */
double orc_interpolate_tet_vtx(const orc_mesh* m, const double* field, int elem,
                               const double bcc[4], int dof, int comp);
int orc_gather_tet_field(const orc_mesh* m, int cap, const unsigned char* mask, const double* x,
                         long stride, const int* elem_ids, const double* field, int dof,
                         double* out);
double orc_interpolate2d_field(const double* data, double gridx0, double gridz0, double dx,
                               double dz, int nx, int nz, const double pos[3], int cyl, int nComp,
                               int comp);
void orc_interp2d_vector(const double* data3, double gridx0, double gridz0, double dx, double dz,
                         int nx, int nz, const double pos[3], double field[3], int cyl);
double orc_interpolate3d_field(double x, double y, double z, int nx, int ny, int nz,
                               const double* gridx, const double* gridy, const double* gridz,
                               const double* data);
void orc_gather_grid2d_vector(int cap, const unsigned char* mask, const double* x, long stride,
                              const double* data3, double gridx0, double gridz0, double dx,
                              double dz, int nx, int nz, int cyl, double* out);
void orc_gather_grid3d(int cap, const unsigned char* mask, const double* x, long stride,
                       const double* data, const double* gridx, const double* gridy,
                       const double* gridz, int nx, int ny, int nz, double* out);

/* --- pushes --- */
/* test/pseudoPushAndSearch.cpp:87-118 */
void orc_push_constant(int cap, const unsigned char* mask, const double* x, double* xtgt,
                       long stride, double distance, double dx, double dy, double dz);
/* test/test_adj.cpp:550-562 */
void orc_push_direction(int cap, const unsigned char* mask, double* tgt, const double* dir,
                        long stride, double distance);
/* test/ellipticalPush.hpp:10-34 (setup) and :36-70 (push) */
void orc_elliptical_setup(int cap, const unsigned char* mask, const double* x, long stride,
                          float* b, float* phi, double h, double k, double d);
void orc_elliptical_push(int cap, const int* slot_elem, const unsigned char* mask, double* xtgt,
                         long stride, const float* b, float* phi, const int* class_ids,
                         double h, double k, double d, double deg);
/* src/pumipic_push.hpp:17-74 (formula spec; the reference function is dead code) */
void orc_push_boris(int n, double* pos, double* pos_prev, double* vel, const double* efield,
                    const double* bfield, double dt);
/* test/pseudoPushAndSearch.cpp:142-154 updatePtclPositions (mask ignored) */
void orc_update_positions(int cap, double* x, double* xtgt, long stride);
/* src/pumipic_ptcl_ops.hpp:33-53 setUnsafeProcs */
void orc_set_unsafe_procs(int cap, const unsigned char* mask, const int* elems, const int* safe,
                          const int* owner, int self, int* new_elems, int* new_procs);

/* --- gyro scatter: test/gyroScatter.hpp --- */
/* :168-229.  scatter_w[nverts] is zeroed and filled. */
void orc_gyro_scatter(const orc_mesh* m, int cap, const int* slot_elem,
                      const unsigned char* mask, const int* v2v, double rmax, int nrings,
                      int points_per_ring, double* scatter_w);
/* :96-166 createGyroRingMappings + :25-90 searchAndBuildMap.  map[3*nverts*nrings*ppr] */
int orc_gyro_ring_map(const orc_mesh* m, double rmax, int nrings, int points_per_ring,
                      double theta_deg, int* map);

#ifdef __cplusplus
}
#endif
#endif
