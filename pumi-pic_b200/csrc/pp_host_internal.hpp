// pp_host_internal.hpp -- shared by the host-only (no CUDA) set-up sources: the mesh container
// with every entity dimension, the little-endian binary streams of the Omega_h file formats, and
// the PICpart record.  Not part of the public C ABI.
#ifndef PP_HOST_INTERNAL_HPP
#define PP_HOST_INTERNAL_HPP

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace pph {

extern const int kTetFace[4][3];
extern const int kTriEdge[3][2];

int type_bytes(int type);
void align(int n, const int32_t* in, int code, int32_t* out);
int find_code(int n, const int32_t* stored, const int32_t* want);

struct HTag {
  std::string name;
  int ncomps = 1;
  int type = 2;  // pp_host_tag_type
  std::vector<char> data;
};

// A simplicial mesh with all entity dimensions, numbered as Omega_h numbers them.
struct HMesh {
  int dim = 0;
  int nents[4] = {0, 0, 0, 0};
  std::vector<int32_t> down[4];   // d -> d-1 entities, [nents[d]*(d+1)], d >= 1
  std::vector<int8_t> codes[4];   // alignment codes of `down`, d >= 2
  std::vector<int32_t> verts[4];  // d -> vertices in template order (derived), d >= 1
  std::vector<HTag> tags[4];
  std::vector<char> trailer;      // class sets + parents flag of a file that was read
  int family = 0, parting = 0, comm_size = 1, comm_rank = 0, nghost = 0, version = 9;

  HTag* find(int d, const char* name);
  const HTag* find(int d, const char* name) const;
  void set_tag(int d, const char* name, int ncomps, int type, const void* data);
  template <class T>
  const T* tag_data(int d, const char* name) const {
    const HTag* t = find(d, name);
    return t ? reinterpret_cast<const T*>(t->data.data()) : nullptr;
  }
  bool derive_verts();
  bool from_elems(int dim, int nverts, const double* coords, int nelems, const int32_t* ev);
  const double* coords() const;
};

struct Reader {
  std::vector<char> buf;
  size_t pos = 0;
  bool ok = true;
  bool load(const char* path);
  bool raw(void* out, size_t n);
  template <class T>
  T value() {
    T v = T();
    raw(&v, sizeof(T));
    return v;
  }
  bool array(int elem_bytes, bool compressed, std::vector<char>& out);
  template <class T>
  bool typed_array(bool compressed, std::vector<T>& out) {
    std::vector<char> a;
    if (!array((int)sizeof(T), compressed, a)) return false;
    out.resize(a.size() / sizeof(T));
    if (!a.empty()) memcpy(out.data(), a.data(), a.size());
    return true;
  }
};

struct Writer {
  std::vector<char> buf;
  void raw(const void* p, size_t n);
  template <class T>
  void value(T v) {
    raw(&v, sizeof(T));
  }
  void array(const void* data, int64_t n, int elem_bytes, bool compressed);
  bool save(const char* path) const;
};

bool read_osh(const char* path, HMesh& m);
bool write_osh(const HMesh& m, const char* path);

struct Up {  // ask_up(bridge_dim, dim): bridge entity -> elements, ascending
  std::vector<int> off, val;
};
Up build_up(int nverts, int nelems, int nv, const int32_t* ev);
// element -> entities of dimension bridge_dim (vertices, edges or sides), `per_elem` each
bool elem_bridges(const HMesh& m, int bridge_dim, std::vector<int32_t>& out, int& per_elem);
void picpart_tags(const Up& u, int nverts, int nelems, const int32_t* owner, int nranks, int rank,
                  int buffer_method, int safe_method, int buffer_layers, int safe_layers,
                  std::vector<int>& is_safe, std::vector<int>& has_part);

// Per-dimension communication record of a PICpart: the members pumipic::Mesh keeps per
// dimension (pumipic_mesh.hpp:118-143) and pumipic::write stores (pumipic_file.cpp:85-114).
struct PicpartDim {
  int64_t num_entities = 0;                // entities of this dimension in the FULL mesh
  int32_t num_cores = 0;                   // other parts with entities here
  std::vector<int32_t> buffered_parts;     // [num_cores]
  std::vector<int32_t> offset_ents_per_rank;  // [nranks+1] entities here per owner, offset-summed
  std::vector<int32_t> ent_to_comm_arr_index; // [nents]
  std::vector<int32_t> is_complete_part;   // [nranks] 0 none here, 1 part of it (boundary), 2 all
  int32_t num_bounds = 0;                  // parts we hold only a boundary of
  int32_t num_boundaries = 0;              // parts that hold only a boundary of ours
  std::vector<int32_t> boundary_parts;     // [num_boundaries]
  std::vector<int32_t> offset_bounded;     // [nranks+1]
  std::vector<int32_t> bounded_ent_ids;    // rank-local ids of our entities on those boundaries
  std::vector<int32_t> ent_l2g;            // full-mesh index of every local entity (not in .ppm)
};

struct Picpart {
  int nranks = 1, rank = 0;
  bool is_full_mesh = false;
  HMesh mesh;
  PicpartDim d[4];
  // safe-zone overlap regions ("sbars") this part belongs to: global id -> sorted parts
  std::vector<int32_t> sbar_ids;
  std::vector<int32_t> sbar_parts_off, sbar_parts;
  int32_t max_sbar = 0;
};

}  // namespace pph

#endif
