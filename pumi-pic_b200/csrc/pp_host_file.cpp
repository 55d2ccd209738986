// pp_host_file.cpp -- host-side mesh container with every entity dimension and its tags, and
// the Omega_h binary mesh format (`.osh` directories) around it.  This is the file side of
// the reference reads its meshes with Omega_h::binary::read /
// read_mesh_file (test/test_file.cpp:24, src/pumipic_file.cpp:135) and writes PICparts with
// Omega_h::binary::write (pumipic_file.cpp:69).  Omega_h is not vendored in the reference tree
// (SCOREC/omega_h, CI pin scorec-v10.8.4), so the format is restated from the files themselves
//  and pinned by reading the reference's own fixtures
// (pumipic-data/xgc/*.osh, */*.ppm/*.osh; tests/test_picpart_file.py).
//
// No CUDA here: set-up code that the reference also runs on the host.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "pp_host_internal.hpp"
#include "pumipic_b200.h"

void pp_set_error(const char* fmt, ...);

namespace pph {

const int kTetFace[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}};  // Omega_h simplex templates
const int kTriEdge[3][2] = {{0, 1}, {1, 2}, {2, 0}};

int type_bytes(int type) {
  switch (type) {
    case PP_TAG_I8: return 1;
    case PP_TAG_I32: return 4;
    case PP_TAG_I64: return 8;
    case PP_TAG_F64: return 8;
  }
  return 0;
}

// Express a down-entity's stored vertex tuple in its parent's frame (SURVEY App. B "alignment
// code"): flip = code & 1, rot = (code >> 1) & 3; out[(j + rot) % n] = in[j]; then swap(out[1],
// out[2]) for a flipped triangle.
void align(int n, const int32_t* in, int code, int32_t* out) {
  const int flip = code & 1, rot = (code >> 1) & 3;
  for (int j = 0; j < n; ++j) out[(j + rot) % n] = in[j];
  if (flip && n == 3) std::swap(out[1], out[2]);
}

// The code that makes align(stored) == want, or -1.
int find_code(int n, const int32_t* stored, const int32_t* want) {
  for (int flip = 0; flip < (n == 3 ? 2 : 1); ++flip)
    for (int rot = 0; rot < n; ++rot) {
      const int code = (rot << 1) | flip;
      int32_t out[3];
      align(n, stored, code, out);
      bool ok = true;
      for (int j = 0; j < n; ++j) ok = ok && out[j] == want[j];
      if (ok) return code;
    }
  return -1;
}

HTag* HMesh::find(int d, const char* name) {
  for (auto& t : tags[d])
    if (t.name == name) return &t;
  return nullptr;
}
const HTag* HMesh::find(int d, const char* name) const {
  for (auto& t : tags[d])
    if (t.name == name) return &t;
  return nullptr;
}
void HMesh::set_tag(int d, const char* name, int ncomps, int type, const void* data) {
  HTag* t = find(d, name);
  if (!t) {
    tags[d].push_back(HTag());
    t = &tags[d].back();
    t->name = name;
  }
  t->ncomps = ncomps;
  t->type = type;
  const size_t nb = (size_t)nents[d] * ncomps * type_bytes(type);
  t->data.assign((const char*)data, (const char*)data + nb);
}

// entity -> vertices from the stored d -> d-1 adjacency and its alignment codes
bool HMesh::derive_verts() {
  if (dim < 1 || dim > 3) return false;
  verts[1] = down[1];
  if (dim >= 2) {
    const int n = nents[2];
    verts[2].resize((size_t)n * 3);
    for (int f = 0; f < n; ++f) {
      int32_t e0[2], e1[2];
      align(2, &verts[1][2 * (size_t)down[2][3 * (size_t)f + 0]], codes[2][3 * (size_t)f + 0], e0);
      align(2, &verts[1][2 * (size_t)down[2][3 * (size_t)f + 1]], codes[2][3 * (size_t)f + 1], e1);
      if (e0[1] != e1[0]) {
        pp_set_error("mesh: edge alignment of face %d is inconsistent", f);
        return false;
      }
      verts[2][3 * (size_t)f + 0] = e0[0];
      verts[2][3 * (size_t)f + 1] = e0[1];
      verts[2][3 * (size_t)f + 2] = e1[1];
    }
  }
  if (dim == 3) {
    const int n = nents[3];
    verts[3].resize((size_t)n * 4);
    for (int t = 0; t < n; ++t) {
      int32_t f0[3], f1[3];
      align(3, &verts[2][3 * (size_t)down[3][4 * (size_t)t + 0]], codes[3][4 * (size_t)t + 0], f0);
      align(3, &verts[2][3 * (size_t)down[3][4 * (size_t)t + 1]], codes[3][4 * (size_t)t + 1], f1);
      // face 0 = (v0, v2, v1), face 1 = (v0, v1, v3)
      if (f1[0] != f0[0] || f1[1] != f0[2]) {
        pp_set_error("mesh: face alignment of tet %d is inconsistent", t);
        return false;
      }
      verts[3][4 * (size_t)t + 0] = f0[0];
      verts[3][4 * (size_t)t + 1] = f0[2];
      verts[3][4 * (size_t)t + 2] = f0[1];
      verts[3][4 * (size_t)t + 3] = f1[2];
    }
  }
  return true;
}

namespace {
// Unique (d-1)-entities of a list of d-simplices given by their vertices (Omega_h find_unique:
// entities numbered by the lexicographic order of their sorted vertex tuple, each keeping the
// vertex order of its last use = highest parent, highest template position), plus the
// parent -> entity adjacency with alignment codes (reflect_down).
struct Key {
  int32_t v[3];
  int64_t use;
};
void unique_down(int d, int nhigh, const std::vector<int32_t>& hv, std::vector<int32_t>& lv,
                 std::vector<int32_t>& h2l, std::vector<int8_t>& hcodes) {
  const int nl = d + 1;      // (d-1)-entities per d-simplex
  const int deg = d;         // vertices per (d-1)-entity
  const int64_t nuses = (int64_t)nhigh * nl;
  auto use_verts = [&](int64_t u, int32_t* out) {
    const int64_t h = u / nl;
    const int k = (int)(u % nl);
    for (int i = 0; i < deg; ++i) {
      const int loc = d == 3 ? kTetFace[k][i] : (d == 2 ? kTriEdge[k][i] : k);
      out[i] = hv[(size_t)h * (d + 1) + loc];
    }
  };
  std::vector<Key> keys((size_t)nuses);
  for (int64_t u = 0; u < nuses; ++u) {
    int32_t v[3] = {-1, -1, -1};
    use_verts(u, v);
    std::sort(v, v + deg);
    keys[(size_t)u] = {{v[0], v[1], v[2]}, u};
  }
  std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
    for (int i = 0; i < 3; ++i)
      if (a.v[i] != b.v[i]) return a.v[i] < b.v[i];
    return a.use < b.use;
  });
  h2l.assign((size_t)nuses, -1);
  hcodes.assign((size_t)nuses, 0);
  lv.clear();
  int32_t n = 0;
  for (int64_t i = 0; i < nuses;) {
    int64_t j = i;
    while (j + 1 < nuses && memcmp(keys[(size_t)j + 1].v, keys[(size_t)i].v, sizeof(keys[0].v)) == 0) ++j;
    // the run's LAST use represents the entity (Omega_h marks a jump where a key differs from
    // the next one); pinned by pumipic-data/xgc/*.osh, whose edges are reproduced exactly
    int32_t rep[3];
    use_verts(keys[(size_t)j].use, rep);
    for (int k = 0; k < deg; ++k) lv.push_back(rep[k]);
    for (int64_t u = i; u <= j; ++u) {
      int32_t v[3];
      use_verts(keys[(size_t)u].use, v);
      h2l[(size_t)keys[(size_t)u].use] = n;
      if (deg > 1) hcodes[(size_t)keys[(size_t)u].use] = (int8_t)find_code(deg, rep, v);
    }
    ++n;
    i = j + 1;
  }
}
}  // namespace

// Build every entity dimension from element -> vertex connectivity the way
// Omega_h::build_from_elems2verts does (faces from tets, edges from faces).
bool HMesh::from_elems(int dim_, int nverts, const double* coords, int nelems, const int32_t* ev) {
  dim = dim_;
  for (int d = 0; d < 4; ++d) nents[d] = 0;
  nents[0] = nverts;
  nents[dim] = nelems;
  verts[dim].assign(ev, ev + (size_t)nelems * (dim + 1));
  for (int d = dim; d >= 2; --d) {
    unique_down(d, nents[d], verts[d], verts[d - 1], down[d], codes[d]);
    nents[d - 1] = (int)(verts[d - 1].size() / d);
  }
  down[1] = verts[1];
  for (int d = 0; d <= dim; ++d) {
    std::vector<int64_t> g((size_t)nents[d]);
    for (int i = 0; i < nents[d]; ++i) g[(size_t)i] = i;
    set_tag(d, "global", 1, PP_TAG_I64, g.data());
    if (d == 0) set_tag(0, "coordinates", dim, PP_TAG_F64, coords);
  }
  return true;
}

const double* HMesh::coords() const {
  const HTag* t = find(0, "coordinates");
  return t ? (const double*)t->data.data() : nullptr;
}

// ------------------------------------------------------------------ binary streams
bool Reader::load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize((size_t)(n > 0 ? n : 0));
  const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  pos = 0;
  return got == buf.size();
}
bool Reader::raw(void* out, size_t n) {
  if (pos + n > buf.size()) {
    ok = false;
    return false;
  }
  memcpy(out, buf.data() + pos, n);
  pos += n;
  return true;
}
// array = I32 n_entries, then (compressed) I64 nbytes + zlib stream, or the raw entries
bool Reader::array(int elem_bytes, bool compressed, std::vector<char>& out) {
  const int32_t n = value<int32_t>();
  if (!ok || n < 0) {
    ok = false;
    return false;
  }
  const size_t want = (size_t)n * elem_bytes;
  if (compressed) {
    const int64_t nb = value<int64_t>();
    // a deflate stream expands by at most ~1032:1: a count that the stream cannot produce is a
    // corrupt file, not a reason to allocate
    if (!ok || nb < 0 || (size_t)nb > buf.size() - pos || want > (size_t)nb * 1040 + 64) {
      ok = false;
      return false;
    }
    out.resize(want);
    uLongf dest = (uLongf)out.size();
    // zlib wants a non-null destination even for empty arrays
    char dummy = 0;
    const int rc = uncompress((Bytef*)(out.empty() ? &dummy : out.data()), &dest,
                              (const Bytef*)(buf.data() + pos), (uLong)nb);
    if (rc != Z_OK || dest != (uLongf)out.size()) {
      ok = false;
      return false;
    }
    pos += (size_t)nb;
  } else {
    if (want > buf.size() - pos) {
      ok = false;
      return false;
    }
    out.resize(want);
    if (!raw(out.data(), out.size())) return false;
  }
  return true;
}

void Writer::raw(const void* p, size_t n) { buf.insert(buf.end(), (const char*)p, (const char*)p + n); }
void Writer::array(const void* data, int64_t n, int elem_bytes, bool compressed) {
  value<int32_t>((int32_t)n);
  const size_t nb = (size_t)n * elem_bytes;
  if (compressed) {
    uLongf dest = compressBound((uLong)nb);
    std::vector<char> tmp((size_t)dest);
    char dummy = 0;
    compress2((Bytef*)tmp.data(), &dest, (const Bytef*)(nb ? data : &dummy), (uLong)nb, Z_BEST_SPEED);
    value<int64_t>((int64_t)dest);
    raw(tmp.data(), (size_t)dest);
  } else {
    raw(data, nb);
  }
}
bool Writer::save(const char* path) const {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  const size_t put = buf.empty() ? 0 : fwrite(buf.data(), 1, buf.size(), f);
  const bool closed = fclose(f) == 0;      // on every path: a short write must not leak the handle
  return put == buf.size() && closed;
}

bool read_small_int(const std::string& path, int* out) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  const int n = fscanf(f, "%d", out);
  fclose(f);
  return n == 1;
}

// `.osh` directory, format version 9: nparts, version, <rank>.osh
bool read_osh(const char* path, HMesh& m) {
  const std::string dir(path);
  int nparts = 1, version = 9;
  if (!read_small_int(dir + "/nparts", &nparts)) {
    pp_set_error("read_osh: %s/nparts is missing (not an .osh directory)", path);
    return false;
  }
  read_small_int(dir + "/version", &version);
  if (nparts != 1) {
    pp_set_error("read_osh: %s has %d parts; only serial meshes are read (the reference loads "
                 "the full mesh in serial on every rank, test_file.cpp:24)", path, nparts);
    return false;
  }
  Reader r;
  if (!r.load((dir + "/0.osh").c_str())) {
    pp_set_error("read_osh: cannot read %s/0.osh", path);
    return false;
  }
  const uint8_t m0 = r.value<uint8_t>(), m1 = r.value<uint8_t>();
  if (m0 != 0xA1 || m1 != 0x1A) {
    pp_set_error("read_osh: %s/0.osh has a bad magic number", path);
    return false;
  }
  const bool comp = r.value<int8_t>() != 0;
  m.family = r.value<int8_t>();
  m.dim = r.value<int8_t>();
  m.comm_size = r.value<int32_t>();
  m.comm_rank = r.value<int32_t>();
  m.parting = r.value<int8_t>();
  m.nghost = r.value<int32_t>();
  const int8_t have_hints = r.value<int8_t>();
  if (m.family != 0 || have_hints != 0 || m.dim < 1 || m.dim > 3) {
    pp_set_error("read_osh: %s: only simplex meshes without parting hints are supported", path);
    return false;
  }
  m.version = version;
  for (int d = 0; d < 4; ++d) m.nents[d] = 0;
  m.nents[0] = r.value<int32_t>();
  for (int d = 1; d <= m.dim; ++d) {
    std::vector<char> a;
    if (!r.array(4, comp, a)) break;
    m.down[d].assign((const int32_t*)a.data(), (const int32_t*)(a.data() + a.size()));
    m.nents[d] = (int)(m.down[d].size() / (d + 1));
    if (d > 1) {
      if (!r.array(1, comp, a)) break;
      m.codes[d].assign((const int8_t*)a.data(), (const int8_t*)(a.data() + a.size()));
    }
  }
  for (int d = 0; d <= m.dim && r.ok; ++d) {
    const int32_t ntags = r.value<int32_t>();
    for (int i = 0; i < ntags && r.ok; ++i) {
      HTag t;
      const int32_t nl = r.value<int32_t>();
      if (!r.ok || nl < 0 || nl > 4096) {
        r.ok = false;
        break;
      }
      t.name.resize((size_t)nl);
      r.raw(&t.name[0], (size_t)nl);
      t.ncomps = r.value<int8_t>();
      t.type = r.value<int8_t>();
      const int tb = type_bytes(t.type);
      if (!tb) {
        pp_set_error("read_osh: %s: tag %s has unknown type %d", path, t.name.c_str(), t.type);
        return false;
      }
      r.array(tb, comp, t.data);
      if (r.ok && t.data.size() != (size_t)m.nents[d] * t.ncomps * tb) r.ok = false;
      m.tags[d].push_back(std::move(t));
    }
    if (m.comm_size > 1 && r.ok) {  // owners (ranks, idxs): not used for serial files
      std::vector<char> a;
      r.array(4, comp, a);
      r.array(4, comp, a);
    }
  }
  if (!r.ok) {
    pp_set_error("read_osh: %s/0.osh is truncated or corrupt", path);
    return false;
  }
  // class sets and the has-parents flag: kept verbatim
  m.trailer.assign(r.buf.begin() + (long)r.pos, r.buf.end());
  return m.derive_verts();
}

bool write_osh(const HMesh& m, const char* path) {
  const std::string dir(path);
  mkdir(dir.c_str(), 0777);
  struct stat st;
  if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) {
    pp_set_error("write_osh: cannot create directory %s", path);
    return false;
  }
  auto put_text = [&](const char* name, int v) {
    FILE* f = fopen((dir + "/" + name).c_str(), "w");
    if (!f) return false;
    fprintf(f, "%d\n", v);
    return fclose(f) == 0;
  };
  if (!put_text("nparts", 1) || !put_text("version", m.version)) {
    pp_set_error("write_osh: cannot write into %s", path);
    return false;
  }
  Writer w;
  const bool comp = true;
  w.value<uint8_t>(0xA1);
  w.value<uint8_t>(0x1A);
  w.value<int8_t>(comp);
  w.value<int8_t>((int8_t)m.family);
  w.value<int8_t>((int8_t)m.dim);
  w.value<int32_t>(1);  // comm size / rank: a serial file
  w.value<int32_t>(0);
  w.value<int8_t>((int8_t)m.parting);
  w.value<int32_t>(m.nghost);
  w.value<int8_t>(0);   // no parting hints
  w.value<int32_t>(m.nents[0]);
  for (int d = 1; d <= m.dim; ++d) {
    w.array(m.down[d].data(), (int64_t)m.down[d].size(), 4, comp);
    if (d > 1) w.array(m.codes[d].data(), (int64_t)m.codes[d].size(), 1, comp);
  }
  for (int d = 0; d <= m.dim; ++d) {
    w.value<int32_t>((int32_t)m.tags[d].size());
    for (const HTag& t : m.tags[d]) {
      w.value<int32_t>((int32_t)t.name.size());
      w.raw(t.name.data(), t.name.size());
      w.value<int8_t>((int8_t)t.ncomps);
      w.value<int8_t>((int8_t)t.type);
      const int tb = type_bytes(t.type);
      w.array(t.data.data(), (int64_t)(t.data.size() / tb), tb, comp);
    }
  }
  if (!m.trailer.empty()) {
    w.raw(m.trailer.data(), m.trailer.size());
  } else {
    w.value<int32_t>(0);  // no class sets
    w.value<int8_t>(0);   // no parent links
  }
  if (!w.save((dir + "/0.osh").c_str())) {
    pp_set_error("write_osh: cannot write %s/0.osh", path);
    return false;
  }
  return true;
}

}  // namespace pph

// ------------------------------------------------------------------ C ABI
using pph::HMesh;

extern "C" pp_status pp_host_mesh_read_osh(const char* path, pp_host_mesh** out) {
  if (!path || !out) {
    pp_set_error("pp_host_mesh_read_osh: bad argument");
    return PP_ERR_INVALID;
  }
  HMesh* m = new HMesh();
  if (!pph::read_osh(path, *m)) {
    delete m;
    return PP_ERR_INVALID;
  }
  *out = reinterpret_cast<pp_host_mesh*>(m);
  return PP_OK;
}

extern "C" pp_status pp_host_mesh_write_osh(const pp_host_mesh* mesh, const char* path) {
  if (!mesh || !path) {
    pp_set_error("pp_host_mesh_write_osh: bad argument");
    return PP_ERR_INVALID;
  }
  return pph::write_osh(*reinterpret_cast<const HMesh*>(mesh), path) ? PP_OK : PP_ERR_INVALID;
}

extern "C" pp_status pp_host_mesh_from_elems(int32_t dim, int32_t nverts, const double* coords,
                                             int32_t nelems, const int32_t* elem2verts,
                                             pp_host_mesh** out) {
  if (!(dim == 2 || dim == 3) || nverts <= 0 || nelems <= 0 || !coords || !elem2verts || !out) {
    pp_set_error("pp_host_mesh_from_elems: bad argument");
    return PP_ERR_INVALID;
  }
  for (int64_t i = 0; i < (int64_t)nelems * (dim + 1); ++i)
    if (elem2verts[i] < 0 || elem2verts[i] >= nverts) {
      pp_set_error("pp_host_mesh_from_elems: vertex id %d out of range", elem2verts[i]);
      return PP_ERR_INVALID;
    }
  HMesh* m = new HMesh();
  m->from_elems(dim, nverts, coords, nelems, elem2verts);
  *out = reinterpret_cast<pp_host_mesh*>(m);
  return PP_OK;
}

extern "C" void pp_host_mesh_destroy(pp_host_mesh* mesh) { delete reinterpret_cast<HMesh*>(mesh); }

extern "C" int32_t pp_host_mesh_dim(const pp_host_mesh* mesh) {
  return mesh ? reinterpret_cast<const HMesh*>(mesh)->dim : -1;
}
extern "C" int32_t pp_host_mesh_nents(const pp_host_mesh* mesh, int32_t d) {
  if (!mesh || d < 0 || d > 3) return -1;
  return reinterpret_cast<const HMesh*>(mesh)->nents[d];
}
extern "C" const int32_t* pp_host_mesh_down(const pp_host_mesh* mesh, int32_t d) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || d < 1 || d > m->dim) return nullptr;
  return m->down[d].data();
}
extern "C" const int8_t* pp_host_mesh_codes(const pp_host_mesh* mesh, int32_t d) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || d < 2 || d > m->dim) return nullptr;
  return m->codes[d].data();
}
extern "C" const int32_t* pp_host_mesh_ent2verts(const pp_host_mesh* mesh, int32_t d) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || d < 1 || d > m->dim) return nullptr;
  return m->verts[d].data();
}
extern "C" const double* pp_host_mesh_coords(const pp_host_mesh* mesh) {
  return mesh ? reinterpret_cast<const HMesh*>(mesh)->coords() : nullptr;
}
extern "C" int32_t pp_host_mesh_ntags(const pp_host_mesh* mesh, int32_t d) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || d < 0 || d > m->dim) return -1;
  return (int32_t)m->tags[d].size();
}
static void fill_tag(const pph::HTag& t, pp_host_tag* out) {
  out->name = t.name.c_str();
  out->ncomps = t.ncomps;
  out->type = t.type;
  out->nvalues = (int64_t)(t.data.size() / (size_t)pph::type_bytes(t.type));
  out->data = t.data.data();
}
extern "C" pp_status pp_host_mesh_tag_at(const pp_host_mesh* mesh, int32_t d, int32_t i,
                                         pp_host_tag* out) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || !out || d < 0 || d > m->dim || i < 0 || i >= (int32_t)m->tags[d].size()) {
    pp_set_error("pp_host_mesh_tag_at: bad argument");
    return PP_ERR_INVALID;
  }
  fill_tag(m->tags[d][(size_t)i], out);
  return PP_OK;
}
extern "C" pp_status pp_host_mesh_find_tag(const pp_host_mesh* mesh, int32_t d, const char* name,
                                           pp_host_tag* out) {
  const HMesh* m = reinterpret_cast<const HMesh*>(mesh);
  if (!m || !out || !name || d < 0 || d > m->dim) {
    pp_set_error("pp_host_mesh_find_tag: bad argument");
    return PP_ERR_INVALID;
  }
  const pph::HTag* t = m->find(d, name);
  if (!t) {
    pp_set_error("pp_host_mesh_find_tag: no tag \"%s\" on dimension %d", name, d);
    return PP_ERR_INVALID;
  }
  fill_tag(*t, out);
  return PP_OK;
}
extern "C" pp_status pp_host_mesh_set_tag(pp_host_mesh* mesh, int32_t d, const char* name,
                                          int32_t ncomps, int32_t type, const void* data) {
  HMesh* m = reinterpret_cast<HMesh*>(mesh);
  if (!m || !name || !data || d < 0 || d > m->dim || ncomps < 1 || !pph::type_bytes(type)) {
    pp_set_error("pp_host_mesh_set_tag: bad argument");
    return PP_ERR_INVALID;
  }
  m->set_tag(d, name, ncomps, type, data);
  return PP_OK;
}

// Partition files (pumipic_input.cpp:44-89): `.ptn` = one owner per element; `.cpn` = N, then
// (class id, owner) pairs, elements owned through their class_id (setOwnerByClassification,
// part_construct.cpp:265-288).
extern "C" pp_status pp_host_read_partition(const char* path, int32_t nelems,
                                            const int32_t* elem_class, int32_t* owner_out) {
  if (!path || !owner_out || nelems < 0) {
    pp_set_error("pp_host_read_partition: bad argument");
    return PP_ERR_INVALID;
  }
  const char* dot = strrchr(path, '.');
  if (!dot) {
    pp_set_error("Filename provided has no extension (%s)", path);
    return PP_ERR_INVALID;
  }
  FILE* f = fopen(path, "r");
  if (!f) {
    pp_set_error("Cannot open file %s", path);
    return PP_ERR_INVALID;
  }
  pp_status st = PP_OK;
  if (strcmp(dot + 1, "ptn") == 0) {
    int own, n = 0;
    for (int i = 0; i < nelems; ++i) owner_out[i] = 0;
    while (fscanf(f, "%d", &own) == 1) {
      if (n < nelems) owner_out[n] = own;
      ++n;
    }
    if (n != nelems) {
      pp_set_error("pp_host_read_partition: %s holds %d owners for %d elements", path, n, nelems);
      st = PP_ERR_INVALID;
    }
  } else if (strcmp(dot + 1, "cpn") == 0) {
    int size = 0;
    if (!elem_class || fscanf(f, "%d", &size) != 1 || size < 0) {
      pp_set_error("pp_host_read_partition: %s needs the elements' class ids and a size line", path);
      st = PP_ERR_INVALID;
    } else {
      std::vector<int32_t> owners((size_t)size + 1, 0);
      int cid, own;
      while (fscanf(f, "%d %d", &cid, &own) == 2)
        if (cid >= 0 && cid <= size) owners[(size_t)cid] = own;
      for (int e = 0; e < nelems && st == PP_OK; ++e) {
        const int c = elem_class[e];
        if (c < 0 || c > size) {
          pp_set_error("Class id %d on element %d is outside the partition file's range [0,%d]", c,
                       e, size);
          st = PP_ERR_INVALID;
        } else {
          owner_out[e] = owners[(size_t)c];
        }
      }
    }
  } else {
    pp_set_error("Only .ptn and .cpn partitions are supported");
    st = PP_ERR_INVALID;
  }
  fclose(f);
  return st;
}
