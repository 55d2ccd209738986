// pp_host_ppm.cpp -- PICpart construction with its communication record, the safe-zone overlap
// regions ("sbars") of the particle balancer, and the `.ppm` file format.
//
// Follows, for one rank:
//   Mesh::constructPICPart   src/pumipic_part_construct.cpp:116-262 (owners :304-323, global
//                            numbering :335-385, entity selection :467-489, sub-mesh :514-595,
//                            tag conversion :597-617)
//   Mesh::setupComm          src/pumipic_comm.cpp:12-184
//   ParticleBalancer ctor    src/pumipic_lb.cpp:23-82 (buildLocalSbarMap :92-110, sendCoreSbars
//                            :112-181, globalNumberSbars :185-335, cleanSbars :337-345,
//                            numberElements :382-432)
//   pumipic::write / read    src/pumipic_file.cpp:45-205
// The reference exchanges boundary lists, safe flags and sbar tables between ranks with MPI; every
// rank holds the same full mesh and partition, so here what a peer would send is evaluated
// locally ("world" below) and no communicator is needed at set-up.
// Pinned against the reference's own output files pumipic-data/xgc/{24k,120k}_4.ppm
// (tests/test_picpart_file.py).  No CUDA.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "pp_host_internal.hpp"
#include "pumipic_b200.h"

void pp_set_error(const char* fmt, ...);

namespace pph {
namespace {

typedef std::set<int> Parts;
// pumipic_lb.hpp:15-24: the iteration order of the reference's table depends on this hash
struct PartsHash {
  size_t operator()(const Parts& res) const {
    int h = 0;
    std::hash<int> hasher;
    for (auto itr = res.begin(); itr != res.end(); ++itr) h ^= hasher(*itr);
    return h;
  }
};
typedef std::unordered_map<Parts, int, PartsHash> SBarUnmap;

// Everything about all ranks that a single rank's PICpart depends on.
struct World {
  int dim = 0, nranks = 1;
  const HMesh* full = nullptr;
  std::vector<int32_t> owner[4];              // per dimension, full mesh
  std::vector<int32_t> rank_offset[4];        // [nranks+1] entities per owner, offset-summed
  std::vector<int64_t> gids[4];
  std::vector<int32_t> rank_lids[4];
  std::vector<std::vector<uint8_t>> safe;     // [rank][nelems]
  std::vector<std::vector<int>> has_part;     // [rank][nranks]
};

// keep flags of rank s for every dimension: elements of buffered cores and their closure
void kept_entities(const World& w, int s, std::vector<uint8_t> keep[4]) {
  const HMesh& f = *w.full;
  const int dim = w.dim;
  keep[dim].assign((size_t)f.nents[dim], 0);
  for (int e = 0; e < f.nents[dim]; ++e) keep[dim][(size_t)e] = w.has_part[(size_t)s][(size_t)w.owner[dim][(size_t)e]] != 0;
  for (int d = dim; d >= 1; --d) {
    keep[d - 1].assign((size_t)f.nents[d - 1], 0);
    const int nd = d + 1;
    for (int i = 0; i < f.nents[d]; ++i)
      if (keep[d][(size_t)i])
        for (int k = 0; k < nd; ++k) keep[d - 1][(size_t)f.down[d][(size_t)i * nd + k]] = 1;
  }
}

bool build_world(const HMesh& full, const int32_t* elem_owner, int nranks, int bm, int sm, int bl,
                 int sl, int bridge_dim, World& w) {
  const int dim = full.dim;
  w.dim = dim;
  w.nranks = nranks;
  w.full = &full;
  // defineOwners: minimum owner of the adjacent elements, propagated down one dimension at a time
  w.owner[dim].assign(elem_owner, elem_owner + full.nents[dim]);
  for (int d = dim; d >= 1; --d) {
    w.owner[d - 1].assign((size_t)full.nents[d - 1], nranks);
    const int nd = d + 1;
    for (int i = 0; i < full.nents[d]; ++i)
      for (int k = 0; k < nd; ++k) {
        int32_t& o = w.owner[d - 1][(size_t)full.down[d][(size_t)i * nd + k]];
        o = std::min(o, w.owner[d][(size_t)i]);
      }
  }
  // createGlobalNumbering / rankLidNumbering: owner-major, stable within an owner
  for (int d = 0; d <= dim; ++d) {
    const int n = full.nents[d];
    w.rank_offset[d].assign((size_t)nranks + 1, 0);
    for (int i = 0; i < n; ++i) w.rank_offset[d][(size_t)w.owner[d][(size_t)i] + 1]++;
    for (int p = 0; p < nranks; ++p) w.rank_offset[d][(size_t)p + 1] += w.rank_offset[d][(size_t)p];
    std::vector<int32_t> seen((size_t)nranks, 0);
    w.gids[d].resize((size_t)n);
    w.rank_lids[d].resize((size_t)n);
    for (int i = 0; i < n; ++i) {
      const int o = w.owner[d][(size_t)i];
      w.rank_lids[d][(size_t)i] = seen[(size_t)o];
      w.gids[d][(size_t)i] = (int64_t)w.rank_offset[d][(size_t)o] + seen[(size_t)o]++;
    }
  }
  // safe zone and buffered parts of every rank
  std::vector<int32_t> bridges;
  int per_elem = 0;
  if (!elem_bridges(full, bridge_dim, bridges, per_elem)) return false;
  const int nbridges = full.nents[bridge_dim];
  const Up up = build_up(nbridges, full.nents[dim], per_elem, bridges.data());
  w.safe.resize((size_t)nranks);
  w.has_part.resize((size_t)nranks);
  for (int s = 0; s < nranks; ++s) {
    std::vector<int> is_safe;
    picpart_tags(up, nbridges, full.nents[dim], elem_owner, nranks, s, bm, sm, bl, sl, is_safe,
                 w.has_part[(size_t)s]);
    w.safe[(size_t)s].assign(is_safe.begin(), is_safe.end());
  }
  return true;
}

// bufferedRanks(dim) of rank s: the other parts with elements in its PICpart
std::vector<int> buffer_ranks_of(const World& w, int s) {
  std::vector<int> out;
  const auto& off = w.rank_offset[w.dim];
  for (int p = 0; p < w.nranks; ++p)
    if (p != s && w.has_part[(size_t)s][(size_t)p] && off[(size_t)p + 1] != off[(size_t)p]) out.push_back(p);
  return out;
}

// The ParticleBalancer constructor for every rank at once.
struct Sbars {
  std::vector<int32_t> elem_sbar;            // global sbar id of every element of the full mesh
  std::vector<SBarUnmap> table;              // per rank, after cleanSbars, values = global ids
  int max_sbar = 0;
};
void build_sbars(const World& w, Sbars& out) {
  const int nranks = w.nranks, dim = w.dim;
  const int nelems = w.full->nents[dim];
  std::vector<std::vector<int>> bufs((size_t)nranks);
  for (int s = 0; s < nranks; ++s) bufs[(size_t)s] = buffer_ranks_of(w, s);
  // buildLocalSbarMap: core elements in rank-local order
  std::vector<SBarUnmap> maps((size_t)nranks);
  std::vector<int32_t> local_id((size_t)nelems, -1);   // owner's local sbar id of each element
  {
    std::vector<std::map<std::vector<uint64_t>, int>> memo((size_t)nranks);
    const size_t words = ((size_t)nranks + 63) / 64;
    std::vector<uint64_t> mask(words);
    for (int e = 0; e < nelems; ++e) {
      const int s = w.owner[dim][(size_t)e];
      std::fill(mask.begin(), mask.end(), 0);
      for (int b : bufs[(size_t)s])
        if (w.safe[(size_t)b][(size_t)e]) mask[(size_t)b / 64] |= 1ull << (b % 64);
      auto it = memo[(size_t)s].find(mask);
      if (it == memo[(size_t)s].end()) {
        Parts parts;
        parts.insert(s);
        for (int b : bufs[(size_t)s])
          if (w.safe[(size_t)b][(size_t)e]) parts.insert(b);
        auto& m = maps[(size_t)s];
        auto f = m.find(parts);
        if (f == m.end()) f = m.insert(std::make_pair(parts, (int)m.size())).first;
        it = memo[(size_t)s].insert(std::make_pair(mask, f->second)).first;
      }
      local_id[(size_t)e] = it->second;
    }
  }
  // sendCoreSbars: every rank sends its table (in iteration order) to its buffer ranks
  std::vector<std::vector<Parts>> message((size_t)nranks);
  for (int s = 0; s < nranks; ++s)
    for (auto it = maps[(size_t)s].begin(); it != maps[(size_t)s].end(); ++it) message[(size_t)s].push_back(it->first);
  for (int r = 0; r < nranks; ++r)
    for (int b : bufs[(size_t)r])
      for (const Parts& p : message[(size_t)b]) {
        auto& m = maps[(size_t)r];
        if (m.find(p) == m.end()) m.insert(std::make_pair(p, (int)m.size()));
      }
  // globalNumberSbars: the smallest part of an sbar numbers it; ids advance by the sbar's size
  std::vector<int> start((size_t)nranks + 1, 0);
  for (int s = 0; s < nranks; ++s) {
    int owned = 0;
    for (auto it = maps[(size_t)s].begin(); it != maps[(size_t)s].end(); ++it)
      if (*(it->first.begin()) == s) owned += (int)it->first.size();
    start[(size_t)s + 1] = start[(size_t)s] + owned;
  }
  out.max_sbar = start[(size_t)nranks];
  std::map<Parts, int> global_id;
  for (int s = 0; s < nranks; ++s) {
    int next = start[(size_t)s];
    for (auto it = maps[(size_t)s].begin(); it != maps[(size_t)s].end(); ++it)
      if (*(it->first.begin()) == s) {
        global_id[it->first] = next;
        next += (int)it->first.size();
      }
  }
  // numberElements: each core numbers its own elements, buffers receive that numbering
  std::vector<std::vector<int>> l2g((size_t)nranks);
  for (int s = 0; s < nranks; ++s) {
    l2g[(size_t)s].assign(maps[(size_t)s].size(), -1);
    for (auto it = maps[(size_t)s].begin(); it != maps[(size_t)s].end(); ++it) {
      auto g = global_id.find(it->first);
      if (g != global_id.end()) l2g[(size_t)s][(size_t)it->second] = g->second;
    }
  }
  out.elem_sbar.resize((size_t)nelems);
  for (int e = 0; e < nelems; ++e)
    out.elem_sbar[(size_t)e] = l2g[(size_t)w.owner[dim][(size_t)e]][(size_t)local_id[(size_t)e]];
  // cleanSbars + the final id conversion of numberElements
  out.table.resize((size_t)nranks);
  for (int r = 0; r < nranks; ++r) {
    SBarUnmap& m = maps[(size_t)r];
    for (auto it = m.begin(); it != m.end();) {
      if (it->first.find(r) == it->first.end())
        it = m.erase(it);
      else {
        it->second = l2g[(size_t)r][(size_t)it->second];
        ++it;
      }
    }
    out.table[(size_t)r].swap(m);
  }
}

void store_sbar_table(const SBarUnmap& t, int max_sbar, Picpart& pp) {
  pp.sbar_ids.clear();
  pp.sbar_parts.clear();
  pp.sbar_parts_off.assign(1, 0);
  for (auto it = t.begin(); it != t.end(); ++it) {
    pp.sbar_ids.push_back(it->second);
    for (int p : it->first) pp.sbar_parts.push_back(p);
    pp.sbar_parts_off.push_back((int32_t)pp.sbar_parts.size());
  }
  pp.max_sbar = max_sbar;
}

// Mesh::setupComm for dimension d of rank r.  keep_of(s) gives the kept entities of rank s.
void setup_comm(const World& w, int r, int d, const std::vector<int32_t>& l2g,
                const std::vector<std::vector<uint8_t>>& keep_all_d, PicpartDim& out) {
  const int nranks = w.nranks, dim = w.dim;
  const int n = (int)l2g.size();
  const auto& gown = w.owner[d];
  const auto& goff = w.rank_offset[d];
  out.num_entities = w.full->nents[d];
  out.ent_l2g = l2g;
  std::vector<int32_t>& poff = out.offset_ents_per_rank;
  poff.assign((size_t)nranks + 1, 0);
  for (int i = 0; i < n; ++i) poff[(size_t)gown[(size_t)l2g[(size_t)i]] + 1]++;
  for (int p = 0; p < nranks; ++p) poff[(size_t)p + 1] += poff[(size_t)p];
  if (d == dim) {
    int cores = 0;
    for (int p = 0; p < nranks; ++p) cores += w.has_part[(size_t)r][(size_t)p] > 0;
    out.num_cores = cores - 1;
  }
  if (out.num_cores == 0 || d != dim) {   // comm.cpp:22-31
    int cores = 0;
    for (int p = 0; p < nranks; ++p) cores += poff[(size_t)p + 1] != poff[(size_t)p];
    out.num_cores = cores - 1;
  }
  out.buffered_parts.assign((size_t)std::max(out.num_cores, 0), 0);
  {
    size_t index = 0;
    for (int p = 0; p < nranks; ++p)
      if (poff[(size_t)p + 1] != poff[(size_t)p] && p != r && index < out.buffered_parts.size())
        out.buffered_parts[index++] = p;
  }
  out.is_complete_part.assign((size_t)nranks, 0);
  for (int p = 0; p < nranks; ++p) {
    const int gdiff = goff[(size_t)p + 1] - goff[(size_t)p], pdiff = poff[(size_t)p + 1] - poff[(size_t)p];
    out.is_complete_part[(size_t)p] = (gdiff == pdiff) + (pdiff != 0);
  }
  // rank-local ids: global numbering for complete parts, ascending entity order for boundaries
  std::vector<int32_t> next((size_t)nranks, 0);
  out.ent_to_comm_arr_index.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    const int g = l2g[(size_t)i], o = gown[(size_t)g];
    const int lid = out.is_complete_part[(size_t)o] == 1 ? next[(size_t)o]++ : w.rank_lids[d][(size_t)g];
    out.ent_to_comm_arr_index[(size_t)i] = lid + poff[(size_t)o];
  }
  out.num_bounds = out.num_boundaries = 0;
  out.boundary_parts.clear();
  out.offset_bounded.clear();
  out.bounded_ent_ids.clear();
  if (d == dim) return;
  out.num_bounds = out.num_cores - /* num_cores[dim] */ [&] {
    int cores = 0;
    for (int p = 0; p < nranks; ++p) cores += w.has_part[(size_t)r][(size_t)p] > 0;
    return cores - 1;
  }();
  // what every other rank holds of OUR entities without holding all of them
  out.offset_bounded.assign((size_t)nranks + 1, 0);
  const int mine = goff[(size_t)r + 1] - goff[(size_t)r];
  for (int s = 0; s < nranks; ++s) {
    std::vector<int32_t> rl;
    if (s != r) {
      const std::vector<uint8_t>& keep = keep_all_d[(size_t)s];
      for (int g = 0; g < w.full->nents[d]; ++g)
        if (keep[(size_t)g] && gown[(size_t)g] == r) rl.push_back(w.rank_lids[d][(size_t)g]);
      if ((int)rl.size() == mine) rl.clear();   // complete copy: not a boundary
    }
    if (!rl.empty()) {
      ++out.num_boundaries;
      out.boundary_parts.push_back(s);
      out.bounded_ent_ids.insert(out.bounded_ent_ids.end(), rl.begin(), rl.end());
    }
    out.offset_bounded[(size_t)s + 1] = (int32_t)out.bounded_ent_ids.size();
  }
}

void convert_tag(const HMesh& full, HMesh& part, int d, const std::vector<int32_t>& l2g, const HTag& t,
                 const char* new_name) {
  const size_t vb = (size_t)t.ncomps * type_bytes(t.type);
  std::vector<char> data(l2g.size() * vb);
  for (size_t i = 0; i < l2g.size(); ++i)
    memcpy(data.data() + i * vb, t.data.data() + (size_t)l2g[i] * vb, vb);
  char dummy = 0;
  part.set_tag(d, new_name, t.ncomps, t.type, data.empty() ? (const void*)&dummy : data.data());
  (void)full;
}

template <class T>
void gather_tag(HMesh& part, int d, const char* name, int type, const std::vector<T>& fullv,
                const std::vector<int32_t>& l2g) {
  std::vector<T> v(l2g.size() + 1);
  for (size_t i = 0; i < l2g.size(); ++i) v[i] = fullv[(size_t)l2g[i]];
  part.set_tag(d, name, 1, type, v.data());
}

bool build_picpart(const HMesh& full, const int32_t* elem_owner, int nranks, int rank, int bm,
                   int sm, int bl, int sl, int bridge_dim, Picpart& pp) {
  enum { FULL = 0 };
  const int dim = full.dim;
  World w;
  if (!build_world(full, elem_owner, nranks, bm, sm, bl, sl, bridge_dim, w)) {
    pp_set_error("pp_host_picpart_build: bridge dimension %d is not below the mesh dimension %d", bridge_dim, dim);
    return false;
  }
  pp.nranks = nranks;
  pp.rank = rank;
  pp.is_full_mesh = bm == FULL;
  // kept entities of every rank (the peers' are needed for the boundary lists)
  std::vector<std::vector<uint8_t>> keep_all[4];
  for (int d = 0; d <= dim; ++d) keep_all[d].resize((size_t)nranks);
  for (int s = 0; s < nranks; ++s) {
    std::vector<uint8_t> keep[4];
    kept_entities(w, s, keep);
    for (int d = 0; d <= dim; ++d) keep_all[d][(size_t)s].swap(keep[d]);
  }
  // numbering of the entities that stay: relative order of the full mesh
  std::vector<int32_t> ids[4], l2g[4];
  for (int d = 0; d <= dim; ++d) {
    ids[d].assign((size_t)full.nents[d], -1);
    for (int i = 0; i < full.nents[d]; ++i)
      if (keep_all[d][(size_t)rank][(size_t)i]) {
        ids[d][(size_t)i] = (int32_t)l2g[d].size();
        l2g[d].push_back(i);
      }
  }
  if (l2g[dim].empty()) {
    pp_set_error("constructPICPart: empty part on rank %d", rank);
    return false;
  }
  // the PICpart's own mesh
  HMesh& m = pp.mesh;
  m = HMesh();
  m.dim = dim;
  m.family = full.family;
  m.parting = 0;
  m.version = full.version;
  for (int d = 0; d <= dim; ++d) m.nents[d] = (int)l2g[d].size();
  for (int d = 1; d <= dim; ++d) {
    const int nd = d + 1;
    m.down[d].resize(l2g[d].size() * nd);
    if (d > 1) m.codes[d].resize(l2g[d].size() * nd);
    for (size_t i = 0; i < l2g[d].size(); ++i)
      for (int k = 0; k < nd; ++k) {
        const size_t src = (size_t)l2g[d][i] * nd + k;
        m.down[d][i * nd + k] = ids[d - 1][(size_t)full.down[d][src]];
        if (d > 1) m.codes[d][i * nd + k] = full.codes[d][src];
      }
  }
  if (!m.derive_verts()) return false;
  Sbars sb;
  build_sbars(w, sb);
  for (int d = 0; d <= dim; ++d) {
    for (const HTag& t : full.tags[d]) {
      if (t.name == "ownership" || t.name == "safe" || t.name == "gids" || t.name == "rank_lids" ||
          t.name == "sbar_id")
        continue;   // ours, written below
      convert_tag(full, m, d, l2g[d], t, t.name.c_str());
      // part_construct.cpp:203-206,238-240: the full-mesh numbering is kept as "global_serial"
      if (t.name == "global") convert_tag(full, m, d, l2g[d], t, "global_serial");
    }
    gather_tag<int32_t>(m, d, "ownership", PP_TAG_I32, w.owner[d], l2g[d]);
    if (d == dim) {
      std::vector<int32_t> safe(w.safe[(size_t)rank].begin(), w.safe[(size_t)rank].end());
      gather_tag<int32_t>(m, d, "safe", PP_TAG_I32, safe, l2g[d]);
    }
    gather_tag<int64_t>(m, d, "gids", PP_TAG_I64, w.gids[d], l2g[d]);
    gather_tag<int32_t>(m, d, "rank_lids", PP_TAG_I32, w.rank_lids[d], l2g[d]);
    if (d == dim) gather_tag<int32_t>(m, d, "sbar_id", PP_TAG_I32, sb.elem_sbar, l2g[d]);
  }
  for (int d = 0; d < 4; ++d) pp.d[d] = PicpartDim();
  for (int d = 0; d <= dim; ++d) setup_comm(w, rank, d, l2g[d], keep_all[d], pp.d[d]);
  store_sbar_table(sb.table[(size_t)rank], sb.max_sbar, pp);
  return true;
}

// ------------------------------------------------------------------ .ppm files
// Array compression of the .ppm stream: the reference's OMEGA_H_USE_ZLIB build switch
// (pumipic_file.cpp:76-80), here a run-time setting (pp_host_ppm_set_compression, or
// PUMIPIC_PPM_ZLIB=0 in the environment); default on, like the reference's usual build.
int g_ppm_compress = -1;
bool ppm_compression() {
  if (g_ppm_compress < 0) {
    const char* env = getenv("PUMIPIC_PPM_ZLIB");
    g_ppm_compress = (env && env[0] == '0') ? 0 : 1;
  }
  return g_ppm_compress != 0;
}
const char* split_path(const char* full_path) {
  const char* s = strrchr(full_path, '/');
  return s ? s + 1 : full_path;
}

bool write_ppm(const Picpart& pp, const char* prefix) {
  const std::string name = split_path(prefix);
  const std::string dir = std::string(prefix) + "_" + std::to_string(pp.nranks) + ".ppm";
  mkdir(dir.c_str(), 0777);
  struct stat st;
  if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) {
    pp_set_error("Failed to create directory %s", dir.c_str());
    return false;
  }
  const std::string base = dir + "/" + name + "_" + std::to_string(pp.rank);
  if (!write_osh(pp.mesh, (base + ".osh").c_str())) return false;
  Writer w;
  const bool comp = ppm_compression();      // the file carries no flag: it must match the reader's build
  w.value<int8_t>(2);                       // version
  w.value<int8_t>(pp.is_full_mesh ? 1 : 0);
  for (int i = 0; i < 4; ++i) {
    const PicpartDim& d = pp.d[i];
    w.value<int64_t>(d.num_entities);
    w.value<int32_t>(d.num_cores);
    w.array(d.buffered_parts.data(), (int64_t)d.buffered_parts.size(), 4, comp);
    w.array(d.offset_ents_per_rank.data(), (int64_t)d.offset_ents_per_rank.size(), 4, comp);
    w.array(d.ent_to_comm_arr_index.data(), (int64_t)d.ent_to_comm_arr_index.size(), 4, comp);
    w.array(d.is_complete_part.data(), (int64_t)d.is_complete_part.size(), 4, comp);
    w.value<int32_t>(d.num_bounds);
    w.value<int32_t>(d.num_boundaries);
    w.array(d.boundary_parts.data(), (int64_t)d.boundary_parts.size(), 4, comp);
    w.array(d.offset_bounded.data(), (int64_t)d.offset_bounded.size(), 4, comp);
    w.array(d.bounded_ent_ids.data(), (int64_t)d.bounded_ent_ids.size(), 4, comp);
  }
  if (!w.save((base + ".ppm").c_str())) {
    pp_set_error("Failed to open file %s.ppm", base.c_str());
    return false;
  }
  return true;
}

// The sbar table of a PICpart read from disk: the reference rebuilds its balancer with a new
// round of messages (pumipic_file.cpp:203); here the peers' `.osh` files in the same directory
// carry what those messages would (their "safe" and "gids" element tags).
void sbars_from_files(const std::string& dir, const std::string& name, Picpart& pp) {
  const int dim = pp.mesh.dim;
  const int32_t* sbar = pp.mesh.tag_data<int32_t>(dim, "sbar_id");
  const int32_t* own = pp.mesh.tag_data<int32_t>(dim, "ownership");
  const int32_t* safe = pp.mesh.tag_data<int32_t>(dim, "safe");
  const int64_t* gids = pp.mesh.tag_data<int64_t>(dim, "gids");
  if (!sbar || !own || !safe || !gids) return;
  const int ne = pp.mesh.nents[dim];
  std::vector<Parts> parts((size_t)ne);
  for (int e = 0; e < ne; ++e) {
    parts[(size_t)e].insert(own[e]);
    if (safe[e]) parts[(size_t)e].insert(pp.rank);
  }
  for (int b = 0; b < pp.nranks; ++b) {
    if (b == pp.rank) continue;
    HMesh peer;
    if (!read_osh((dir + "/" + name + "_" + std::to_string(b) + ".osh").c_str(), peer)) return;
    const int32_t* psafe = peer.tag_data<int32_t>(dim, "safe");
    const int64_t* pgid = peer.tag_data<int64_t>(dim, "gids");
    if (!psafe || !pgid) return;
    std::unordered_map<int64_t, int> where;
    for (int e = 0; e < peer.nents[dim]; ++e)
      if (psafe[e]) where[pgid[e]] = 1;
    for (int e = 0; e < ne; ++e)
      if (where.count(gids[e])) parts[(size_t)e].insert(b);
  }
  std::map<int, Parts> table;
  int max_id = 0;
  for (int e = 0; e < ne; ++e) {
    max_id = std::max(max_id, sbar[e] + (int)parts[(size_t)e].size());
    if (parts[(size_t)e].count(pp.rank)) table[sbar[e]] = parts[(size_t)e];
  }
  pp.sbar_ids.clear();
  pp.sbar_parts.clear();
  pp.sbar_parts_off.assign(1, 0);
  for (auto& kv : table) {
    pp.sbar_ids.push_back(kv.first);
    for (int p : kv.second) pp.sbar_parts.push_back(p);
    pp.sbar_parts_off.push_back((int32_t)pp.sbar_parts.size());
  }
  pp.max_sbar = max_id;   // lower bound: the largest id seen from this part
}

bool read_ppm(const char* prefix, int nranks, int rank, Picpart& pp) {
  const std::string name = split_path(prefix);
  const std::string dir = std::string(prefix) + "_" + std::to_string(nranks) + ".ppm";
  struct stat st;
  if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) {
    pp_set_error("Directory %s does not exist", dir.c_str());
    return false;
  }
  const std::string base = dir + "/" + name + "_" + std::to_string(rank);
  if (!read_osh((base + ".osh").c_str(), pp.mesh)) return false;
  Reader r;
  if (!r.load((base + ".ppm").c_str())) {
    pp_set_error("Cannot open file %s.ppm", base.c_str());
    return false;
  }
  // The reference compresses the arrays only when Omega_h was built with zlib
  // (pumipic_file.cpp:76-80) and the file does not say which: parse with the configured setting
  // first and, if the arrays do not decode, once more the other way.
  int8_t version = 0;
  auto parse = [&](bool comp) -> bool {
    r.pos = 0;
    r.ok = true;
    version = r.value<int8_t>();
    pp.is_full_mesh = r.value<int8_t>() != 0;
    if (!r.ok || version < 1 || version > 2) return false;
    for (int i = 0; i < 4 && r.ok; ++i) {
      PicpartDim& d = pp.d[i];
      d = PicpartDim();
      if (version >= 2) d.num_entities = r.value<int64_t>();
      d.num_cores = r.value<int32_t>();
      r.typed_array(comp, d.buffered_parts);
      r.typed_array(comp, d.offset_ents_per_rank);
      r.typed_array(comp, d.ent_to_comm_arr_index);
      r.typed_array(comp, d.is_complete_part);
      d.num_bounds = r.value<int32_t>();
      d.num_boundaries = r.value<int32_t>();
      r.typed_array(comp, d.boundary_parts);
      r.typed_array(comp, d.offset_bounded);
      r.typed_array(comp, d.bounded_ent_ids);
    }
    return r.ok && r.pos == r.buf.size();
  };
  pp.nranks = nranks;
  pp.rank = rank;
  const bool first = ppm_compression();
  if (!parse(first) && !parse(!first)) {
    if (r.buf.size() >= 1 && (r.buf[0] < 1 || r.buf[0] > 2))
      pp_set_error("%s.ppm: unsupported version %d", base.c_str(), (int)r.buf[0]);
    else
      pp_set_error("%s.ppm is truncated or corrupt", base.c_str());
    return false;
  }
  for (int i = 0; i <= pp.mesh.dim; ++i)
    if ((int)pp.d[i].offset_ents_per_rank.size() != nranks + 1 || (int)pp.d[i].is_complete_part.size() != nranks) {
      pp_set_error("%s.ppm was not written for %d ranks (dimension %d: %d offsets, %d completeness flags)",
                   base.c_str(), nranks, i, (int)pp.d[i].offset_ents_per_rank.size(),
                   (int)pp.d[i].is_complete_part.size());
      return false;
    }
  for (int i = 0; i <= pp.mesh.dim; ++i)
    if ((int)pp.d[i].ent_to_comm_arr_index.size() != pp.mesh.nents[i]) {
      pp_set_error("%s.ppm does not match its mesh (dimension %d: %d entries for %d entities)",
                   base.c_str(), i, (int)pp.d[i].ent_to_comm_arr_index.size(), pp.mesh.nents[i]);
      return false;
    }
  sbars_from_files(dir, name, pp);
  return true;
}

}  // namespace
}  // namespace pph

using pph::HMesh;
using pph::Picpart;

extern "C" void pp_host_ppm_set_compression(int32_t on) { pph::g_ppm_compress = on ? 1 : 0; }

extern "C" pp_status pp_host_picpart_build(const pp_host_mesh* full, const int32_t* elem_owner,
                                           int32_t nranks, int32_t rank, int32_t buffer_method,
                                           int32_t safe_method, int32_t buffer_layers,
                                           int32_t safe_layers, pp_host_picpart** out) {
  return pp_host_picpart_build_bridged(full, elem_owner, nranks, rank, buffer_method, safe_method,
                                       buffer_layers, safe_layers, 0, out);
}

extern "C" pp_status pp_host_picpart_build_bridged(const pp_host_mesh* full, const int32_t* elem_owner,
                                                   int32_t nranks, int32_t rank,
                                                   int32_t buffer_method, int32_t safe_method,
                                                   int32_t buffer_layers, int32_t safe_layers,
                                                   int32_t bridge_dim, pp_host_picpart** out) {
  const HMesh* f = reinterpret_cast<const HMesh*>(full);
  if (!f || !elem_owner || !out || nranks < 1 || rank < 0 || rank >= nranks || buffer_method < 0 ||
      buffer_method > 3 || safe_method < 0 || safe_method > 3 || !(f->dim == 2 || f->dim == 3) ||
      bridge_dim < 0 || bridge_dim >= f->dim) {
    pp_set_error("pp_host_picpart_build: bad argument");
    return PP_ERR_INVALID;
  }
  bool mine = false;
  for (int e = 0; e < f->nents[f->dim]; ++e) {
    if (elem_owner[e] < 0 || elem_owner[e] >= nranks) {
      pp_set_error("pp_host_picpart_build: element %d has owner %d outside [0,%d)", e, elem_owner[e], nranks);
      return PP_ERR_INVALID;
    }
    mine = mine || elem_owner[e] == rank;
  }
  if (!mine) {   // setOwnerByClassification / constructPICPart assert on this
    pp_set_error("pp_host_picpart_build: rank %d with no owned elements detected", rank);
    return PP_ERR_INVALID;
  }
  if (buffer_layers < 0) buffer_layers = 3;   // pumipic_input.cpp:103-105
  if (safe_layers < 0) safe_layers = 1;
  Picpart* pp = new Picpart();
  if (!pph::build_picpart(*f, elem_owner, nranks, rank, buffer_method, safe_method, buffer_layers,
                          safe_layers, bridge_dim, *pp)) {
    delete pp;
    return PP_ERR_INVALID;
  }
  *out = reinterpret_cast<pp_host_picpart*>(pp);
  return PP_OK;
}

extern "C" void pp_host_picpart_destroy(pp_host_picpart* pp) { delete reinterpret_cast<Picpart*>(pp); }

extern "C" const pp_host_mesh* pp_host_picpart_mesh(const pp_host_picpart* pp) {
  return pp ? reinterpret_cast<const pp_host_mesh*>(&reinterpret_cast<const Picpart*>(pp)->mesh) : nullptr;
}

extern "C" pp_status pp_host_picpart_get(const pp_host_picpart* pp_, int32_t d,
                                         pp_host_picpart_dim* out) {
  const Picpart* pp = reinterpret_cast<const Picpart*>(pp_);
  if (!pp || !out || d < 0 || d > 3) {
    pp_set_error("pp_host_picpart_get: bad argument");
    return PP_ERR_INVALID;
  }
  const pph::PicpartDim& s = pp->d[d];
  out->num_entities = s.num_entities;
  out->nents = (int32_t)s.ent_to_comm_arr_index.size();
  out->num_cores = s.num_cores;
  out->buffered_parts = s.buffered_parts.data();
  out->offset_ents_per_rank = s.offset_ents_per_rank.data();
  out->ent_to_comm_arr_index = s.ent_to_comm_arr_index.data();
  out->is_complete_part = s.is_complete_part.data();
  out->num_bounds = s.num_bounds;
  out->num_boundaries = s.num_boundaries;
  out->boundary_parts = s.boundary_parts.data();
  out->offset_bounded = s.offset_bounded.data();
  out->n_offset_bounded = (int32_t)s.offset_bounded.size();
  out->bounded_ent_ids = s.bounded_ent_ids.data();
  out->n_bounded_ent_ids = (int32_t)s.bounded_ent_ids.size();
  out->ent_l2g = s.ent_l2g.empty() ? nullptr : s.ent_l2g.data();
  return PP_OK;
}

extern "C" int32_t pp_host_picpart_is_full_mesh(const pp_host_picpart* pp) {
  return pp ? (int32_t)reinterpret_cast<const Picpart*>(pp)->is_full_mesh : -1;
}
extern "C" int32_t pp_host_picpart_nranks(const pp_host_picpart* pp) {
  return pp ? reinterpret_cast<const Picpart*>(pp)->nranks : -1;
}
extern "C" int32_t pp_host_picpart_rank(const pp_host_picpart* pp) {
  return pp ? reinterpret_cast<const Picpart*>(pp)->rank : -1;
}

extern "C" pp_status pp_host_picpart_write(const pp_host_picpart* pp, const char* prefix) {
  if (!pp || !prefix) {
    pp_set_error("pp_host_picpart_write: bad argument");
    return PP_ERR_INVALID;
  }
  return pph::write_ppm(*reinterpret_cast<const Picpart*>(pp), prefix) ? PP_OK : PP_ERR_INVALID;
}

extern "C" pp_status pp_host_picpart_read(const char* prefix, int32_t nranks, int32_t rank,
                                          pp_host_picpart** out) {
  if (!prefix || !out || nranks < 1 || rank < 0 || rank >= nranks) {
    pp_set_error("pp_host_picpart_read: bad argument");
    return PP_ERR_INVALID;
  }
  Picpart* pp = new Picpart();
  if (!pph::read_ppm(prefix, nranks, rank, *pp)) {
    delete pp;
    return PP_ERR_INVALID;
  }
  *out = reinterpret_cast<pp_host_picpart*>(pp);
  return PP_OK;
}

extern "C" pp_status pp_host_picpart_sbars(const pp_host_picpart* pp_, int32_t* nsbars,
                                           const int32_t** sbar_ids, const int32_t** parts_off,
                                           const int32_t** parts, int32_t* max_sbar) {
  const Picpart* pp = reinterpret_cast<const Picpart*>(pp_);
  if (!pp || !nsbars || !sbar_ids || !parts_off || !parts) {
    pp_set_error("pp_host_picpart_sbars: bad argument");
    return PP_ERR_INVALID;
  }
  *nsbars = (int32_t)pp->sbar_ids.size();
  *sbar_ids = pp->sbar_ids.data();
  *parts_off = pp->sbar_parts_off.data();
  *parts = pp->sbar_parts.data();
  if (max_sbar) *max_sbar = pp->max_sbar;
  return PP_OK;
}
