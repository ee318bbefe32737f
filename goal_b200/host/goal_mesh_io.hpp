// goal_mesh_io.hpp -- reads the reference's mesh inputs into the flat arrays gx_desc wants, without PUMI.
//
// The reference loads its meshes through SCOREC/core (Disc::Disc, src/goal_disc.cpp:31-62: apf::loadMdsMesh of a
// .dmg geometric model + per-part .smb files, then an "assoc file" that names node / side / elem sets by model
// entity, src/goal_disc.cpp:64-93, 334-396).  SCOREC/core is not available here, so this header restates the
// on-disk formats as decoded for SURVEY.md 8(c):
//
//   .smb  (MDS "version 5", big-endian)
//         u32 magic = 0, version, dim, nparts; u32 counts[8] (vert, edge, tri, quad, hex, prism, pyramid, tet);
//         edge->vert (2 u32 each), tri->edge (3), tet->tri (4); coords nv*3 f64; params nv*2 f64;
//         remotes: u32 np, peers[np], counts[np], then per peer the local ids of the shared vertices (the same
//         order on both sides of a part boundary); then (model tag, model dim) u32 pairs per entity.
//   .dmg  text: "nregion nface nedge nvertex", bounding box (6), vertices (tag x y z), edges (tag v0 v1), faces
//         (tag nloops {nedges {edge dir}}) ...
//   assoc text: "<node set|side set|elem set> <name> <n>" followed by n lines "<model dim> <model tag>".
//
// A vertex belongs to a node set when its classification lies in the closure of one of the set's model entities
// (Disc::compute_node_sets, src/goal_disc.cpp:369-396); a boundary triangle / an element belongs to a side / elem
// set when it is classified on one of the set's model entities (compute_side_sets / compute_elem_sets, :334-367).
// Tet vertex order: the sets come from the downward adjacencies; the canonical MDS order is not reproduced (a
// permutation of an element's nodes changes nothing but round-off) -- each tet is oriented to positive volume.
#ifndef GOAL_MESH_IO_HPP
#define GOAL_MESH_IO_HPP

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gx {

struct SmbPart {
  int dim = 0, nparts = 0;
  std::vector<double> coords;                 // [nv*3]
  std::vector<int32_t> tets;                  // [ntet*4], positively oriented
  std::vector<int32_t> tris;                  // [ntri*3]
  std::vector<int32_t> peers;                 // neighbouring parts
  std::vector<std::vector<int32_t>> remotes;  // per peer: local ids of the shared vertices, agreed order
  // classification (model dim, model tag) per entity
  std::vector<std::pair<int, int>> vert_cls, tri_cls, tet_cls;
  int32_t n_verts() const { return (int32_t)(coords.size() / 3); }
  int32_t n_tets() const { return (int32_t)(tets.size() / 4); }
  int32_t n_tris() const { return (int32_t)(tris.size() / 3); }
};

namespace detail {
struct Reader {
  std::vector<unsigned char> b;
  size_t off = 0;
  explicit Reader(std::string const& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    b.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  }
  uint32_t u32() {
    if (off + 4 > b.size()) throw std::runtime_error("smb: truncated file");
    uint32_t v = ((uint32_t)b[off] << 24) | ((uint32_t)b[off + 1] << 16) | ((uint32_t)b[off + 2] << 8) | (uint32_t)b[off + 3];
    off += 4;
    return v;
  }
  double f64() {
    if (off + 8 > b.size()) throw std::runtime_error("smb: truncated file");
    uint64_t v = 0;
    for (int k = 0; k < 8; ++k) v = (v << 8) | b[off + k];
    off += 8;
    double d;
    std::memcpy(&d, &v, 8);
    return d;
  }
};
}  // namespace detail

inline SmbPart read_smb(std::string const& path) {
  detail::Reader r(path);
  SmbPart m;
  uint32_t const magic = r.u32(), version = r.u32();
  m.dim = (int)r.u32();
  m.nparts = (int)r.u32();
  uint32_t cnt[8];
  for (auto& c : cnt) c = r.u32();
  if (magic != 0 || version != 5) throw std::runtime_error("smb: unsupported magic/version in " + path);
  if (m.dim != 3 || cnt[3] || cnt[4] || cnt[5] || cnt[6]) throw std::runtime_error("smb: only 3D all-tet meshes are supported");
  uint32_t const nv = cnt[0], ne = cnt[1], nt = cnt[2], ntet = cnt[7];
  std::vector<std::array<uint32_t, 2>> e2v(ne);
  for (auto& e : e2v) { e[0] = r.u32(); e[1] = r.u32(); }
  std::vector<std::array<uint32_t, 3>> t2e(nt);
  for (auto& t : t2e) for (auto& x : t) x = r.u32();
  std::vector<std::array<uint32_t, 4>> tet2t(ntet);
  for (auto& t : tet2t) for (auto& x : t) x = r.u32();
  m.coords.resize(3 * (size_t)nv);
  for (auto& x : m.coords) x = r.f64();
  for (size_t k = 0; k < 2 * (size_t)nv; ++k) r.f64();  // parametric coordinates, unused
  uint32_t const np = r.u32();
  m.peers.resize(np);
  for (auto& p : m.peers) p = (int32_t)r.u32();
  std::vector<uint32_t> pc(np);
  for (auto& c : pc) c = r.u32();
  m.remotes.resize(np);
  for (uint32_t p = 0; p < np; ++p) {
    m.remotes[p].resize(pc[p]);
    for (auto& v : m.remotes[p]) v = (int32_t)r.u32();
  }
  auto read_cls = [&](uint32_t n, std::vector<std::pair<int, int>>* out) {
    for (uint32_t k = 0; k < n; ++k) {
      int const tag = (int)r.u32(), dim = (int)r.u32();
      if (out) out->push_back({dim, tag});
    }
  };
  read_cls(nv, &m.vert_cls);
  read_cls(ne, nullptr);
  read_cls(nt, &m.tri_cls);
  read_cls(ntet, &m.tet_cls);
  // triangles and tets as vertex sets
  m.tris.resize(3 * (size_t)nt);
  for (uint32_t t = 0; t < nt; ++t) {
    auto const &a = e2v.at(t2e[t][0]), &c = e2v.at(t2e[t][1]);
    uint32_t shared = (a[0] == c[0] || a[0] == c[1]) ? a[0] : a[1];
    if (shared != c[0] && shared != c[1]) throw std::runtime_error("smb: triangle edges do not meet");
    m.tris[3 * t] = (int32_t)(a[0] == shared ? a[1] : a[0]);
    m.tris[3 * t + 1] = (int32_t)shared;
    m.tris[3 * t + 2] = (int32_t)(c[0] == shared ? c[1] : c[0]);
  }
  m.tets.resize(4 * (size_t)ntet);
  for (uint32_t k = 0; k < ntet; ++k) {
    int32_t v[4];
    for (int j = 0; j < 3; ++j) v[j] = m.tris.at(3 * (size_t)tet2t[k][0] + j);
    v[3] = -1;
    for (int j = 0; j < 3; ++j) {
      int32_t const w = m.tris.at(3 * (size_t)tet2t[k][1] + j);
      if (w != v[0] && w != v[1] && w != v[2]) v[3] = w;
    }
    if (v[3] < 0) throw std::runtime_error("smb: tet faces do not close");
    double e[3][3];
    for (int a = 0; a < 3; ++a) for (int j = 0; j < 3; ++j) e[a][j] = m.coords[3 * (size_t)v[a + 1] + j] - m.coords[3 * (size_t)v[0] + j];
    double const det = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) - e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                       e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
    if (det < 0) std::swap(v[1], v[2]);
    for (int j = 0; j < 4; ++j) m.tets[4 * (size_t)k + j] = v[j];
  }
  return m;
}

// geometric model: closure of a model face = its edges and their vertices
struct DmgModel {
  std::map<int, std::pair<int, int>> edges;   // tag -> (v0, v1)
  std::map<int, std::vector<int>> faces;      // tag -> edge tags
  std::set<std::pair<int, int>> closure(int dim, int tag) const {
    std::set<std::pair<int, int>> out{{dim, tag}};
    if (dim == 2) {
      auto it = faces.find(tag);
      if (it != faces.end())
        for (int e : it->second) {
          out.insert({1, e});
          auto ed = edges.find(e);
          if (ed != edges.end()) { out.insert({0, ed->second.first}); out.insert({0, ed->second.second}); }
        }
    } else if (dim == 1) {
      auto ed = edges.find(tag);
      if (ed != edges.end()) { out.insert({0, ed->second.first}); out.insert({0, ed->second.second}); }
    }
    return out;
  }
};

inline DmgModel read_dmg(std::string const& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  DmgModel g;
  long nreg, nface, nedge, nvert;
  f >> nreg >> nface >> nedge >> nvert;
  double d;
  for (int k = 0; k < 6; ++k) f >> d;
  for (long k = 0; k < nvert; ++k) { long tag; f >> tag >> d >> d >> d; }
  for (long k = 0; k < nedge; ++k) { int tag, a, b; f >> tag >> a >> b; g.edges[tag] = {a, b}; }
  for (long k = 0; k < nface; ++k) {
    int tag, nloops;
    f >> tag >> nloops;
    auto& fe = g.faces[tag];
    for (int l = 0; l < nloops; ++l) {
      int n;
      f >> n;
      for (int j = 0; j < n; ++j) { int e, dir; f >> e >> dir; fe.push_back(e); }
    }
  }
  if (!f) throw std::runtime_error("dmg: parse error in " + path);
  return g;
}

// assoc file: kind ("node set" / "side set" / "elem set") -> name -> model entities (dim, tag)
using AssocSets = std::map<std::string, std::map<std::string, std::vector<std::pair<int, int>>>>;
inline AssocSets read_assoc(std::string const& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open " + path);
  AssocSets s;
  std::string line;
  while (std::getline(f, line)) {
    if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
    std::string const kind = line.substr(0, 8);
    std::istringstream is(line.substr(8));
    std::string name;
    int n = 0;
    is >> name >> n;
    auto& ents = s[kind][name];
    for (int k = 0; k < n; ++k) {
      if (!std::getline(f, line)) throw std::runtime_error("assoc: truncated set " + name);
      std::istringstream es(line);
      int dim, tag;
      es >> dim >> tag;
      ents.push_back({dim, tag});
    }
  }
  return s;
}

// Disc::compute_node_sets / compute_side_sets / compute_elem_sets (src/goal_disc.cpp:334-396) on one part
struct MeshSets {
  std::map<std::string, std::vector<int32_t>> node_sets;  // vertex ids
  std::map<std::string, std::vector<int32_t>> side_sets;  // [n*3] vertex ids of the boundary triangles
  std::map<std::string, std::vector<int32_t>> elem_sets;  // tet ids
};
inline MeshSets make_sets(SmbPart const& m, DmgModel const& g, AssocSets const& a) {
  MeshSets out;
  auto find = [&](const char* kind) -> std::map<std::string, std::vector<std::pair<int, int>>> const* {
    auto it = a.find(kind);
    return it == a.end() ? nullptr : &it->second;
  };
  if (auto ns = find("node set"))
    for (auto const& kv : *ns) {
      std::set<std::pair<int, int>> cl;
      for (auto const& e : kv.second) { auto c = g.closure(e.first, e.second); cl.insert(c.begin(), c.end()); }
      auto& v = out.node_sets[kv.first];
      for (int32_t i = 0; i < m.n_verts(); ++i) if (cl.count(m.vert_cls[i])) v.push_back(i);
    }
  if (auto ss = find("side set"))
    for (auto const& kv : *ss) {
      std::set<std::pair<int, int>> want(kv.second.begin(), kv.second.end());
      auto& v = out.side_sets[kv.first];
      for (int32_t t = 0; t < m.n_tris(); ++t)
        if (want.count(m.tri_cls[t])) v.insert(v.end(), {m.tris[3 * t], m.tris[3 * t + 1], m.tris[3 * t + 2]});
    }
  if (auto es = find("elem set"))
    for (auto const& kv : *es) {
      std::set<std::pair<int, int>> want(kv.second.begin(), kv.second.end());
      auto& v = out.elem_sets[kv.first];
      for (int32_t e = 0; e < m.n_tets(); ++e) if (want.count(m.tet_cls[e])) v.push_back(e);
    }
  return out;
}

}  // namespace gx
#endif
