// gx_meshdump -- prints what goal_mesh_io.hpp reads from a PUMI mesh part as JSON (used by tests/test_mesh_io.py
// and handy for turning the reference's meshes into gx_desc inputs):  gx_meshdump part.smb [model.dmg assoc.txt]
#include <cstdio>

#include "goal_mesh_io.hpp"

template <class T> static void arr(const char* name, std::vector<T> const& v, const char* fmt, bool last = false) {
  std::printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) { if (i) std::printf(","); std::printf(fmt, v[i]); }
  std::printf("]%s", last ? "" : ",\n");
}
static void sets(const char* name, std::map<std::string, std::vector<int32_t>> const& s, bool last = false) {
  std::printf("\"%s\": {", name);
  bool first = true;
  for (auto const& kv : s) {
    if (!first) std::printf(",");
    first = false;
    std::printf("\"%s\": [", kv.first.c_str());
    for (size_t i = 0; i < kv.second.size(); ++i) std::printf("%s%d", i ? "," : "", kv.second[i]);
    std::printf("]");
  }
  std::printf("}%s", last ? "" : ",\n");
}

int main(int argc, char** argv) {
  if (argc != 2 && argc != 4) { std::fprintf(stderr, "usage: %s part.smb [model.dmg assoc.txt]\n", argv[0]); return 2; }
  try {
    gx::SmbPart m = gx::read_smb(argv[1]);
    std::printf("{\"dim\": %d, \"nparts\": %d,\n", m.dim, m.nparts);
    arr("coords", m.coords, "%.17g");
    arr("tets", m.tets, "%d");
    arr("tris", m.tris, "%d");
    arr("peers", m.peers, "%d");
    std::printf("\"remotes\": [");
    for (size_t p = 0; p < m.remotes.size(); ++p) {
      std::printf("%s[", p ? "," : "");
      for (size_t i = 0; i < m.remotes[p].size(); ++i) std::printf("%s%d", i ? "," : "", m.remotes[p][i]);
      std::printf("]");
    }
    std::printf("]");
    if (argc == 4) {
      gx::MeshSets s = gx::make_sets(m, gx::read_dmg(argv[2]), gx::read_assoc(argv[3]));
      std::printf(",\n");
      sets("node_sets", s.node_sets);
      sets("side_sets", s.side_sets);
      sets("elem_sets", s.elem_sets, true);
    }
    std::printf("}\n");
  } catch (std::exception const& e) {
    std::fprintf(stderr, "gx_meshdump: %s\n", e.what());
    return 1;
  }
  return 0;
}
