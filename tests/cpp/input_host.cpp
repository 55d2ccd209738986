// input_host.cpp -- host-only check (plain g++, no GPU): the C++ mirror header compiles without nvcc and
// pumipic::Input applies the reference constructor's adjustments (pumipic_input.cpp:94-110).
#include <cassert>
#include <cstdio>
#include <fstream>
#include "pumipic_b200.hpp"
namespace p = pumipic;
int main() {
  double* coords; int32_t* ev; int32_t nv, ne;
  if (pp_host_plate(6, 1.0, &nv, &coords, &ne, &ev) != PP_OK) return 2;
  pp_host_mesh* full = nullptr;
  if (pp_host_mesh_from_elems(2, nv, coords, ne, ev, &full) != PP_OK) return 3;
  std::vector<int> cls((size_t)ne);
  for (int e = 0; e < ne; ++e) cls[e] = e % 3;
  pp_host_mesh_set_tag(full, 2, "class_id", 1, PP_TAG_I32, cls.data());
  p::Input a(full, p::Input::CLASSIFICATION, std::vector<int>{2, 0, 1}, p::Input::NONE, p::Input::MINIMUM);
  assert(a.bufferBFSLayers == 0 && a.safeBFSLayers == 0 && a.getRule() == p::Input::CLASSIFICATION);
  a.printMethod();
  std::printf("ok %d elems\n", ne);
  return 0;
}
