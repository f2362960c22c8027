// mesh_vmec.cpp -- grid_kind = 3 (placeholder until the VMEC pipeline lands)
#include "mesh_common.hpp"
namespace gbhost {
int build_vmec(const gorilla_grid_settings &, const gorilla_settings &, Mesh &, std::string &err)
{
  err = "grid_kind 3 (VMEC) mesh builder not implemented yet";
  return GORILLA_ERR_UNSUPPORTED;
}
}
