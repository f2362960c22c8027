// mesh_api.cpp -- C ABI of the host-side mesh construction (include/gorilla_b200.h, last section).
// Replaces the grid/physics half of initialize_gorilla (SRC/orbit_timestep_gorilla.f90:151-274).
#include "mesh_common.hpp"
#include <new>

struct gorilla_mesh {
  gbhost::Mesh m;
};

extern "C" const char *gorilla_b200_last_error(void);
namespace gbhost { void set_last_error(const std::string &s); }

extern "C" int gorilla_mesh_build(const gorilla_grid_settings *grid, const gorilla_settings *settings, gorilla_mesh **out)
{
  if (!grid || !settings || !out) { gbhost::set_last_error("gorilla_mesh_build: null argument"); return GORILLA_ERR_ARG; }
  gorilla_mesh *gm = new (std::nothrow) gorilla_mesh();
  if (!gm) return GORILLA_ERR_ARG;
  std::string err;
  int rc = gbhost::set_species(gm->m, settings->ispecies, err);
  if (rc == GORILLA_OK) {
    switch (grid->grid_kind) {
      case 5: rc = gbhost::build_analytic_circ(*grid, *settings, gm->m, err); break;
      case 3: rc = gbhost::build_vmec(*grid, *settings, gm->m, err); break;
      case 1: rc = gbhost::build_efit_rect(*grid, *settings, gm->m, err); break;
      case 2: rc = gbhost::build_efit_flux(*grid, *settings, gm->m, err); break;
      case 4: rc = gbhost::build_soledge3x(*grid, *settings, gm->m, err); break;
      default:
        err = "grid_kind must be 1 (EFIT, rectangular), 2 (EFIT, field aligned), 3 (VMEC, field aligned), 4 (SOLEDGE3X-EIRENE) "
              "or 5 (analytic circular tokamak)";
        rc = GORILLA_ERR_UNSUPPORTED;
    }
  }
  if (rc != GORILLA_OK) {
    gbhost::set_last_error(err);
    delete gm;
    return rc;
  }
  *out = gm;
  return GORILLA_OK;
}

extern "C" int gorilla_mesh_get_desc(const gorilla_mesh *mesh, gorilla_mesh_desc *d)
{
  if (!mesh || !d) return GORILLA_ERR_ARG;
  const gbhost::Mesh &m = mesh->m;
  d->ntetr = m.ntetr;
  d->tetra_physics = m.tetra_physics.data();
  d->tetra_grid = m.tetra_grid.data();
  d->cm_over_e = m.cm_over_e;
  d->particle_mass = m.particle_mass;
  d->particle_charge = m.particle_charge;
  d->sign_sqg = m.sign_sqg;
  d->coord_system = m.coord_system;
  d->n_field_periods = m.n_field_periods;
  d->grid_kind = m.grid_kind;
  for (int i = 0; i < 3; i++) d->grid_size[i] = m.grid_size[i];
  d->pad0 = 0;
  d->Rmin = m.Rmin; d->Rmax = m.Rmax; d->Zmin = m.Zmin; d->Zmax = m.Zmax;
  d->sfc_s_min = m.sfc_s_min;
  return GORILLA_OK;
}

extern "C" int gorilla_mesh_get_vertices(const gorilla_mesh *mesh, int64_t *nvert, const double **verts_rphiz,
                                         const double **verts_sthetaphi)
{
  if (!mesh) return GORILLA_ERR_ARG;
  if (nvert) *nvert = mesh->m.nvert;
  if (verts_rphiz) *verts_rphiz = mesh->m.verts_rphiz.data();
  if (verts_sthetaphi) *verts_sthetaphi = mesh->m.verts_sthetaphi.empty() ? nullptr : mesh->m.verts_sthetaphi.data();
  return GORILLA_OK;
}

extern "C" void gorilla_mesh_free(gorilla_mesh *mesh) { delete mesh; }
