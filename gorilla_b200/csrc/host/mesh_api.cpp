// mesh_api.cpp -- C ABI of the host-side mesh construction (include/gorilla_b200.h, last section).
// Replaces the grid/physics half of initialize_gorilla (SRC/orbit_timestep_gorilla.f90:151-274).
#include "mesh_common.hpp"
#include <new>
#include <cstdio>
#include <cstring>

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
  gm->m.handover_processing_kind = settings->handover_processing_kind;
  gm->m.bmod_multiplier = grid->bmod_multiplier != 0.0 ? grid->bmod_multiplier : 1.0;
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

extern "C" int gorilla_b200_abi_struct_sizes(int64_t sizes_out[GORILLA_ABI_N_STRUCTS])
{
  if (!sizes_out) return GORILLA_ERR_ARG;
  const size_t sz[GORILLA_ABI_N_STRUCTS] = {sizeof(gorilla_settings), sizeof(gorilla_mesh_desc), sizeof(gorilla_counters),
                                            sizeof(gorilla_diag), sizeof(gorilla_grid_settings), sizeof(gorilla_event),
                                            sizeof(gorilla_event_settings)};
  for (int i = 0; i < GORILLA_ABI_N_STRUCTS; i++) sizes_out[i] = (int64_t)sz[i];
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
  d->tetra_skew_coord = m.tetra_skew_coord.empty() ? nullptr : m.tetra_skew_coord.data();
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

// ---- .gmesh: versioned on-disk form of a host mesh -----------------------------------------
// The reference has no mesh file: every run rebuilds tetra_grid / tetra_physics in initialize_gorilla.  A .gmesh file is
// those two arrays as they sit in memory (the `sequence` derived types, [ntetr][142] doubles and [ntetr][20] int32), the
// vertex tables and the module scalars the hot path reads, behind a fixed little-endian header:
//   magic "GMESHB2\0" | u32 version | u32 endian tag 0x01020304 | u32 ndoubles per record (142) | u32 nints (20)
//   | i64 ntetr | i64 nvert | i32 flags | i32 sign_sqg, coord_system, n_field_periods, grid_kind, grid_size[3]
//   | f64 cm_over_e, particle_mass, particle_charge, Rmin, Rmax, Zmin, Zmax, sfc_s_min | u64 FNV-1a of the payload
// followed by the payload: tetra_physics, tetra_grid, verts_rphiz, [verts_sthetaphi], [tetra_skew_coord].
// flags: bit 0 = verts_sthetaphi present, bit 1 (version 2) = tetra_skew_coord present ([ntetr][168] doubles, the records
// of handover_processing_kind = 2).  Version 1 files (no skew records) are still read.
namespace {
const char GMESH_MAGIC[8] = {'G', 'M', 'E', 'S', 'H', 'B', '2', '\0'};
const uint32_t GMESH_VERSION = 2, GMESH_ENDIAN = 0x01020304u;
const int32_t GMESH_F_STHETAPHI = 1, GMESH_F_SKEW = 2;
struct GmeshHeader {
  char magic[8];
  uint32_t version, endian, ndoubles, nints;
  int64_t ntetr, nvert;
  int32_t has_sthetaphi, sign_sqg, coord_system, n_field_periods, grid_kind, grid_size[3];
  double cm_over_e, particle_mass, particle_charge, Rmin, Rmax, Zmin, Zmax, sfc_s_min;
  uint64_t checksum;
};
uint64_t fnv1a(uint64_t h, const void *p, size_t n)
{
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}
int io_fail(const std::string &msg, FILE *f = nullptr)
{
  if (f) fclose(f);
  gbhost::set_last_error(msg);
  return GORILLA_ERR_IO;
}
} // namespace

extern "C" int gorilla_mesh_save(const gorilla_mesh_desc *d, int64_t nvert, const double *verts_rphiz,
                                 const double *verts_sthetaphi, const char *path)
{
  if (!d || !path || !d->tetra_physics || !d->tetra_grid || d->ntetr < 1 || nvert < 0 || (nvert > 0 && !verts_rphiz)) {
    gbhost::set_last_error("gorilla_mesh_save: null argument or empty mesh");
    return GORILLA_ERR_ARG;
  }
  GmeshHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, GMESH_MAGIC, 8);
  h.version = GMESH_VERSION; h.endian = GMESH_ENDIAN;
  h.ndoubles = GORILLA_TETRA_PHYSICS_NDOUBLES; h.nints = GORILLA_TETRA_GRID_NINTS;
  h.ntetr = d->ntetr; h.nvert = nvert;
  const bool has_sth = verts_sthetaphi && nvert > 0, has_skew = d->tetra_skew_coord != nullptr;
  h.has_sthetaphi = (has_sth ? GMESH_F_STHETAPHI : 0) | (has_skew ? GMESH_F_SKEW : 0);
  h.sign_sqg = d->sign_sqg; h.coord_system = d->coord_system; h.n_field_periods = d->n_field_periods; h.grid_kind = d->grid_kind;
  for (int i = 0; i < 3; i++) h.grid_size[i] = d->grid_size[i];
  h.cm_over_e = d->cm_over_e; h.particle_mass = d->particle_mass; h.particle_charge = d->particle_charge;
  h.Rmin = d->Rmin; h.Rmax = d->Rmax; h.Zmin = d->Zmin; h.Zmax = d->Zmax; h.sfc_s_min = d->sfc_s_min;
  const size_t n_tp = (size_t)d->ntetr * h.ndoubles * sizeof(double), n_tg = (size_t)d->ntetr * h.nints * sizeof(int32_t),
               n_v = (size_t)nvert * 3 * sizeof(double),
               n_sk = (size_t)d->ntetr * GORILLA_TETRA_SKEW_NDOUBLES * sizeof(double);
  uint64_t c = 14695981039346656037ull;
  c = fnv1a(c, d->tetra_physics, n_tp);
  c = fnv1a(c, d->tetra_grid, n_tg);
  if (nvert > 0) c = fnv1a(c, verts_rphiz, n_v);
  if (has_sth) c = fnv1a(c, verts_sthetaphi, n_v);
  if (has_skew) c = fnv1a(c, d->tetra_skew_coord, n_sk);
  h.checksum = c;
  FILE *f = fopen(path, "wb");
  if (!f) return io_fail(std::string("gorilla_mesh_save: cannot open ") + path);
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(d->tetra_physics, 1, n_tp, f) == n_tp &&
            fwrite(d->tetra_grid, 1, n_tg, f) == n_tg && (nvert == 0 || fwrite(verts_rphiz, 1, n_v, f) == n_v) &&
            (!has_sth || fwrite(verts_sthetaphi, 1, n_v, f) == n_v) &&
            (!has_skew || fwrite(d->tetra_skew_coord, 1, n_sk, f) == n_sk);
  if (fclose(f) != 0) ok = false;
  if (!ok) return io_fail(std::string("gorilla_mesh_save: short write to ") + path);
  return GORILLA_OK;
}

extern "C" int gorilla_mesh_load(const char *path, gorilla_mesh **out)
{
  if (!path || !out) { gbhost::set_last_error("gorilla_mesh_load: null argument"); return GORILLA_ERR_ARG; }
  FILE *f = fopen(path, "rb");
  if (!f) return io_fail(std::string("gorilla_mesh_load: cannot open ") + path);
  GmeshHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1) return io_fail("gorilla_mesh_load: file shorter than the .gmesh header", f);
  if (memcmp(h.magic, GMESH_MAGIC, 8) != 0) return io_fail("gorilla_mesh_load: not a .gmesh file (bad magic)", f);
  if (h.endian != GMESH_ENDIAN) return io_fail("gorilla_mesh_load: file written with the other byte order", f);
  if (h.version != GMESH_VERSION && h.version != 1)
    return io_fail("gorilla_mesh_load: unsupported .gmesh version " + std::to_string(h.version) + " (this build reads version " +
                   std::to_string(GMESH_VERSION) + ")", f);
  if (h.ndoubles != GORILLA_TETRA_PHYSICS_NDOUBLES || h.nints != GORILLA_TETRA_GRID_NINTS)
    return io_fail("gorilla_mesh_load: record sizes differ from tetrahedron_physics (142 doubles) / tetrahedron_grid (20 ints)", f);
  if (h.ntetr < 1 || h.nvert < 0 || h.ntetr > (int64_t)1 << 40 || h.nvert > (int64_t)1 << 40)
    return io_fail("gorilla_mesh_load: implausible sizes in the header", f);
  gorilla_mesh *gm = new (std::nothrow) gorilla_mesh();
  if (!gm) return io_fail("gorilla_mesh_load: out of memory", f);
  gbhost::Mesh &m = gm->m;
  m.ntetr = h.ntetr; m.nvert = h.nvert;
  m.sign_sqg = h.sign_sqg; m.coord_system = h.coord_system; m.n_field_periods = h.n_field_periods; m.grid_kind = h.grid_kind;
  for (int i = 0; i < 3; i++) m.grid_size[i] = h.grid_size[i];
  m.cm_over_e = h.cm_over_e; m.particle_mass = h.particle_mass; m.particle_charge = h.particle_charge;
  m.Rmin = h.Rmin; m.Rmax = h.Rmax; m.Zmin = h.Zmin; m.Zmax = h.Zmax; m.sfc_s_min = h.sfc_s_min;
  bool ok = true;
  const bool has_sth = (h.has_sthetaphi & GMESH_F_STHETAPHI) != 0;
  const bool has_skew = h.version >= 2 && (h.has_sthetaphi & GMESH_F_SKEW) != 0;
  if (h.has_sthetaphi & ~(h.version >= 2 ? (GMESH_F_STHETAPHI | GMESH_F_SKEW) : GMESH_F_STHETAPHI)) {
    delete gm;
    return io_fail("gorilla_mesh_load: unknown payload flags in the header", f);
  }
  m.handover_processing_kind = has_skew ? 2 : 1;
  try {
    m.tetra_physics.resize((size_t)h.ntetr * h.ndoubles);
    m.tetra_grid.resize((size_t)h.ntetr * h.nints);
    m.verts_rphiz.resize((size_t)h.nvert * 3);
    if (has_sth) m.verts_sthetaphi.resize((size_t)h.nvert * 3);
    if (has_skew) m.tetra_skew_coord.resize((size_t)h.ntetr * GORILLA_TETRA_SKEW_NDOUBLES);
  } catch (...) {
    ok = false;
  }
  const size_t n_tp = m.tetra_physics.size() * sizeof(double), n_tg = m.tetra_grid.size() * sizeof(int32_t),
               n_v = m.verts_rphiz.size() * sizeof(double), n_sk = m.tetra_skew_coord.size() * sizeof(double);
  ok = ok && fread(m.tetra_physics.data(), 1, n_tp, f) == n_tp && fread(m.tetra_grid.data(), 1, n_tg, f) == n_tg &&
       (n_v == 0 || fread(m.verts_rphiz.data(), 1, n_v, f) == n_v) &&
       (!has_sth || fread(m.verts_sthetaphi.data(), 1, n_v, f) == n_v) &&
       (!has_skew || fread(m.tetra_skew_coord.data(), 1, n_sk, f) == n_sk);
  const bool at_end = ok && fgetc(f) == EOF;
  fclose(f);
  if (!ok || !at_end) {
    delete gm;
    return io_fail(ok ? "gorilla_mesh_load: trailing bytes after the payload" : "gorilla_mesh_load: file is truncated");
  }
  uint64_t c = 14695981039346656037ull;
  c = fnv1a(c, m.tetra_physics.data(), n_tp);
  c = fnv1a(c, m.tetra_grid.data(), n_tg);
  if (n_v) c = fnv1a(c, m.verts_rphiz.data(), n_v);
  if (has_sth) c = fnv1a(c, m.verts_sthetaphi.data(), n_v);
  if (has_skew) c = fnv1a(c, m.tetra_skew_coord.data(), n_sk);
  if (c != h.checksum) {
    delete gm;
    return io_fail("gorilla_mesh_load: checksum mismatch (file is corrupted)");
  }
  *out = gm;
  return GORILLA_OK;
}
