// gb_orbit_k4p.cu -- EXT = 4 variant of polynomial order 4: precomputed coefficients, i_precomp = 1
// (tetra_physics_poly4 records; see gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<4, 0, 4>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 1, 4>(gorilla_b200_handle *, const Batch &, cudaStream_t);
