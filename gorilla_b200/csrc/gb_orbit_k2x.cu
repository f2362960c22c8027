// gb_orbit_k2x.cu -- EXT = 2 variant of polynomial order 2: time tracing option read at run time + the optional quantities of pusher_tetra_poly
// (see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<2, 0, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 1, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 2, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
