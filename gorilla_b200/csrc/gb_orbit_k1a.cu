// gb_orbit_k1a.cu -- EXT = 3 variant of polynomial order 1: adaptive energy-controlled sub-stepping
// (boole_adaptive_time_steps; see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<1, 0, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 1, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 2, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
