// gb_orbit_k2a.cu -- EXT = 3 variant of polynomial order 2: adaptive energy-controlled sub-stepping
// (boole_adaptive_time_steps; see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<2, 0, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 1, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 2, 3>(gorilla_b200_handle *, const Batch &, cudaStream_t);
