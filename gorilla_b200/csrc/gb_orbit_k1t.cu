// gb_orbit_k1t.cu -- EXT = 1 variant of polynomial order 1: Hamiltonian time tracing (i_time_tracing_option = 2)
// (see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<1, 0, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 1, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 2, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
