// gb_orbit_k4t.cu -- EXT = 1 variant of polynomial order 4: Hamiltonian time tracing (i_time_tracing_option = 2)
// (see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<4, 0, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 1, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 2, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
