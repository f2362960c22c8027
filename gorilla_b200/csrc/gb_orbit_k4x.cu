// gb_orbit_k4x.cu -- EXT variant of polynomial order 4: orbit_kernel<4, *, true> / orbit_kernel_g with Hamiltonian time
// tracing (i_time_tracing_option = 2) and the optional quantities of pusher_tetra_poly (see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<4, 0, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 1, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 2, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
