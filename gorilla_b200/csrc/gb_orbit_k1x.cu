// gb_orbit_k1x.cu -- EXT variant of polynomial order 1: orbit_kernel<1, *, true> with Hamiltonian time
// tracing (i_time_tracing_option = 2) and the optional quantities of pusher_tetra_poly (see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<1, 0, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 1, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 2, true>(gorilla_b200_handle *, const Batch &, cudaStream_t);
