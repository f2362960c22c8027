// gb_orbit_k4ax.cu -- EXT = 5 variant of polynomial order 4: adaptive energy-controlled sub-stepping combined with the list
// consumers (Hamiltonian time tracing, optional quantities, orbit events); see gb_internal.cuh, gb_poly.cuh
#include "gb_internal.cuh"
template int launch_orbit_t<4, 0, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 1, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 2, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
