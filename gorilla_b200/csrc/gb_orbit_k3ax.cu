// gb_orbit_k3ax.cu -- EXT = 5 variant of polynomial order 3: adaptive energy-controlled sub-stepping combined with the list
// consumers (Hamiltonian time tracing, optional quantities, orbit events); see gb_internal.cuh, gb_poly.cuh
#include "gb_internal.cuh"
template int launch_orbit_t<3, 0, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<3, 1, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<3, 2, 5>(gorilla_b200_handle *, const Batch &, cudaStream_t);
