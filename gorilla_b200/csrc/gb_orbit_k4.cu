// gb_orbit_k4.cu -- orbit_kernel<4, *>: polynomial order 4 of the persistent push kernel (see gb_internal.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<4, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<4, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
