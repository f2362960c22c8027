// gb_orbit_k2.cu -- orbit_kernel<2, *>: polynomial order 2 of the persistent push kernel (see gb_internal.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<2, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<2, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
