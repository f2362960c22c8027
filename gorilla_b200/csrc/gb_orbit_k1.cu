// gb_orbit_k1.cu -- orbit_kernel<1, *>: polynomial order 1 of the persistent push kernel (see gb_internal.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<1, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<1, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
