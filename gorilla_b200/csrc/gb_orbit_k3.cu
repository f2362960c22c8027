// gb_orbit_k3.cu -- orbit_kernel<3, *>: polynomial order 3 of the persistent push kernel (see gb_internal.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<3, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<3, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<3, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
