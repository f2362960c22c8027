// gb_orbit_rk.cu -- orbit_kernel<0, *>: the RK4 pusher (ipusher = 1) variant of the persistent push kernel
#include "gb_internal.cuh"
template int launch_orbit_t<0, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<0, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<0, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
