// gb_orbit_rkx.cu -- orbit_kernel<0, *, 2>: the RK4 pusher with handover_processing_kind = 2 (position exchange via
// Cartesian skew coordinates; see gb_internal.cuh, gb_poly.cuh)
#include "gb_internal.cuh"
template int launch_orbit_t<0, 0, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<0, 1, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
template int launch_orbit_t<0, 2, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
