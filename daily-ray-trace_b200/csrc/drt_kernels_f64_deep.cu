/* csrc/drt_kernels_f64_deep.cu -- instantiates drt::render_kernel<double, NS, 0, PAIRED, true> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 0, records overflowing to global memory (deep renders). */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f64_deep, double, 0, true)
