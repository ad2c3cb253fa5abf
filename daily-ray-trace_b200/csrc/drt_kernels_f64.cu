/* csrc/drt_kernels_f64.cu -- instantiates drt::render_kernel<double, NS, 0, PAIRED, false> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 0. */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f64, double, 0, false)
