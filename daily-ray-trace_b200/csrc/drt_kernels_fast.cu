/* csrc/drt_kernels_fast.cu -- instantiates drt::render_kernel<float, NS, 1, PAIRED, false> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 1. */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f32_fast, float, 1, false)
