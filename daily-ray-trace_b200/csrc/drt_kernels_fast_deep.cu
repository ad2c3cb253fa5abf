/* csrc/drt_kernels_fast_deep.cu -- instantiates drt::render_kernel<float, NS, 1, PAIRED, true> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 1, records overflowing to global memory (deep renders). */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f32_fast_deep, float, 1, true)
