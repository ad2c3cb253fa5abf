/* csrc/drt_kernels_classed.cu -- instantiates drt::render_kernel<float, NS, 2, PAIRED, false> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 2. */
#define DRT_PHILOX_ROLLED 1   /* this kernel is bound by instruction fetch (hot code > 32 KB): smaller beats straight-line */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f32_classed, float, 2, false)
