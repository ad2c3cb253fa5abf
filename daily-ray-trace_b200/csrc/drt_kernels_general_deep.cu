/* csrc/drt_kernels_general_deep.cu -- instantiates drt::render_kernel<float, NS, 0, PAIRED, true> (drt_render.cuh) for NS = 2, 3, 5, 8:
 * kernel mode 0, records overflowing to global memory (deep renders). */
#define DRT_PHILOX_ROLLED 1   /* this kernel is bound by instruction fetch (hot code > 32 KB): smaller beats straight-line */
#include "drt_render.cuh"

DRT_DEFINE_LAUNCHER(drt_launch_render_f32_general_deep, float, 0, true)
