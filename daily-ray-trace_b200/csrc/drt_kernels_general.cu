/* csrc/drt_kernels_general.cu -- instantiates drt::render_kernel<float, NS, 0, PAIRED> (drt_render.cuh) for NS = 2, 3, 5, 8. */
#define DRT_PHILOX_ROLLED 1   /* this kernel is bound by instruction fetch (hot code > 32 KB): smaller beats straight-line */
#include "drt_render.cuh"

cudaError_t drt_launch_render_f32_general(const RenderLaunch &L, bool paired, int nslots, int grid, int warps, size_t smem, cudaStream_t stream)
{
    return paired ? drt_launch_render_ns<float, 0, true>(L, nslots, grid, warps, smem, stream)
                  : drt_launch_render_ns<float, 0, false>(L, nslots, grid, warps, smem, stream);
}
