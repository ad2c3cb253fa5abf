/* csrc/drt_kernels_fast.cu -- instantiates drt::render_kernel<float, NS, 1, PAIRED> (drt_render.cuh) for NS = 2, 3, 5, 8. */
#include "drt_render.cuh"

cudaError_t drt_launch_render_f32_fast(const RenderLaunch &L, bool paired, int nslots, int grid, int warps, size_t smem, cudaStream_t stream)
{
    return paired ? drt_launch_render_ns<float, 1, true>(L, nslots, grid, warps, smem, stream)
                  : drt_launch_render_ns<float, 1, false>(L, nslots, grid, warps, smem, stream);
}
