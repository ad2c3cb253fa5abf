/* csrc/drt_kernels_f64.cu -- instantiates drt::render_kernel<double, NS, 0, PAIRED> (drt_render.cuh) for NS = 2, 3, 5, 8. */
#include "drt_render.cuh"

cudaError_t drt_launch_render_f64(const RenderLaunch &L, bool paired, int nslots, int grid, int warps, size_t smem, cudaStream_t stream)
{
    return paired ? drt_launch_render_ns<double, 0, true>(L, nslots, grid, warps, smem, stream)
                  : drt_launch_render_ns<double, 0, false>(L, nslots, grid, warps, smem, stream);
}
