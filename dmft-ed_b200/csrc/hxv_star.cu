// hxv_star.cu -- star-product layout and tiled H*v kernels (placeholder until the tiled kernels land).
#include "edgpu_internal.h"

int build_star_layout(edgpu_ctx *ctx, SpinBasis *, const std::vector<HopPair> &, const std::vector<double> &)
{
    return edgpu_fail(ctx, "star-product layout not available in this build");
}

int hxv_star(edgpu_sector *s, const double *, double *)
{
    return edgpu_fail(s->ctx, "star-product kernels not available in this build");
}
