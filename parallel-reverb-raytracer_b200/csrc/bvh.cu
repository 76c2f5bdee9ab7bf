#include "rvb_internal.cuh"
namespace rvb {
int scene_build_bvh(Scene *) { return fail(RVB_ERR_UNSUPPORTED, "BVH build not implemented yet"); }
}
