/* tgb_shade.cu -- K3 placeholder (filled in next milestone). */
#include "tgb_device.cuh"

extern "C" b32 tgbd_render_shading(struct tgb_device* d, const tg_camera_rays* p_cam, u32 gi_enabled, u32 frame_seed, u32 debug_visualization)
{
    tgb_set_error("tgbd_render_shading: not built yet");
    return TG_FALSE;
}
