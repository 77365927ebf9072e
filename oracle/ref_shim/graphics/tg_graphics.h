/* ORACLE -- TEST INFRASTRUCTURE. Stand-in for /root/reference/tg/src/graphics/tg_graphics.h, which only selects the Vulkan
 * back end (#error otherwise, tg_graphics.h:7-33): the portable files need the core types alone. */
#ifndef TG_GRAPHICS_H
#define TG_GRAPHICS_H
#include "graphics/tg_graphics_core.h"
#include "graphics/vulkan/tgvk_raytracer.h"
#endif
