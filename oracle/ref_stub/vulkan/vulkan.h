/* Stand-in for <vulkan/vulkan.h>, which this image does not have: just enough for the reference's render/GpuBuffers.h to compile
 * UNMODIFIED inside oracle/_ref (oracle/Makefile, target _ref/libanerend_ref_host.so).  The two structs are the plain-C layouts
 * the Vulkan specification defines; nothing on the DDGI path uses them.  Test infrastructure only. */
#ifndef ANEREND_REF_STUB_VULKAN_H
#define ANEREND_REF_STUB_VULKAN_H
#include <stdint.h>
typedef struct VkDrawIndexedIndirectCommand { uint32_t indexCount, instanceCount, firstIndex; int32_t vertexOffset; uint32_t firstInstance; } VkDrawIndexedIndirectCommand;
typedef struct VkDrawIndirectCommand { uint32_t vertexCount, instanceCount, firstVertex, firstInstance; } VkDrawIndirectCommand;
#endif
