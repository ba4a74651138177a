// Opaque-handle stand-in for <vulkan/vulkan.h>: src/engine/include/sc_assets.h:3 only needs the
// handle types as struct members. Test infrastructure only (oracle build of the reference).
#pragma once
#include <cstdint>
#define VK_NULL_HANDLE nullptr
#define SC_VK_HANDLE(name) typedef struct name##_T* name;
SC_VK_HANDLE(VkInstance)
SC_VK_HANDLE(VkDevice)
SC_VK_HANDLE(VkPhysicalDevice)
SC_VK_HANDLE(VkQueue)
SC_VK_HANDLE(VkCommandPool)
SC_VK_HANDLE(VkCommandBuffer)
SC_VK_HANDLE(VkDescriptorPool)
SC_VK_HANDLE(VkDescriptorSetLayout)
SC_VK_HANDLE(VkDescriptorSet)
SC_VK_HANDLE(VkImage)
SC_VK_HANDLE(VkDeviceMemory)
SC_VK_HANDLE(VkImageView)
SC_VK_HANDLE(VkSampler)
SC_VK_HANDLE(VkBuffer)
SC_VK_HANDLE(VkPipeline)
SC_VK_HANDLE(VkPipelineLayout)
typedef uint32_t VkFormat;
typedef uint32_t VkFlags;
typedef uint64_t VkDeviceSize;
