// cudaengine.cpp — device registry of the drop-in (see dropin/include/common/cudaengine.cuh).
#include "common/cudaengine.cuh"

#include "b200sv.h"

#include <cstdlib>
#include <stdexcept>
#include <string>

namespace Qrack {

CUDADeviceContext::CUDADeviceContext(int64_t dev, size_t totalBytes, int smCount)
    : device_id(dev)
    , context_id(dev)
    , globalSize(totalBytes)
{
    // Page granule: the largest power of two not above 1/8 of device memory (16 GiB on a 180 GB B200), so that a
    // 33-qubit fp32 register splits into exactly 8 pages of 2^30 amplitudes = one page per GPU of the box, with room
    // for the out-of-place scratch that Compose/Decompose and QPager's re-paging need.  QRACK_MAX_ALLOC_MB overrides
    // it like in the reference (include/common/cudaengine.cuh:96-98 there).
    size_t limit = totalBytes / 8U;
    if (const char* mb = getenv("QRACK_MAX_ALLOC_MB")) {
        const size_t v = (size_t)std::stoull(std::string(mb)) << 20U;
        if (v) {
            limit = v;
        }
    }
    size_t p = 1U;
    while ((p << 1U) <= limit) {
        p <<= 1U;
    }
    maxAlloc = p;
    // one work item per resident thread: SMs x 2048, rounded up to a power of two
    size_t c = 1U;
    while (c < (size_t)smCount * 2048U) {
        c <<= 1U;
    }
    preferredConcurrency = c;
}

CUDAEngine::CUDAEngine()
{
    int n = 0;
    if (b200sv_device_count(&n) != B200SV_OK) {
        n = 0;
    }
    for (int d = 0; d < n; ++d) {
        uint64_t total = 0, freeb = 0;
        int sms = 148;
        if (b200sv_device_info(d, &total, &freeb, &sms) != B200SV_OK) {
            continue;
        }
        all_device_contexts.push_back(std::make_shared<CUDADeviceContext>(d, (size_t)total, sms));
    }
    activeAllocSizes.assign(all_device_contexts.size(), 0U);
    if (!all_device_contexts.empty()) {
        size_t def = 0U;
        if (const char* e = getenv("QRACK_OCL_DEFAULT_DEVICE")) {
            const long v = std::atol(e);
            if (v >= 0 && (size_t)v < all_device_contexts.size()) {
                def = (size_t)v;
            }
        }
        default_device = all_device_contexts[def];
    }
}

size_t CUDAEngine::Index(const int64_t& dev)
{
    if (dev < 0) {
        return GetDefaultDeviceID();
    }
    if ((size_t)dev >= all_device_contexts.size()) {
        throw std::invalid_argument("Invalid CUDA device selection");
    }
    return (size_t)dev;
}

DeviceContextPtr CUDAEngine::GetDeviceContextPtr(const int64_t& dev)
{
    if (all_device_contexts.empty()) {
        throw std::runtime_error("No CUDA device available");
    }
    return all_device_contexts[Index(dev)];
}

void CUDAEngine::SetDeviceContextPtrVector(std::vector<DeviceContextPtr> vec, DeviceContextPtr dcp)
{
    all_device_contexts = vec;
    activeAllocSizes.assign(vec.size(), 0U);
    if (dcp) {
        default_device = dcp;
    }
}

size_t CUDAEngine::GetActiveAllocSize(const int64_t& dev)
{
    std::lock_guard<std::mutex> lock(allocMutex);
    return activeAllocSizes[Index(dev)];
}
size_t CUDAEngine::AddToActiveAllocSize(const int64_t& dev, size_t size)
{
    std::lock_guard<std::mutex> lock(allocMutex);
    return activeAllocSizes[Index(dev)] += size;
}
size_t CUDAEngine::SubtractFromActiveAllocSize(const int64_t& dev, size_t size)
{
    std::lock_guard<std::mutex> lock(allocMutex);
    size_t& a = activeAllocSizes[Index(dev)];
    a = (size < a) ? (a - size) : 0U;
    return a;
}
void CUDAEngine::ResetActiveAllocSize(const int64_t& dev)
{
    std::lock_guard<std::mutex> lock(allocMutex);
    activeAllocSizes[Index(dev)] = 0U;
}

} // namespace Qrack
