// common/cudaengine.cuh — drop-in replacement of the reference's device singleton header
// (/root/reference/include/common/cudaengine.cuh:35-243) for the B200-native engine.
//
// Only what the reference's callers actually use is provided: QPager (src/qpager.cpp:98-272,
// include/qpager.hpp:141,504), QHybrid (src/qhybrid.cpp:38,43), QUnit (src/qunit.cpp:84), QUnitMulti
// (src/qunitmulti.cpp:63-204), QStabilizerHybrid, QTensorNetwork, the factory (include/qfactory.hpp:263) and the
// test/benchmark mains.  No CUDA headers are needed here: everything goes through the C ABI (include/b200sv.h).
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

namespace Qrack {

class CUDADeviceContext {
public:
    const int64_t device_id;
    const int64_t context_id; // one context per device

private:
    size_t globalSize;
    size_t maxAlloc;
    size_t preferredConcurrency;

public:
    CUDADeviceContext(int64_t dev, size_t totalBytes, int smCount);

    /// Largest single allocation QPager should place on this device.  QPager derives its page size from it
    /// (qubits per page = log2(maxAlloc / sizeof(complex)) - 1, reference src/qpager.cpp:129).
    size_t GetMaxAlloc() const { return maxAlloc; }
    size_t GetGlobalSize() const { return globalSize; }
    /// Number of work items that saturates the device; QHybrid / QUnitMulti derive the CPU<->GPU switch-over width
    /// from it (reference src/qhybrid.cpp:37-40).
    size_t GetPreferredConcurrency() const { return preferredConcurrency; }
    size_t GetPreferredSizeMultiple() const { return 32U; }
    size_t GetGlobalAllocLimit() const { return globalSize; }
};

typedef std::shared_ptr<CUDADeviceContext> DeviceContextPtr;

/** Process-wide registry of the visible B200s (replaces Qrack::CUDAEngine). */
class CUDAEngine {
public:
    static CUDAEngine& Instance()
    {
        static CUDAEngine instance;
        return instance;
    }

    int GetDeviceCount() { return (int)all_device_contexts.size(); }
    size_t GetDefaultDeviceID() { return default_device ? (size_t)default_device->device_id : 0U; }
    DeviceContextPtr GetDeviceContextPtr(const int64_t& dev = -1);
    std::vector<DeviceContextPtr> GetDeviceContextPtrVector() { return all_device_contexts; }
    void SetDeviceContextPtrVector(std::vector<DeviceContextPtr> vec, DeviceContextPtr dcp = nullptr);
    void SetDefaultDeviceContext(DeviceContextPtr dcp) { default_device = dcp; }

    /// Allocation accounting used by QUnitMulti's load balancer (reference src/qunitmulti.cpp:156-204)
    size_t GetActiveAllocSize(const int64_t& dev);
    size_t AddToActiveAllocSize(const int64_t& dev, size_t size);
    size_t SubtractFromActiveAllocSize(const int64_t& dev, size_t size);
    void ResetActiveAllocSize(const int64_t& dev);

    CUDAEngine(CUDAEngine const&) = delete;
    void operator=(CUDAEngine const&) = delete;

private:
    CUDAEngine();
    size_t Index(const int64_t& dev);

    std::vector<size_t> activeAllocSizes;
    std::mutex allocMutex;
    std::vector<DeviceContextPtr> all_device_contexts;
    DeviceContextPtr default_device;
};

} // namespace Qrack
