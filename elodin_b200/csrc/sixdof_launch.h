// Host-side launch helpers shared by the kernel translation units.
#pragma once
#include <cstdlib>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

namespace b200 {

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): remember the size each pair has been
// raised to (a process may hold handles on several GPUs; several kernels share a function-pointer type;
// one kernel may be launched with several tile sizes).
template <typename K>
static cudaError_t ensure_dynamic_smem(K kernel, size_t bytes)
{
    struct Entry { const void *kernel; int device; size_t bytes; };
    static std::mutex mu;
    static std::vector<Entry> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const void *key = reinterpret_cast<const void *>(kernel);
    std::lock_guard<std::mutex> lock(mu);
    Entry *hit = nullptr;
    for (auto &d : done) if (d.kernel == key && d.device == dev) hit = &d;
    if (hit && hit->bytes >= bytes) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (hit) hit->bytes = bytes; else done.push_back(Entry{key, dev, bytes});
    return cudaSuccess;
}


inline int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

} // namespace b200
