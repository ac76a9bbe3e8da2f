// emu_runtime.h -- TEST INFRASTRUCTURE ONLY: the slice of the CUDA runtime API that lpm_v2_b200/csrc uses,
// for the SIMT emulator.  "Device" memory is host memory, streams are synchronous (every operation has
// completed when the call returns), events are wall-clock stamps.  LPM_EMU_DEVICES (default 1) emulated
// devices share that memory, so peer stores simply work.
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <map>
#include <mutex>
#include <string>

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801,
                   cudaErrorPeerAccessAlreadyEnabled = 704 };
inline const char* cudaGetErrorString(cudaError_t e)
{
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorMemoryAllocation: return "out of memory (emulated)";
        case cudaErrorNotSupported: return "not supported by the emulator";
        default: return "emulated CUDA error";
    }
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct EmuStream { int id; };
using cudaStream_t = EmuStream*;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
using cudaEvent_t = EmuEvent*;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostRegisterPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; };
struct cudaIpcMemHandle_t { char reserved[64]; };

inline int& emu_current_device() { static thread_local int d = 0; return d; }
inline int emu_device_count()
{
    const char* e = std::getenv("LPM_EMU_DEVICES");
    const int n = e ? std::atoi(e) : 1;
    return n < 1 ? 1 : (n > 8 ? 8 : n);
}
inline cudaError_t cudaGetDeviceCount(int* n) { *n = emu_device_count(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { emu_current_device() = d; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = emu_current_device(); return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    std::snprintf(p->name, sizeof(p->name), "emulated sm_100a (tests/cuda_emu)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 2;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

// Device allocations are POSIX shared-memory segments, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map
// one process's "device" buffer into another (rank mode: the shared slabs the ranks store their slices into).
struct EmuAlloc { std::string name; size_t bytes; };
inline std::map<void*, EmuAlloc>& emu_allocs() { static std::map<void*, EmuAlloc> m; return m; }
inline std::map<void*, size_t>& emu_ipc_maps() { static std::map<void*, size_t> m; return m; }
inline std::mutex& emu_alloc_mutex() { static std::mutex m; return m; }
// Segments are unlinked by cudaFree, at exit for whatever the process still holds, and -- for processes that
// were killed -- by the next emulated process that starts (names carry the owner's pid).
struct EmuShmJanitor {
    EmuShmJanitor()
    {
        if (DIR* d = opendir("/dev/shm")) {
            while (dirent* e = readdir(d)) {
                int pid = 0, k = 0;
                if (std::sscanf(e->d_name, "lpmemu_%d_%d", &pid, &k) == 2 && pid > 0 && kill(pid, 0) != 0 && errno == ESRCH)
                    shm_unlink((std::string("/") + e->d_name).c_str());
            }
            closedir(d);
        }
    }
    ~EmuShmJanitor()
    {
        for (auto& a : emu_allocs())
            if (!a.second.name.empty()) shm_unlink(a.second.name.c_str());
    }
};
inline EmuShmJanitor& emu_janitor() { static EmuShmJanitor j; return j; }
inline cudaError_t emu_malloc(void** p, size_t bytes)
{
    static int counter = 0;
    emu_allocs();           // constructed before the janitor, so destroyed after it
    emu_janitor();
    std::lock_guard<std::mutex> lock(emu_alloc_mutex());
    if (bytes == 0) bytes = 1;
    char name[64];
    std::snprintf(name, sizeof(name), "/lpmemu_%d_%d", (int)getpid(), counter++);
    // posix_fallocate reserves the pages now: a full /dev/shm must not turn into a SIGBUS at first touch.  If the
    // segment cannot be had, fall back to private memory (such a buffer has no IPC handle).
    void* q = MAP_FAILED;
    const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd >= 0) {
        if (posix_fallocate(fd, 0, (off_t)bytes) == 0) q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (q == MAP_FAILED) shm_unlink(name);
    }
    if (q == MAP_FAILED) {
        name[0] = 0;
        q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q == MAP_FAILED) { *p = nullptr; return cudaErrorMemoryAllocation; }
    }
    std::memset(q, 0xdb, bytes);        // poison: nothing may rely on fresh device memory being zero
    emu_allocs()[q] = EmuAlloc{name, bytes};
    *p = q;
    return cudaSuccess;
}
template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return emu_malloc((void**)p, bytes); }
inline cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lock(emu_alloc_mutex());
    auto it = emu_allocs().find(p);
    if (it == emu_allocs().end()) return cudaErrorInvalidValue;
    munmap(p, it->second.bytes);
    if (!it->second.name.empty()) shm_unlink(it->second.name.c_str());
    emu_allocs().erase(it);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* s, size_t n) { std::memcpy(&sym, s, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaGetSymbolAddress(void** p, T& sym) { *p = (void*)&sym; return cudaSuccess; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new EmuStream{0}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    if (!(*ms > 1e-6f)) *ms = 1e-6f;
    return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
    std::lock_guard<std::mutex> lock(emu_alloc_mutex());
    auto it = emu_allocs().find(p);
    if (it == emu_allocs().end() || it->second.name.empty()) return cudaErrorInvalidValue;
    std::memset(h, 0, sizeof(*h));
    std::snprintf(h->reserved, 48, "%s", it->second.name.c_str());
    const uint64_t bytes = it->second.bytes;
    std::memcpy(h->reserved + 48, &bytes, 8);
    return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
    uint64_t bytes = 0;
    std::memcpy(&bytes, h.reserved + 48, 8);
    const int fd = shm_open(h.reserved, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    void* q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (q == MAP_FAILED) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(emu_alloc_mutex());
    emu_ipc_maps()[q] = bytes;
    *p = q;
    return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void* p)
{
    std::lock_guard<std::mutex> lock(emu_alloc_mutex());
    auto it = emu_ipc_maps().find(p);
    if (it == emu_ipc_maps().end()) return cudaErrorInvalidValue;
    munmap(p, it->second);
    emu_ipc_maps().erase(it);
    return cudaSuccess;
}
