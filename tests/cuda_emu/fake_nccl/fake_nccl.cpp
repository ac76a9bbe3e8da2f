// fake_nccl.cpp -- TEST INFRASTRUCTURE ONLY: the nine NCCL entry points liblpmgpu resolves with dlsym
// (csrc/runtime.cuh, load_nccl), implemented over one POSIX shared-memory segment and a process-shared barrier,
// for the ranks of an emulated run (tests/cuda_emu).  Built as tests/cuda_emu/fake_nccl/libnccl.so.2 and found
// through LD_LIBRARY_PATH by the emulated library only.  Collectives are synchronous (the emulator's streams are):
// every rank copies its contribution into its slot, a barrier, every rank reads what it needs, a barrier.
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace {
constexpr size_t kSlot = 4u << 20;          // bytes per rank and round
constexpr int kMaxRanks = 8;
constexpr uint32_t kMagic = 0x4c504d4eu;
struct Control {
    volatile uint32_t ready;
    pthread_barrier_t barrier;
};
struct Comm {
    char name[64];
    int nranks, rank;
    unsigned char* base;
    size_t bytes;
    Control* ctl() const { return reinterpret_cast<Control*>(base); }
    unsigned char* slot(int r) const { return base + 4096 + (size_t)r * kSlot; }
    void sync() const { pthread_barrier_wait(&ctl()->barrier); }
};
size_t esize(int dtype)
{
    switch (dtype) {
        case 0: case 1: return 1;       // ncclInt8, ncclUint8
        case 2: case 3: case 7: return 4;       // ncclInt32, ncclUint32, ncclFloat32
        case 4: case 5: case 8: return 8;       // ncclInt64, ncclUint64, ncclFloat64
        default: return 0;
    }
}
}  // namespace

struct ncclUniqueId { char internal[128]; };
#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT int ncclGetUniqueId(ncclUniqueId* id)
{
    std::memset(id, 0, sizeof(*id));
    std::snprintf(id->internal, sizeof(id->internal), "/lpmnccl_%d_%ld", (int)getpid(), (long)time(nullptr));
    return 0;
}
EXPORT int ncclCommInitRank(void** comm, int nranks, ncclUniqueId id, int rank)
{
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return 4;
    Comm* c = new Comm();
    std::snprintf(c->name, sizeof(c->name), "%s", id.internal);
    c->nranks = nranks; c->rank = rank;
    c->bytes = 4096 + (size_t)nranks * kSlot;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(c->name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)c->bytes) != 0) return 2;
    } else {
        for (int tries = 0; tries < 60000 && fd < 0; ++tries) {
            fd = shm_open(c->name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < c->bytes)) { close(fd); fd = -1; }
            if (fd < 0) usleep(1000);
        }
        if (fd < 0) return 2;
    }
    c->base = (unsigned char*)mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->base == (unsigned char*)MAP_FAILED) return 2;
    if (rank == 0) {
        pthread_barrierattr_t a;
        pthread_barrierattr_init(&a);
        pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
        pthread_barrier_init(&c->ctl()->barrier, &a, (unsigned)nranks);
        __sync_synchronize();
        c->ctl()->ready = kMagic;
    } else {
        while (c->ctl()->ready != kMagic) usleep(200);
    }
    c->sync();
    *comm = c;
    return 0;
}
EXPORT int ncclCommDestroy(void* comm)
{
    Comm* c = (Comm*)comm;
    if (!c) return 0;
    c->sync();
    munmap(c->base, c->bytes);
    if (c->rank == 0) shm_unlink(c->name);
    delete c;
    return 0;
}
EXPORT int ncclGroupStart() { return 0; }
EXPORT int ncclGroupEnd() { return 0; }
EXPORT const char* ncclGetErrorString(int e) { return e == 0 ? "no error" : "fake NCCL error (tests/cuda_emu)"; }

EXPORT int ncclBroadcast(const void* send, void* recv, size_t count, int dtype, int root, void* comm, void*)
{
    Comm* c = (Comm*)comm;
    const size_t es = esize(dtype);
    if (!es) return 4;
    const size_t total = count * es;
    for (size_t off = 0; off < total || (total == 0 && off == 0); off += kSlot) {
        const size_t n = total - off < kSlot ? total - off : kSlot;
        if (c->rank == root) std::memcpy(c->slot(0), (const unsigned char*)send + off, n);
        c->sync();
        if (c->rank != root || recv != send) std::memcpy((unsigned char*)recv + off, c->slot(0), n);
        c->sync();
        if (total == 0) break;
    }
    return 0;
}
EXPORT int ncclAllGather(const void* send, void* recv, size_t sendcount, int dtype, void* comm, void*)
{
    Comm* c = (Comm*)comm;
    const size_t es = esize(dtype);
    if (!es) return 4;
    const size_t total = sendcount * es;
    for (size_t off = 0; off < total; off += kSlot) {
        const size_t n = total - off < kSlot ? total - off : kSlot;
        std::memcpy(c->slot(c->rank), (const unsigned char*)send + off, n);
        c->sync();
        for (int r = 0; r < c->nranks; ++r) std::memcpy((unsigned char*)recv + (size_t)r * total + off, c->slot(r), n);
        c->sync();
    }
    return 0;
}
EXPORT int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void*)
{
    Comm* c = (Comm*)comm;
    if (op != 0 || (dtype != 2 && dtype != 8 && dtype != 4)) return 4;        // sums of int32 / int64 / float64 only
    const size_t es = esize(dtype);
    const size_t per = kSlot / es;
    for (size_t off = 0; off < count; off += per) {
        const size_t n = count - off < per ? count - off : per;
        std::memcpy(c->slot(c->rank), (const unsigned char*)send + off * es, n * es);
        c->sync();
        if (dtype == 8) {
            double* out = (double*)recv + off;
            for (size_t i = 0; i < n; ++i) {
                double s = 0.0;
                for (int r = 0; r < c->nranks; ++r) s += ((const double*)c->slot(r))[i];     // rank order: every rank gets the same bits
                out[i] = s;
            }
        } else if (dtype == 4) {
            int64_t* out = (int64_t*)recv + off;
            for (size_t i = 0; i < n; ++i) {
                int64_t s = 0;
                for (int r = 0; r < c->nranks; ++r) s += ((const int64_t*)c->slot(r))[i];
                out[i] = s;
            }
        } else {
            int32_t* out = (int32_t*)recv + off;
            for (size_t i = 0; i < n; ++i) {
                int32_t s = 0;
                for (int r = 0; r < c->nranks; ++r) s += ((const int32_t*)c->slot(r))[i];
                out[i] = s;
            }
        }
        c->sync();
    }
    return 0;
}
