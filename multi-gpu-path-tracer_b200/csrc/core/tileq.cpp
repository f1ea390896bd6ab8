// tileq.cpp — node-wide dynamic tile counter in POSIX shared memory (include/ptcore.h: pt_tileq_*).
//
// One rank per GPU (torchrun) pulls tile indices from the same 64-bit counter, so a rank whose
// tiles are cheap (the ~19 % of cornell_duck pixels that leave through the open front) simply
// claims more.  Replaces the fixed per-frame rectangles of the reference
// (src/RenderManager.h:42-59,105-110; src/Scheduling/TaskGenerator.h:58-80).
#include "../../../include/ptcore.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <new>
#include <string>

struct pt_tileq {
    std::string name;
    std::atomic<int64_t> *counter = nullptr;
    int fd = -1;
};

static_assert(sizeof(std::atomic<int64_t>) == 8 && std::atomic<int64_t>::is_always_lock_free, "need an 8-byte lock-free atomic");

extern "C" {

int pt_tileq_open(const char *name, int create, pt_tileq_t **out) {
    if (!name || !out || name[0] != '/') return PT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return PT_ERR_SYSTEM;
    if (create && ftruncate(fd, 64) != 0) { close(fd); return PT_ERR_SYSTEM; }
    void *p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) { close(fd); return PT_ERR_SYSTEM; }
    pt_tileq *q = new (std::nothrow) pt_tileq();
    if (!q) { munmap(p, 64); close(fd); return PT_ERR_SYSTEM; }
    q->name = name;
    q->fd = fd;
    q->counter = reinterpret_cast<std::atomic<int64_t> *>(p);
    if (create) q->counter->store(0, std::memory_order_seq_cst);
    *out = q;
    return PT_OK;
}

int64_t pt_tileq_claim(pt_tileq_t *q, int64_t count, int64_t limit) {
    if (!q || count <= 0) return -1;
    int64_t first = q->counter->fetch_add(count, std::memory_order_acq_rel);
    return first < limit ? first : -1;
}

int pt_tileq_reset(pt_tileq_t *q) {
    if (!q) return PT_ERR_INVALID_ARGUMENT;
    q->counter->store(0, std::memory_order_seq_cst);
    return PT_OK;
}

int pt_tileq_close(pt_tileq_t *q, int unlink_name) {
    if (!q) return PT_OK;
    munmap((void *)q->counter, 64);
    close(q->fd);
    if (unlink_name) shm_unlink(q->name.c_str());
    delete q;
    return PT_OK;
}

}  // extern "C"
