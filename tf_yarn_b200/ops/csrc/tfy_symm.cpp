// Symmetric HBM arenas for one-process-per-GPU jobs on an NVSwitch box.
//
// Every rank allocates one physical arena with the CUDA VMM API
// (cuMemCreate), exports it as a POSIX file descriptor, ships the fd to every
// peer over a unix-domain socket (SCM_RIGHTS), and maps every peer's arena
// into its own address space.  When the fabric supports it, rank 0 also
// creates a multicast object; every rank binds its arena to it so that
// `multimem.ld_reduce` / `multimem.st` on the multicast VA are reduced /
// replicated inside the NVSwitch (NVLS).
//
// This replaces the host-side transport the reference reaches through its
// dependencies (Horovod gloo context bootstrap, reference:
// tf_yarn/tensorflow/tasks/gloo_allred_task.py:42-54,119-122; c10d TCPStore +
// NCCL communicator bootstrap, reference: tf_yarn/pytorch/tasks/worker.py:101).
// Rendezvous between the steps (socket paths, barriers) goes through the
// launcher's KV store on the Python side; this file owns only the CUDA and fd
// plumbing.  The driver API is resolved at run time through
// cudaGetDriverEntryPoint so the library loads on machines without libcuda.
#include <cuda.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#define TFY_MAX_RANKS 16

namespace {

thread_local char g_err[1024] = "";

void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct DriverApi {
    bool loaded = false;
#define TFY_DRV(name) decltype(&name) p_##name = nullptr;
    TFY_DRV(cuMemCreate)
    TFY_DRV(cuMemRelease)
    TFY_DRV(cuMemAddressReserve)
    TFY_DRV(cuMemAddressFree)
    TFY_DRV(cuMemMap)
    TFY_DRV(cuMemUnmap)
    TFY_DRV(cuMemSetAccess)
    TFY_DRV(cuMemGetAllocationGranularity)
    TFY_DRV(cuMemExportToShareableHandle)
    TFY_DRV(cuMemImportFromShareableHandle)
    TFY_DRV(cuMulticastCreate)
    TFY_DRV(cuMulticastAddDevice)
    TFY_DRV(cuMulticastBindMem)
    TFY_DRV(cuMulticastUnbind)
    TFY_DRV(cuMulticastGetGranularity)
    TFY_DRV(cuDeviceGetAttribute)
    TFY_DRV(cuDeviceGet)
    TFY_DRV(cuGetErrorString)
#undef TFY_DRV
};
DriverApi g_drv;

template <typename F>
bool load_sym(const char* name, F* out) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st);
    if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || fn == nullptr) {
        set_err("driver entry point %s not found (%s)", name, cudaGetErrorString(e));
        return false;
    }
    *out = reinterpret_cast<F>(fn);
    return true;
}

bool load_driver() {
    if (g_drv.loaded) return true;
#define TFY_LOAD(name) \
    if (!load_sym(#name, &g_drv.p_##name)) return false;
    TFY_LOAD(cuMemCreate)
    TFY_LOAD(cuMemRelease)
    TFY_LOAD(cuMemAddressReserve)
    TFY_LOAD(cuMemAddressFree)
    TFY_LOAD(cuMemMap)
    TFY_LOAD(cuMemUnmap)
    TFY_LOAD(cuMemSetAccess)
    TFY_LOAD(cuMemGetAllocationGranularity)
    TFY_LOAD(cuMemExportToShareableHandle)
    TFY_LOAD(cuMemImportFromShareableHandle)
    TFY_LOAD(cuMulticastCreate)
    TFY_LOAD(cuMulticastAddDevice)
    TFY_LOAD(cuMulticastBindMem)
    TFY_LOAD(cuMulticastUnbind)
    TFY_LOAD(cuMulticastGetGranularity)
    TFY_LOAD(cuDeviceGetAttribute)
    TFY_LOAD(cuDeviceGet)
    TFY_LOAD(cuGetErrorString)
#undef TFY_LOAD
    g_drv.loaded = true;
    return true;
}

bool cu_ok(CUresult r, const char* what) {
    if (r == CUDA_SUCCESS) return true;
    const char* s = nullptr;
    if (g_drv.p_cuGetErrorString) g_drv.p_cuGetErrorString(r, &s);
    set_err("%s failed: %d (%s)", what, (int)r, s ? s : "?");
    return false;
}

struct Msg {
    int32_t tag;
    int32_t src;
};
enum { TAG_MEM = 1, TAG_MC = 2 };

struct Symm {
    int device = 0, rank = 0, world = 1;
    size_t size = 0;  // rounded arena size
    size_t gran = 0;
    std::string sock_prefix;
    int listen_fd = -1;
    CUmemGenericAllocationHandle local_h = 0;
    int local_fd = -1;
    CUmemGenericAllocationHandle peer_h[TFY_MAX_RANKS] = {0};
    CUdeviceptr peer_va[TFY_MAX_RANKS] = {0};
    bool mc_supported = false;
    CUmemGenericAllocationHandle mc_h = 0;
    bool mc_bound = false;
    CUdeviceptr mc_va = 0;
    std::map<std::pair<int, int>, int> stash;  // (tag, src) -> fd
};

std::string sock_path(const Symm* s, int rank) { return s->sock_prefix + "." + std::to_string(rank); }

bool send_fd(const Symm* s, int dst, int tag, int fd) {
    int sock = socket(AF_UNIX, SOCK_STREAM, 0);
    if (sock < 0) { set_err("socket: %s", strerror(errno)); return false; }
    sockaddr_un addr;
    memset(&addr, 0, sizeof(addr));
    addr.sun_family = AF_UNIX;
    std::string p = sock_path(s, dst);
    strncpy(addr.sun_path, p.c_str(), sizeof(addr.sun_path) - 1);
    int tries = 0;
    while (connect(sock, (sockaddr*)&addr, sizeof(addr)) != 0) {
        if (++tries > 600) { set_err("connect %s: %s", p.c_str(), strerror(errno)); close(sock); return false; }
        usleep(50 * 1000);
    }
    Msg m{tag, s->rank};
    iovec iov{&m, sizeof(m)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    memset(cbuf, 0, sizeof(cbuf));
    msghdr mh;
    memset(&mh, 0, sizeof(mh));
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    mh.msg_control = cbuf;
    mh.msg_controllen = sizeof(cbuf);
    cmsghdr* cm = CMSG_FIRSTHDR(&mh);
    cm->cmsg_level = SOL_SOCKET;
    cm->cmsg_type = SCM_RIGHTS;
    cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    ssize_t n = sendmsg(sock, &mh, 0);
    if (n != (ssize_t)sizeof(m)) { set_err("sendmsg: %s", strerror(errno)); close(sock); return false; }
    // No acknowledgement: the message (and the in-flight fd, which holds its own reference to the
    // file) stays queued on the peer's not-yet-accepted connection after we close.  Waiting for an
    // ack here would deadlock, because every rank sends to all peers before it starts accepting.
    close(sock);
    return true;
}

// receive one message from the listening socket into the stash
bool recv_one(Symm* s, int timeout_ms) {
    pollfd pfd{s->listen_fd, POLLIN, 0};
    int pr = poll(&pfd, 1, timeout_ms);
    if (pr <= 0) { set_err("timeout waiting for a peer handle (rank %d)", s->rank); return false; }
    int conn = accept(s->listen_fd, nullptr, nullptr);
    if (conn < 0) { set_err("accept: %s", strerror(errno)); return false; }
    Msg m{0, 0};
    iovec iov{&m, sizeof(m)};
    char cbuf[CMSG_SPACE(sizeof(int))];
    msghdr mh;
    memset(&mh, 0, sizeof(mh));
    mh.msg_iov = &iov;
    mh.msg_iovlen = 1;
    mh.msg_control = cbuf;
    mh.msg_controllen = sizeof(cbuf);
    ssize_t n = recvmsg(conn, &mh, MSG_WAITALL);
    if (n != (ssize_t)sizeof(m)) { set_err("recvmsg: %s", strerror(errno)); close(conn); return false; }
    int fd = -1;
    for (cmsghdr* cm = CMSG_FIRSTHDR(&mh); cm; cm = CMSG_NXTHDR(&mh, cm))
        if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) memcpy(&fd, CMSG_DATA(cm), sizeof(int));
    close(conn);
    if (fd < 0) { set_err("message without fd"); return false; }
    s->stash[{m.tag, m.src}] = fd;
    return true;
}

bool take_fd(Symm* s, int tag, int src, int* fd, int timeout_ms) {
    for (;;) {
        auto it = s->stash.find({tag, src});
        if (it != s->stash.end()) {
            *fd = it->second;
            s->stash.erase(it);
            return true;
        }
        if (!recv_one(s, timeout_ms)) return false;
    }
}

bool map_handle(Symm* s, CUmemGenericAllocationHandle h, CUdeviceptr* va) {
    if (!cu_ok(g_drv.p_cuMemAddressReserve(va, s->size, s->gran, 0, 0), "cuMemAddressReserve")) return false;
    if (!cu_ok(g_drv.p_cuMemMap(*va, s->size, 0, h, 0), "cuMemMap")) return false;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = s->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (!cu_ok(g_drv.p_cuMemSetAccess(*va, s->size, &acc, 1), "cuMemSetAccess")) return false;
    return true;
}

}  // namespace

extern "C" {

const char* tfy_symm_last_error() { return g_err; }

// Host-only self test of the fd exchange (no CUDA): every rank publishes a memfd holding its rank
// and must read back every peer's.  Run by tests/test_symm_fdx.py with several processes.
int tfy_fdx_selftest(int rank, int world, const char* sock_prefix, int timeout_ms) {
    Symm s;
    s.rank = rank; s.world = world; s.sock_prefix = sock_prefix;
    char name[64];
    snprintf(name, sizeof(name), "/tmp/tfy_fdx_%d_%d", (int)getpid(), rank);
    int fd = open(name, O_RDWR | O_CREAT | O_TRUNC, 0600);
    if (fd < 0) { set_err("open: %s", strerror(errno)); return -1; }
    unlink(name);
    int32_t payload = 1000 + rank;
    if (write(fd, &payload, 4) != 4) { set_err("write"); return -1; }
    s.listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un addr;
    memset(&addr, 0, sizeof(addr));
    addr.sun_family = AF_UNIX;
    std::string p = sock_path(&s, rank);
    unlink(p.c_str());
    strncpy(addr.sun_path, p.c_str(), sizeof(addr.sun_path) - 1);
    if (bind(s.listen_fd, (sockaddr*)&addr, sizeof(addr)) != 0 || listen(s.listen_fd, 4 * TFY_MAX_RANKS) != 0) {
        set_err("bind/listen %s: %s", p.c_str(), strerror(errno));
        return -1;
    }
    int rc = 0;
    for (int r = 0; r < world && rc == 0; ++r)
        if (r != rank && !send_fd(&s, r, TAG_MEM, fd)) rc = -1;
    for (int r = 0; r < world && rc == 0; ++r) {
        if (r == rank) continue;
        int pfd = -1;
        if (!take_fd(&s, TAG_MEM, r, &pfd, timeout_ms)) { rc = -1; break; }
        int32_t got = 0;
        if (pread(pfd, &got, 4, 0) != 4 || got != 1000 + r) { set_err("bad payload from %d: %d", r, got); rc = -2; }
        close(pfd);
    }
    close(fd);
    close(s.listen_fd);
    unlink(p.c_str());
    return rc;
}

// Step 1: allocate + map the local arena, export it, start listening.
// Returns an opaque handle or nullptr.
void* tfy_symm_open(int device, int rank, int world, size_t size, const char* sock_prefix) {
    if (world > TFY_MAX_RANKS) { set_err("world %d > %d", world, TFY_MAX_RANKS); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {
        set_err("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    if (!load_driver()) return nullptr;
    Symm* s = new Symm();
    s->device = device; s->rank = rank; s->world = world;
    s->sock_prefix = sock_prefix ? sock_prefix : "";

    CUdevice dev;
    if (!cu_ok(g_drv.p_cuDeviceGet(&dev, device), "cuDeviceGet")) { delete s; return nullptr; }
    int mc = 0;
    g_drv.p_cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    s->mc_supported = (mc != 0) && world > 1;

    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (!cu_ok(g_drv.p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
               "cuMemGetAllocationGranularity")) { delete s; return nullptr; }
    if (s->mc_supported) {
        CUmulticastObjectProp mp;
        memset(&mp, 0, sizeof(mp));
        mp.numDevices = world;
        mp.size = size;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (g_drv.p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
            mg > gran)
            gran = mg;
    }
    s->gran = gran;
    s->size = (size + gran - 1) / gran * gran;
    if (!cu_ok(g_drv.p_cuMemCreate(&s->local_h, s->size, &prop, 0), "cuMemCreate")) { delete s; return nullptr; }
    s->peer_h[rank] = s->local_h;
    if (!map_handle(s, s->local_h, &s->peer_va[rank])) { delete s; return nullptr; }
    if (cudaMemset((void*)s->peer_va[rank], 0, s->size) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        set_err("memset of arena failed");
        delete s;
        return nullptr;
    }
    if (world > 1) {
        if (!cu_ok(g_drv.p_cuMemExportToShareableHandle(&s->local_fd, s->local_h,
                                                        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
                   "cuMemExportToShareableHandle")) { delete s; return nullptr; }
        s->listen_fd = socket(AF_UNIX, SOCK_STREAM, 0);
        sockaddr_un addr;
        memset(&addr, 0, sizeof(addr));
        addr.sun_family = AF_UNIX;
        std::string p = sock_path(s, rank);
        unlink(p.c_str());
        strncpy(addr.sun_path, p.c_str(), sizeof(addr.sun_path) - 1);
        if (bind(s->listen_fd, (sockaddr*)&addr, sizeof(addr)) != 0 || listen(s->listen_fd, 4 * TFY_MAX_RANKS) != 0) {
            set_err("bind/listen %s: %s", p.c_str(), strerror(errno));
            delete s;
            return nullptr;
        }
    }
    return s;
}

// Step 2 (after a job-wide barrier: every rank is listening): swap arena fds.
int tfy_symm_exchange(void* h, int timeout_ms) {
    Symm* s = (Symm*)h;
    for (int r = 0; r < s->world; ++r)
        if (r != s->rank && !send_fd(s, r, TAG_MEM, s->local_fd)) return -1;
    for (int r = 0; r < s->world; ++r) {
        if (r == s->rank) continue;
        int fd = -1;
        if (!take_fd(s, TAG_MEM, r, &fd, timeout_ms)) return -1;
        if (!cu_ok(g_drv.p_cuMemImportFromShareableHandle(&s->peer_h[r], (void*)(uintptr_t)fd,
                                                          CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
                   "cuMemImportFromShareableHandle")) return -1;
        close(fd);
        if (!map_handle(s, s->peer_h[r], &s->peer_va[r])) return -1;
    }
    return 0;
}

int tfy_symm_mc_supported(void* h) { return ((Symm*)h)->mc_supported ? 1 : 0; }

// Step 3: create / import the multicast object and add this device.
// Returns 0 ok, 1 multicast unavailable (not an error), -1 error.
int tfy_symm_mc_create(void* h, int timeout_ms) {
    Symm* s = (Symm*)h;
    if (!s->mc_supported) return 1;
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = s->world;
    mp.size = s->size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (s->rank == 0) {
        CUresult r = g_drv.p_cuMulticastCreate(&s->mc_h, &mp);
        int fd = -1;
        if (r == CUDA_SUCCESS)
            r = g_drv.p_cuMemExportToShareableHandle(&fd, s->mc_h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
        if (r != CUDA_SUCCESS) {
            cu_ok(r, "cuMulticastCreate/export");
            // tell the peers there is no multicast: send /dev/null with tag MC; they detect import failure
            int nfd = open("/dev/null", O_RDONLY);
            for (int p = 1; p < s->world; ++p) send_fd(s, p, TAG_MC, nfd);
            close(nfd);
            s->mc_supported = false;
            s->mc_h = 0;
            return 1;
        }
        for (int p = 1; p < s->world; ++p)
            if (!send_fd(s, p, TAG_MC, fd)) return -1;
        close(fd);
    } else {
        int fd = -1;
        if (!take_fd(s, TAG_MC, 0, &fd, timeout_ms)) return -1;
        CUresult r = g_drv.p_cuMemImportFromShareableHandle(&s->mc_h, (void*)(uintptr_t)fd,
                                                            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(fd);
        if (r != CUDA_SUCCESS) {
            s->mc_supported = false;
            s->mc_h = 0;
            return 1;
        }
    }
    CUdevice dev;
    g_drv.p_cuDeviceGet(&dev, s->device);
    if (!cu_ok(g_drv.p_cuMulticastAddDevice(s->mc_h, dev), "cuMulticastAddDevice")) return -1;
    return 0;
}

// Step 4 (after a barrier: every device was added): bind + map the multicast VA.
int tfy_symm_mc_bind(void* h) {
    Symm* s = (Symm*)h;
    if (!s->mc_supported || !s->mc_h) return 1;
    if (!cu_ok(g_drv.p_cuMulticastBindMem(s->mc_h, 0, s->local_h, 0, s->size, 0), "cuMulticastBindMem")) return -1;
    s->mc_bound = true;
    if (!map_handle(s, s->mc_h, &s->mc_va)) return -1;
    return 0;
}

uint64_t tfy_symm_peer_ptr(void* h, int r) { return (uint64_t)((Symm*)h)->peer_va[r]; }
uint64_t tfy_symm_mc_ptr(void* h) { return (uint64_t)((Symm*)h)->mc_va; }
uint64_t tfy_symm_size(void* h) { return (uint64_t)((Symm*)h)->size; }

void tfy_symm_close(void* h) {
    Symm* s = (Symm*)h;
    if (!s) return;
    cudaDeviceSynchronize();
    if (s->mc_va) {
        g_drv.p_cuMemUnmap(s->mc_va, s->size);
        g_drv.p_cuMemAddressFree(s->mc_va, s->size);
    }
    if (s->mc_bound) {
        CUdevice dev;
        g_drv.p_cuDeviceGet(&dev, s->device);
        g_drv.p_cuMulticastUnbind(s->mc_h, dev, 0, s->size);
    }
    if (s->mc_h) g_drv.p_cuMemRelease(s->mc_h);
    for (int r = 0; r < s->world; ++r) {
        if (s->peer_va[r]) {
            g_drv.p_cuMemUnmap(s->peer_va[r], s->size);
            g_drv.p_cuMemAddressFree(s->peer_va[r], s->size);
        }
        if (s->peer_h[r]) g_drv.p_cuMemRelease(s->peer_h[r]);
    }
    if (s->local_fd >= 0) close(s->local_fd);
    if (s->listen_fd >= 0) {
        close(s->listen_fd);
        unlink(sock_path(s, s->rank).c_str());
    }
    for (auto& kv : s->stash) close(kv.second);
    delete s;
}

}  // extern "C"
