// peer_probe.cu -- microbenchmark: random 32-byte loads from ANOTHER GPU's memory over NVLink (the access
// pattern of the peer-probed partitioned table).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_probe peer_probe.cu
// usage: peer_probe [table_GB=4] [loads_per_thread=64]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CU(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; return x ^ (x >> 33); }

template <int MODE, int ILP>
__global__ void probe(const uint4* __restrict__ tbl, uint64_t nbuckets, int iters, uint64_t* out) {
    uint64_t s = mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
    uint64_t acc = 0;
    for (int it = 0; it < iters; it++) {
        unsigned long long a[ILP], b[ILP], c[ILP], d[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            s = mix(s + j + 1);
            const uint4* p = tbl + 2 * (s % nbuckets);
            if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[j]), "=l"(b[j]), "=l"(c[j]), "=l"(d[j]) : "l"(p));
            if (MODE == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[j]), "=l"(b[j]), "=l"(c[j]), "=l"(d[j]) : "l"(p));
            if (MODE == 2) { asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(a[j]), "=l"(b[j]) : "l"(p)); c[j] = d[j] = 0; }
            if (MODE == 3) asm volatile("ld.relaxed.sys.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[j]), "=l"(b[j]), "=l"(c[j]), "=l"(d[j]) : "l"(p));
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) acc += a[j] ^ b[j] ^ c[j] ^ d[j];
    }
    if (acc == 0x1234567) out[0] = acc;
}

template <int MODE, int ILP>
void run(const char* what, const uint4* tbl, uint64_t nb, int iters, uint64_t* out, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int grid = 148 * ctas_per_sm, block = 256;
    probe<MODE, ILP><<<grid, block>>>(tbl, nb, 4, out);
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(e0));
    probe<MODE, ILP><<<grid, block>>>(tbl, nb, iters, out);
    CU(cudaEventRecord(e1));
    CU(cudaDeviceSynchronize());
    float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
    const double n = (double)grid * block * iters * ILP;
    printf("%-44s ctas/SM %d ILP %d: %8.2f G loads/s  %7.1f GB/s (32 B each)  %.2f ms\n", what, ctas_per_sm, ILP, n / ms / 1e6, n * 32 / ms / 1e6, ms);
}

#include <string.h>
#include <unistd.h>
#include <cuda.h>
#include <sys/socket.h>
#include <sys/un.h>
#define CD(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { const char* m; cuGetErrorString(r, &m); printf("%s: %s\n", #x, m); exit(1); } } while (0)
static int send_fd(int sock, int fd) {
    char dummy = 'x', cbuf[CMSG_SPACE(sizeof(int))];
    iovec io{&dummy, 1};
    msghdr msg{}; msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = cbuf; msg.msg_controllen = sizeof(cbuf);
    cmsghdr* c = CMSG_FIRSTHDR(&msg); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return (int)sendmsg(sock, &msg, 0);
}
static int recv_fd(int sock) {
    char dummy, cbuf[CMSG_SPACE(sizeof(int))];
    iovec io{&dummy, 1};
    msghdr msg{}; msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = cbuf; msg.msg_controllen = sizeof(cbuf);
    if (recvmsg(sock, &msg, 0) <= 0) return -1;
    int fd; memcpy(&fd, CMSG_DATA(CMSG_FIRSTHDR(&msg)), sizeof(int));
    return fd;
}
template <int MODE, int ILP> void run(const char* what, const uint4* tbl, uint64_t nb, int iters, uint64_t* out, int ctas_per_sm);
// VMM allocation (cuMemCreate, 2 MB granularity) shared through a POSIX fd over a unix socket
int vmm_main(int argc, char** argv) {
    const size_t bytes = (size_t)2048 << 21;  // 4 GiB
    const size_t align = argc > 2 ? (size_t)atoll(argv[2]) << 20 : (size_t)2 << 20;  // VA alignment in MiB
    const uint64_t nb = bytes / 32;
    sockaddr_un addr{}; addr.sun_family = AF_UNIX; strcpy(addr.sun_path, "/tmp/peer_probe.sock");
    CU(cudaFree(0));
    if (!strcmp(argv[1], "vserver")) {
        CU(cudaSetDevice(1)); CU(cudaFree(0));
        CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = 1; prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CUmemGenericAllocationHandle h; CD(cuMemCreate(&h, bytes, &prop, 0));
        size_t g0 = 0, g1 = 0;
        cuMemGetAllocationGranularity(&g0, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
        cuMemGetAllocationGranularity(&g1, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        printf("granularity: minimum %zu, recommended %zu; VA alignment %zu\n", g0, g1, align);
        CUdeviceptr p; CD(cuMemAddressReserve(&p, bytes, align, 0, 0)); CD(cuMemMap(p, bytes, 0, h, 0));
        CUmemAccessDesc ad{}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; CD(cuMemSetAccess(p, bytes, &ad, 1));
        CU(cudaMemset((void*)p, 1, bytes)); CU(cudaDeviceSynchronize());
        int fd; CD(cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        int ls = socket(AF_UNIX, SOCK_STREAM, 0); unlink(addr.sun_path); bind(ls, (sockaddr*)&addr, sizeof(addr)); listen(ls, 1);
        int cs = accept(ls, nullptr, nullptr); send_fd(cs, fd);
        char done; read(cs, &done, 1);
        return 0;
    }
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    while (connect(s, (sockaddr*)&addr, sizeof(addr)) != 0) usleep(100000);
    int fd = recv_fd(s);
    CU(cudaSetDevice(0)); CU(cudaFree(0));
    CUmemGenericAllocationHandle h; CD(cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr p; CD(cuMemAddressReserve(&p, bytes, align, 0, 0)); CD(cuMemMap(p, bytes, 0, h, 0));
    CUmemAccessDesc ad{}; ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = 0; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CD(cuMemSetAccess(p, bytes, &ad, 1));
    uint64_t* out; CU(cudaMalloc(&out, 8));
    printf("---- REMOTE table of another PROCESS (cuMemCreate + POSIX fd), 4 GiB ----\n");
    run<0, 1>("ld.global.nc.L1::no_allocate.v4.u64", (const uint4*)p, nb, 64, out, 8);
    run<0, 4>("ld.global.nc.L1::no_allocate.v4.u64", (const uint4*)p, nb, 16, out, 8);
    run<1, 4>("ld.global.v4.u64", (const uint4*)p, nb, 16, out, 8);
    write(s, "d", 1);
    return 0;
}
// two processes: "server" allocates on GPU 1 and exports a CUDA IPC handle, "client" maps it on GPU 0 and probes it
int ipc_main(int argc, char** argv) {
    const double gb = 4.0;
    const uint64_t nb = (uint64_t)(gb * 1e9 / 32);
    if (!strcmp(argv[1], "server")) {
        uint4* mem; CU(cudaSetDevice(1)); CU(cudaMalloc(&mem, nb * 32)); CU(cudaMemset(mem, 1, nb * 32)); CU(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h; CU(cudaIpcGetMemHandle(&h, mem));
        FILE* f = fopen("/tmp/peer_probe.h", "wb"); fwrite(&h, sizeof(h), 1, f); fclose(f);
        rename("/tmp/peer_probe.h", "/tmp/peer_probe.handle");
        while (access("/tmp/peer_probe.done", F_OK) != 0) usleep(100000);
        return 0;
    }
    while (access("/tmp/peer_probe.handle", F_OK) != 0) usleep(100000);
    cudaIpcMemHandle_t h; FILE* f = fopen("/tmp/peer_probe.handle", "rb"); fread(&h, sizeof(h), 1, f); fclose(f);
    CU(cudaSetDevice(0));
    uint4* remote; uint64_t* out; CU(cudaMalloc(&out, 8));
    CU(cudaIpcOpenMemHandle((void**)&remote, h, cudaIpcMemLazyEnablePeerAccess));
    printf("---- REMOTE table of another PROCESS (CUDA IPC), %.1f GB ----\n", gb);
    run<0, 1>("ld.global.nc.L1::no_allocate.v4.u64", remote, nb, 64, out, 8);
    run<0, 4>("ld.global.nc.L1::no_allocate.v4.u64", remote, nb, 16, out, 8);
    run<1, 4>("ld.global.v4.u64", remote, nb, 16, out, 8);
    run<0, 1>("ld.global.nc.L1::no_allocate.v4.u64", remote, nb, 64, out, 4);
    f = fopen("/tmp/peer_probe.done", "wb"); fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && (!strcmp(argv[1], "server") || !strcmp(argv[1], "client"))) return ipc_main(argc, argv);
    if (argc > 1 && (!strcmp(argv[1], "vserver") || !strcmp(argv[1], "vclient"))) return vmm_main(argc, argv);
    const double gb = argc > 1 ? atof(argv[1]) : 4.0;
    const int iters = argc > 2 ? atoi(argv[2]) : 64;
    int ndev = 0; CU(cudaGetDeviceCount(&ndev));
    const uint64_t nb = (uint64_t)(gb * 1e9 / 32);
    uint4 *local = nullptr, *remote = nullptr; uint64_t* out;
    CU(cudaSetDevice(0)); CU(cudaMalloc(&local, nb * 32)); CU(cudaMemset(local, 1, nb * 32)); CU(cudaMalloc(&out, 8));
    if (ndev > 1) {
        int can = 0; CU(cudaDeviceCanAccessPeer(&can, 0, 1));
        printf("devices %d, 0 -> 1 peer access possible: %d\n", ndev, can);
        CU(cudaSetDevice(1)); CU(cudaMalloc(&remote, nb * 32)); CU(cudaMemset(remote, 1, nb * 32)); CU(cudaDeviceSynchronize());
        CU(cudaSetDevice(0)); CU(cudaDeviceEnablePeerAccess(1, 0));
    }
    for (int pass = 0; pass < (remote ? 2 : 1); pass++) {
        const uint4* t = pass ? remote : local;
        printf("---- %s table, %.1f GB ----\n", pass ? "REMOTE (peer)" : "local", gb);
        run<0, 1>("ld.global.nc.L1::no_allocate.v4.u64", t, nb, iters, out, 8);
        run<0, 4>("ld.global.nc.L1::no_allocate.v4.u64", t, nb, iters / 4, out, 8);
        run<1, 1>("ld.global.v4.u64", t, nb, iters, out, 8);
        run<1, 4>("ld.global.v4.u64", t, nb, iters / 4, out, 8);
        run<2, 4>("ld.global.v2.u64 (16 B)", t, nb, iters / 4, out, 8);
        run<3, 4>("ld.relaxed.sys.global.v4.u64", t, nb, iters / 4, out, 8);
        run<1, 4>("ld.global.v4.u64", t, nb, iters / 4, out, 2);
    }
    return 0;
}
