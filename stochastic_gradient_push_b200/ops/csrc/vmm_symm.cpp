// Symmetric memory through the CUDA virtual-memory-management API, with an NVSwitch MULTICAST view.
//
// The cudaMalloc + cudaIpc* rendezvous of bindings.cpp::symm_alloc cannot host a multicast object
// (VERDICT r1).  This allocator does what NVLS needs:
//
//   every rank   cuMemCreate (POSIX-fd exportable physical memory on its GPU), map it
//   exchange     the fds travel as integers over the c10d control plane; a peer duplicates them
//                out of the owner's process with pidfd_open + pidfd_getfd (no fabric / IMEX
//                daemon, no unix-socket protocol) and cuMemImportFromShareableHandle's them
//   unicast      every peer's allocation is mapped (cuMemAddressReserve / cuMemMap / cuMemSetAccess):
//                plain P2P loads / stores / cp.async.bulk over NVLink work as before
//   multicast    symmetric rank 0 creates the multicast object (cuMulticastCreate), everyone adds
//                its device and binds its physical memory (cuMulticastBindMem), then maps the
//                object: stores to that address land in EVERY GPU (multimem.st), loads can be
//                reduced inside the switch (multimem.ld_reduce)
//
// libcuda is not linked (the extension must import on CPU-only hosts): every driver entry point is
// resolved through cudaGetDriverEntryPoint at first use.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>

#include <sys/syscall.h>
#include <unistd.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace py = pybind11;

namespace {

template <typename Fn> Fn drv(const char* name)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess ||
        p == nullptr)
        throw std::runtime_error(std::string("driver entry point unavailable: ") + name);
    return reinterpret_cast<Fn>(p);
}

#define DRV(name) static auto name##_ = drv<decltype(&name)>(#name)

void ck(CUresult r, const char* what)
{
    if (r == CUDA_SUCCESS) return;
    static auto get_str = drv<decltype(&cuGetErrorString)>("cuGetErrorString");
    const char* s = nullptr;
    get_str(r, &s);
    throw std::runtime_error(std::string(what) + " failed: " + (s ? s : "unknown CUDA driver error") + " (" +
                             std::to_string((int)r) + ")");
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

CUmemAllocationProp alloc_prop(int device)
{
    CUmemAllocationProp p;
    std::memset(&p, 0, sizeof(p));
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return p;
}

// duplicate file descriptor `fd` of process `pid` into this process
int steal_fd(int pid, int fd)
{
    if (pid == (int)getpid()) return dup(fd);
    const int pidfd = (int)syscall(SYS_pidfd_open, pid, 0);
    if (pidfd < 0) throw std::runtime_error("pidfd_open failed (errno " + std::to_string(errno) + ")");
    const int got = (int)syscall(SYS_pidfd_getfd, pidfd, fd, 0);
    const int err = errno;
    close(pidfd);
    if (got < 0) throw std::runtime_error("pidfd_getfd failed (errno " + std::to_string(err) +
                                          "): cannot import the peer's memory handle");
    return got;
}

struct Mapping {                 // one mapped range; unmapped + released with the last tensor using it
    CUdeviceptr va = 0;
    size_t size = 0;
    CUmemGenericAllocationHandle handle = 0;
    bool own_handle = false;
    ~Mapping()
    {
        try {
            DRV(cuMemUnmap); DRV(cuMemAddressFree); DRV(cuMemRelease);
            if (va) { cuMemUnmap_(va, size); cuMemAddressFree_(va, size); }
            if (own_handle && handle) cuMemRelease_(handle);
        } catch (...) {}
    }
};

torch::Tensor tensor_over(std::shared_ptr<Mapping> m, size_t nbytes, int device)
{
    auto opts = torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, device);
    return torch::from_blob(reinterpret_cast<void*>(m->va), {(int64_t)nbytes}, [m](void*) mutable { m.reset(); }, opts);
}

std::shared_ptr<Mapping> map_handle(CUmemGenericAllocationHandle h, size_t size, size_t align, int device, bool own)
{
    DRV(cuMemAddressReserve); DRV(cuMemMap); DRV(cuMemSetAccess);
    auto m = std::make_shared<Mapping>();
    m->size = size;
    m->handle = h;
    m->own_handle = own;
    ck(cuMemAddressReserve_(&m->va, size, align, 0, 0), "cuMemAddressReserve");
    ck(cuMemMap_(m->va, size, 0, h, 0), "cuMemMap");
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    ck(cuMemSetAccess_(m->va, size, &acc, 1), "cuMemSetAccess");
    return m;
}

}  // namespace

// ---------------------------------------------------------------------------
// capability probe
// ---------------------------------------------------------------------------
static py::dict vmm_caps(int device)
{
    py::dict d;
    int mc = 0, posix = 0, fabric = 0;
    try {
        DRV(cuDeviceGet); DRV(cuDeviceGetAttribute);
        c10::cuda::CUDAGuard guard(device);
        cudaFree(0);
        CUdevice dev;
        ck(cuDeviceGet_(&dev, device), "cuDeviceGet");
        cuDeviceGetAttribute_(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
        cuDeviceGetAttribute_(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
        cuDeviceGetAttribute_(&fabric, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev);
    } catch (const std::exception& e) {
        d["error"] = std::string(e.what());
    }
    d["multicast"] = mc != 0;
    d["posix_fd"] = posix != 0;
    d["fabric"] = fabric != 0;
    return d;
}

// size a symmetric allocation must have so that it can also be bound to a multicast object
static int64_t vmm_padded_size(int64_t nbytes, int device, int world)
{
    DRV(cuMemGetAllocationGranularity); DRV(cuMulticastGetGranularity);
    CUmemAllocationProp p = alloc_prop(device);
    size_t g = 0;
    ck(cuMemGetAllocationGranularity_(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    size_t mg = g;
    if (world > 1) {
        CUmulticastObjectProp mp;
        std::memset(&mp, 0, sizeof(mp));
        mp.numDevices = (unsigned)world;
        mp.size = round_up((size_t)nbytes, g);
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        if (cuMulticastGetGranularity_(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS) mg = g;
    }
    return (int64_t)round_up((size_t)nbytes, mg > g ? mg : g);
}

// ---------------------------------------------------------------------------
// VmmBuffer: this rank's physical allocation + the views mapped so far
// ---------------------------------------------------------------------------
class VmmBuffer {
public:
    VmmBuffer(int64_t nbytes, int device, int world) : device_(device), world_(world)
    {
        c10::cuda::CUDAGuard guard(device);
        cudaFree(0);
        DRV(cuMemCreate); DRV(cuMemExportToShareableHandle);
        size_ = (size_t)vmm_padded_size(nbytes, device, world);
        nbytes_ = (size_t)nbytes;
        CUmemAllocationProp p = alloc_prop(device);
        ck(cuMemCreate_(&handle_, size_, &p, 0), "cuMemCreate");
        ck(cuMemExportToShareableHandle_(&fd_, handle_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
           "cuMemExportToShareableHandle");
        local_map_ = map_handle(handle_, size_, size_ >= (2u << 20) ? (2u << 20) : 0, device, /*own=*/true);
        cudaMemset(reinterpret_cast<void*>(local_map_->va), 0, size_);
        cudaDeviceSynchronize();
    }
    ~VmmBuffer()
    {
        if (fd_ >= 0) close(fd_);
        if (mc_fd_ >= 0) close(mc_fd_);
    }

    int fd() const { return fd_; }
    int pid() const { return (int)getpid(); }
    int64_t size() const { return (int64_t)size_; }
    torch::Tensor local() { return tensor_over(local_map_, nbytes_, device_); }

    // map the allocation of a peer process (its pid + the fd number it exported)
    torch::Tensor open_peer(int pid, int fd)
    {
        c10::cuda::CUDAGuard guard(device_);
        DRV(cuMemImportFromShareableHandle);
        const int mine = steal_fd(pid, fd);
        CUmemGenericAllocationHandle h;
        CUresult r = cuMemImportFromShareableHandle_(&h, reinterpret_cast<void*>((uintptr_t)mine),
                                                     CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(mine);
        ck(r, "cuMemImportFromShareableHandle");
        auto m = map_handle(h, size_, size_ >= (2u << 20) ? (2u << 20) : 0, device_, /*own=*/true);
        return tensor_over(m, nbytes_, device_);
    }

    // --- multicast: symmetric rank 0 creates, everyone imports / adds / binds / maps ---------
    int mc_create()
    {
        c10::cuda::CUDAGuard guard(device_);
        DRV(cuMulticastCreate); DRV(cuMemExportToShareableHandle);
        CUmulticastObjectProp mp;
        std::memset(&mp, 0, sizeof(mp));
        mp.numDevices = (unsigned)world_;
        mp.size = size_;
        mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        ck(cuMulticastCreate_(&mc_, &mp), "cuMulticastCreate");
        ck(cuMemExportToShareableHandle_(&mc_fd_, mc_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
           "cuMemExportToShareableHandle(multicast)");
        have_mc_ = true;
        return mc_fd_;
    }
    void mc_import(int pid, int fd)
    {
        c10::cuda::CUDAGuard guard(device_);
        DRV(cuMemImportFromShareableHandle);
        const int mine = steal_fd(pid, fd);
        CUresult r = cuMemImportFromShareableHandle_(&mc_, reinterpret_cast<void*>((uintptr_t)mine),
                                                     CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        close(mine);
        ck(r, "cuMemImportFromShareableHandle(multicast)");
        have_mc_ = true;
    }
    void mc_add_device()
    {
        c10::cuda::CUDAGuard guard(device_);
        DRV(cuMulticastAddDevice); DRV(cuDeviceGet);
        TORCH_CHECK(have_mc_, "no multicast object");
        CUdevice dev;
        ck(cuDeviceGet_(&dev, device_), "cuDeviceGet");
        ck(cuMulticastAddDevice_(mc_, dev), "cuMulticastAddDevice");
    }
    // call after EVERY rank has added its device
    torch::Tensor mc_bind_and_map()
    {
        c10::cuda::CUDAGuard guard(device_);
        DRV(cuMulticastBindMem);
        TORCH_CHECK(have_mc_, "no multicast object");
        ck(cuMulticastBindMem_(mc_, 0, handle_, 0, size_, 0), "cuMulticastBindMem");
        mc_map_ = map_handle(mc_, size_, size_ >= (2u << 20) ? (2u << 20) : 0, device_, /*own=*/true);
        return tensor_over(mc_map_, nbytes_, device_);
    }

private:
    int device_, world_;
    size_t size_ = 0, nbytes_ = 0;
    CUmemGenericAllocationHandle handle_ = 0, mc_ = 0;
    bool have_mc_ = false;
    int fd_ = -1, mc_fd_ = -1;
    std::shared_ptr<Mapping> local_map_, mc_map_;
};

void bind_vmm(py::module& mod)
{
    mod.def("vmm_caps", &vmm_caps, "multicast / handle-type capabilities of a device");
    mod.def("vmm_padded_size", &vmm_padded_size);
    py::class_<VmmBuffer, std::shared_ptr<VmmBuffer>>(mod, "VmmBuffer")
        .def(py::init<int64_t, int, int>(), py::arg("nbytes"), py::arg("device"), py::arg("world"))
        .def("fd", &VmmBuffer::fd)
        .def("pid", &VmmBuffer::pid)
        .def("size", &VmmBuffer::size)
        .def("local", &VmmBuffer::local)
        .def("open_peer", &VmmBuffer::open_peer)
        .def("mc_create", &VmmBuffer::mc_create)
        .def("mc_import", &VmmBuffer::mc_import)
        .def("mc_add_device", &VmmBuffer::mc_add_device)
        .def("mc_bind_and_map", &VmmBuffer::mc_bind_and_map);
}
