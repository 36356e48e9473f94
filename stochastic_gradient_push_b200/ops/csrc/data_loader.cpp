// Host-side input pipeline helper (data/shards.py): gather rows of memory-mapped uint8 shards into
// one (pinned) batch buffer with several threads.  Pure CPU code; the GIL is released while it runs.
//
// The reference's loader does its per-sample work in DataLoader worker PROCESSES
// (gossip_sgd.py:539-583); with pre-decoded shards the per-batch work is a strided memcpy of
// B x 196 KB, which a handful of threads do at memory bandwidth without pickling anything.
#include <torch/extension.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace py = pybind11;

// out[i] = shard[shard_id[i]] row local_id[i]   (row_bytes each)
// shard_ptrs: base addresses of the mapped arrays (numpy `arr.ctypes.data`), one per shard
static void gather_rows_u8(std::vector<int64_t> shard_ptrs, std::vector<int64_t> shard_rows, int64_t row_bytes,
                           torch::Tensor shard_id, torch::Tensor local_id, torch::Tensor out, int n_threads)
{
    TORCH_CHECK(out.device().is_cpu() && out.scalar_type() == torch::kUInt8 && out.is_contiguous(), "out: contiguous uint8 CPU tensor");
    TORCH_CHECK(shard_id.device().is_cpu() && local_id.device().is_cpu());
    TORCH_CHECK(shard_id.scalar_type() == torch::kInt64 && local_id.scalar_type() == torch::kInt64);
    TORCH_CHECK(shard_id.is_contiguous() && local_id.is_contiguous());
    const int64_t n = shard_id.numel();
    TORCH_CHECK(local_id.numel() == n && row_bytes > 0 && out.numel() == n * row_bytes, "shape mismatch");
    TORCH_CHECK(shard_ptrs.size() == shard_rows.size(), "one row count per shard");
    const int64_t* sid = shard_id.data_ptr<int64_t>();
    const int64_t* lid = local_id.data_ptr<int64_t>();
    for (int64_t i = 0; i < n; ++i) {
        TORCH_CHECK(sid[i] >= 0 && sid[i] < (int64_t)shard_ptrs.size(), "shard index out of range");
        TORCH_CHECK(lid[i] >= 0 && lid[i] < shard_rows[(size_t)sid[i]], "row index out of range");
    }
    uint8_t* dst = out.data_ptr<uint8_t>();
    const int workers = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads < 1 ? 1 : n_threads, n));
    py::gil_scoped_release nogil;
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) return;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(shard_ptrs[(size_t)sid[i]]) + lid[i] * row_bytes;
            std::memcpy(dst + i * row_bytes, src, (size_t)row_bytes);
        }
    };
    if (workers == 1) { work(); return; }
    std::vector<std::thread> pool;
    pool.reserve((size_t)workers - 1);
    for (int t = 1; t < workers; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

void bind_data(py::module& mod)
{
    mod.def("gather_rows_u8", &gather_rows_u8, py::arg("shard_ptrs"), py::arg("shard_rows"), py::arg("row_bytes"),
            py::arg("shard_id"), py::arg("local_id"), py::arg("out"), py::arg("n_threads") = 4,
            "gather rows of memory-mapped uint8 shards into one batch buffer (multi-threaded memcpy)");
}
