// download.cu -- device -> host copies of the CSR arrays.
//
// The caller of the reference-shaped entry points owns the destination arrays:
// H_sp%init / S_sp%init (src/tools/sparse_array_tools.f90:557-569, called from
// src/mat_els/hamiltonian.f90:137-139) allocate ordinary PAGEABLE memory.  A plain
// cudaMemcpyAsync into pageable memory is staged by the driver through one thread
// (a fraction of the link rate), so large copies into pageable memory are pipelined
// here: D2H into a ring of pinned bounce buffers on the context's stream, a few
// worker threads copy each landed chunk to its destination while the next chunks
// are in flight.  Pinned / registered destinations are copied directly.
#include <condition_variable>
#include <cstring>
#include <deque>
#include <thread>

#include "ctx.h"

struct bs2e_ctx::Stager {
    static constexpr int kBuf = 8;
    static constexpr size_t kChunk = (size_t)32 << 20;
    char* pinned[kBuf] = {};
    cudaEvent_t ev[kBuf] = {};
    bool busy[kBuf] = {};
    int next = 0;
    struct Task { int buf; char* dst; size_t bytes; };
    std::deque<Task> queue;
    int in_flight = 0;
    bool stop = false;
    std::mutex mu;
    std::condition_variable cv_work, cv_free;
    std::vector<std::thread> workers;
    int device = 0;
    std::string error;
};

namespace bs2e {

namespace {
constexpr size_t kStagedMin = (size_t)4 << 20;   // smaller copies: the driver's own staging is fine

bs2e_ctx::Stager* stager_get(bs2e_ctx* c)
{
    if (c->stager) return c->stager;
    std::unique_ptr<bs2e_ctx::Stager> s(new bs2e_ctx::Stager());
    s->device = c->device;
    for (int q = 0; q < bs2e_ctx::Stager::kBuf; ++q) {
        BS2E_CUDA(cudaHostAlloc(&s->pinned[q], bs2e_ctx::Stager::kChunk, cudaHostAllocDefault));
        BS2E_CUDA(cudaEventCreateWithFlags(&s->ev[q], cudaEventDisableTiming));
    }
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(2u, std::min(8u, nthr / 2));
    bs2e_ctx::Stager* sp = s.get();
    for (unsigned w = 0; w < nthr; ++w)
        s->workers.emplace_back([sp] {
            cudaSetDevice(sp->device);
            for (;;) {
                bs2e_ctx::Stager::Task t;
                {
                    std::unique_lock<std::mutex> lk(sp->mu);
                    sp->cv_work.wait(lk, [&] { return sp->stop || !sp->queue.empty(); });
                    if (sp->queue.empty()) return;
                    t = sp->queue.front();
                    sp->queue.pop_front();
                }
                const cudaError_t e = cudaEventSynchronize(sp->ev[t.buf]);
                if (e == cudaSuccess) std::memcpy(t.dst, sp->pinned[t.buf], t.bytes);
                {
                    std::lock_guard<std::mutex> lk(sp->mu);
                    if (e != cudaSuccess && sp->error.empty()) sp->error = cudaGetErrorString(e);
                    sp->busy[t.buf] = false;
                    --sp->in_flight;
                }
                sp->cv_free.notify_all();
            }
        });
    c->stager = s.release();
    return c->stager;
}

bool host_pinned(const void* p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
}  // namespace

void download_to_host(bs2e_ctx* c, void* dst, const void* d_src, size_t bytes)
{
    if (!bytes) return;
    cudaStream_t st = c->stream;
    const char* force = getenv("BS2E_PAGEABLE");   // "driver": leave pageable destinations to cudaMemcpyAsync (A/B)
    if (bytes < kStagedMin || host_pinned(dst) || (force && strcmp(force, "driver") == 0)) {
        BS2E_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        return;
    }
    bs2e_ctx::Stager* s = stager_get(c);
    const char* src = static_cast<const char*>(d_src);
    char* out = static_cast<char*>(dst);
    for (size_t off = 0; off < bytes; off += bs2e_ctx::Stager::kChunk) {
        const size_t n = std::min(bs2e_ctx::Stager::kChunk, bytes - off);
        int buf;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            buf = s->next;
            s->next = (s->next + 1) % bs2e_ctx::Stager::kBuf;
            s->cv_free.wait(lk, [&] { return !s->busy[buf]; });
            s->busy[buf] = true;
            ++s->in_flight;
        }
        BS2E_CUDA(cudaMemcpyAsync(s->pinned[buf], src + off, n, cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaEventRecord(s->ev[buf], st));
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->queue.push_back(bs2e_ctx::Stager::Task{buf, out + off, n});
        }
        s->cv_work.notify_one();
    }
}

// everything queued by download_to_host has reached its destination
void download_flush(bs2e_ctx* c)
{
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
    if (!c->stager) return;
    bs2e_ctx::Stager* s = c->stager;
    std::unique_lock<std::mutex> lk(s->mu);
    s->cv_free.wait(lk, [&] { return s->in_flight == 0; });
    if (!s->error.empty()) {
        const std::string e = s->error;
        s->error.clear();
        throw Error("download: " + e);
    }
}

void stager_destroy(bs2e_ctx* c)
{
    bs2e_ctx::Stager* s = c->stager;
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv_work.notify_all();
    for (auto& t : s->workers) t.join();
    for (int q = 0; q < bs2e_ctx::Stager::kBuf; ++q) {
        if (s->pinned[q]) cudaFreeHost(s->pinned[q]);
        if (s->ev[q]) cudaEventDestroy(s->ev[q]);
    }
    delete s;
    c->stager = nullptr;
}

}  // namespace bs2e
