// plan_dev.cu -- the plan of one symmetry block, built ON THE DEVICE from the
// configuration list (term%configs of src/tools/orbital_tools.f90:15-19 in the order of
// count_configs, :157-193).
//
// The reference's count_nnz (src/mat_els/hamiltonian.f90:348-416) scans all n_config^2
// pairs; here the list is reduced to its (l1,l2) group structure:
//   conf_scan_kernel    group boundaries, argument checks, largest n(2)        -> host (one small read-back)
//   [host]              exact 3j/6j tables of the group pairs, cached per context (plan.cpp)
//   ncrow_build_kernel  per group and n(1): range of n(2) and first configuration index
//   site_enum_kernel    radial sites (n1,n2) that carry planned rows, as sort keys
//   cub radix sort      sites with exchange windows first, heaviest first
//   count + scan        (block.cu) row counts -> 1-based index_ptr             -> host (totals)
// No per-row table is built or uploaded: the site kernels derive the rows of a site
// from the group tables.  Host work per plan is O(number of groups^2) on a cache miss
// and O(number of groups) otherwise.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <set>

#include "ctx.h"
#include "site_core.h"
#include "site_mma.h"

namespace bs2e {

namespace {

constexpr int kErrRangeN = 1, kErrOrderL = 2, kErrN2 = 4, kErrN1 = 8, kErrBounds = 16;
constexpr int kBoundCap = kMaxBlocks + 1;
// counters of a plan (device ints): [0] nsites, [1] nsites_x, [2] group boundaries found,
// [3] error bits, [4] largest n(2), [5] largest l(1)
// [6] n_b + 1 - (smallest first index a of an R^k row the sites read), [7] the largest such a
constexpr int kCntSites = 0, kCntSitesX = 1, kCntBounds = 2, kCntErr = 3, kCntMaxNd = 4, kCntLmax = 5, kCntALo = 6, kCntAHi = 7,
              kCounters = 8;

__global__ void conf_scan_kernel(long long n, const long long* __restrict__ cn, const long long* __restrict__ cl,
                                 int nb, int* __restrict__ counters, int4* __restrict__ bounds)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int err = 0, mnd = 0, ml = 0;
    if (i < n) {
        const long long n1 = cn[2 * i], n2 = cn[2 * i + 1], l1 = cl[2 * i], l2 = cl[2 * i + 1];
        if (n1 < 1 || n1 > nb || n2 < 1 || n2 > nb) err |= kErrRangeN;
        if (l2 < 0 || l1 < l2 || l1 > 120) err |= kErrOrderL;
        mnd = err ? 0 : (int)n2;
        ml = err ? 0 : (int)l1;
        const bool newblk = i == 0 || cl[2 * i - 2] != l1 || cl[2 * i - 1] != l2;
        if (newblk) {
            const int q = atomicAdd(&counters[kCntBounds], 1);
            if (q < kBoundCap) bounds[q] = make_int4((int)i, (int)l1, (int)l2, 0);
            else err |= kErrBounds;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        err |= __shfl_xor_sync(0xffffffffu, err, o);
        mnd = max(mnd, __shfl_xor_sync(0xffffffffu, mnd, o));
        ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (err) atomicOr(&counters[kCntErr], err);
        atomicMax(&counters[kCntMaxNd], mnd);
        atomicMax(&counters[kCntLmax], ml);
    }
}

__global__ void ncrow_init_kernel(size_t n, NcRow* __restrict__ ncrow)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ncrow[i] = NcRow{1, 0, 0, 0};
}

// group of configuration i (0-based): largest b with blk_start[b] <= i
__device__ __forceinline__ int group_of(const int* __restrict__ blk_start, int nblk, int i)
{
    int lo = 0, n = nblk;
    while (n > 1) {
        const int half = n >> 1;
        if (blk_start[lo + half] <= i) lo += half;
        n -= half;
    }
    return lo;
}

__global__ void ncrow_build_kernel(long long n, const long long* __restrict__ cn, int nblk,
                                   const int* __restrict__ blk_start, int stride, NcRow* __restrict__ ncrow,
                                   BlockDesc* __restrict__ blk, int* __restrict__ counters)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bi = group_of(blk_start, nblk, (int)i);
    const bool first = i == blk_start[bi], last = i + 1 == blk_start[bi + 1];
    const int n1 = (int)cn[2 * i], n2 = (int)cn[2 * i + 1];
    int err = 0;
    bool start = first, end = last;
    if (!first) {
        const int p1 = (int)cn[2 * i - 2], p2 = (int)cn[2 * i - 1];
        if (p1 != n1) {
            start = true;
            if (n1 < p1) err |= kErrN1;          // n(1) must ascend inside an (l1,l2) group
        } else if (n2 != p2 + 1) err |= kErrN2;  // n(2) consecutive inside an n(1) row
    }
    if (!last && (int)cn[2 * i + 2] != n1) end = true;
    NcRow* row = ncrow + (size_t)bi * stride + n1;
    if (start) { row->nd_lo = n2; row->start = (int)i + 1; }
    if (end) row->nd_hi = n2;
    if (first) blk[bi].nc_lo = n1;
    if (last) blk[bi].nc_hi = n1;
    if (err) atomicOr(&counters[kCntErr], err);
}

// one thread per (n_a, n_b): number of planned rows on the site
__global__ void site_enum_kernel(Geom g, Plan pl, int cap, unsigned long long* __restrict__ keys,
                                 int* __restrict__ counters)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int na = idx / pl.max_nd + 1, nb = idx % pl.max_nd + 1;
    if (na > g.nb) return;
    int cnt = 0;
    for (int bi = 0; bi < pl.nblk; ++bi) {
        const int row = config_index(g, pl, bi, na, nb);
        if (row > 0 && row_local_of(pl.rr, row) >= 0) ++cnt;
    }
    if (cnt == 0) return;
    const bool wantX = site_wants_X(g, pl.max_nd, na);
    const int q = atomicAdd(&counters[kCntSites], 1);
    if (wantX) atomicAdd(&counters[kCntSitesX], 1);
    if (q < cap) keys[q] = site_sort_key(wantX, cnt, na, nb);
    // rows of R^k the site reads: pairs (n_a, .) for the direct window, (n_b, .) for the exchange window (site_own_cand)
    const int alo = wantX ? imin(na, nb) : na, ahi = wantX ? imax(na, nb) : na;
    atomicMax(&counters[kCntALo], g.nb + 1 - alo);
    atomicMax(&counters[kCntAHi], ahi);
}

__global__ void row_tables_kernel(long long n, const long long* __restrict__ cn, int nblk,
                                  const int* __restrict__ blk_start, unsigned short* __restrict__ rn1,
                                  unsigned short* __restrict__ rn2, unsigned short* __restrict__ rblk)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rn1[i] = (unsigned short)cn[2 * i];
    rn2[i] = (unsigned short)cn[2 * i + 1];
    rblk[i] = (unsigned short)group_of(blk_start, nblk, (int)i);
}

void ensure_pin(bs2e_ctx* c, size_t bytes)
{
    if (c->h_pin_bytes >= bytes) return;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr;
    c->h_pin_bytes = 0;
    BS2E_CUDA(cudaHostAlloc(&c->h_pin, bytes, cudaHostAllocDefault));
    c->h_pin_bytes = bytes;
}

std::shared_ptr<AngDev> ang_tables(bs2e_ctx* c, const std::vector<BlockDesc>& blocks, int L, cudaStream_t st)
{
    std::vector<int> key;
    key.reserve(2 * blocks.size() + 2);
    key.push_back(L);
    key.push_back(c->hg.K1);
    for (const BlockDesc& b : blocks) { key.push_back(b.l1); key.push_back(b.l2); }
    auto it = c->ang_cache.find(key);
    if (it != c->ang_cache.end()) return it->second;
    auto a = std::make_shared<AngDev>();
    a->host = build_ang_tables(blocks, L, c->hg.K1);
    a->flags = dev_upload(a->host.flags, st);
    a->krange = dev_upload(a->host.krange, st);
    a->angD = dev_upload(a->host.angD, st);
    a->angX = dev_upload(a->host.angX, st);
    if (a->host.nkp > 0) a->angP = dev_upload(a->host.angP, st);
    BS2E_CUDA(cudaStreamSynchronize(st));   // the host vectors may be touched again only after the copies
    c->ang_cache.emplace(std::move(key), a);
    return a;
}

}  // namespace

void build_row_tables(cudaStream_t st, long long n_config, const long long* d_conf_n, int nblk,
                      const int* d_blk_start, unsigned short* row_n1, unsigned short* row_n2,
                      unsigned short* row_blk)
{
    if (n_config <= 0) return;
    row_tables_kernel<<<(unsigned)((n_config + 255) / 256), 256, 0, st>>>(n_config, d_conf_n, nblk, d_blk_start,
                                                                         row_n1, row_n2, row_blk);
    BS2E_LAUNCHED();
}

bs2e_configs* configs_upload(bs2e_ctx* c, long long n_config, const int64_t* conf_n, const int64_t* conf_l)
{
    if (n_config < 0) throw Error("bs2e_configs_upload: negative n_config");
    std::unique_ptr<bs2e_configs> cfg(new bs2e_configs());
    cfg->ctx = c;
    cfg->n = n_config;
    if (n_config > 0) {
        cfg->d_n = dev_alloc<long long>(2 * (size_t)n_config);
        cfg->d_l = dev_alloc<long long>(2 * (size_t)n_config);
        BS2E_CUDA(cudaMemcpyAsync(cfg->d_n, conf_n, sizeof(long long) * 2 * n_config, cudaMemcpyHostToDevice, c->stream));
        BS2E_CUDA(cudaMemcpyAsync(cfg->d_l, conf_l, sizeof(long long) * 2 * n_config, cudaMemcpyHostToDevice, c->stream));
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
    }
    return cfg.release();
}

void configs_free(bs2e_configs* cfg)
{
    if (!cfg) return;
    cudaFree(cfg->d_n);
    cudaFree(cfg->d_l);
    delete cfg;
}

// Plan buffers come in power-of-two size classes from a stream-ordered memory pool of their own -- NOT from cudaMalloc: with
// the CSR arrays of earlier blocks cached in the pool (release threshold = all), a cudaMalloc beside a running fill
// was measured at 50-120 ms (the driver trims the pool), once per new buffer, well into the timed steps
// (gpurun_out/r03a trace: block_plan 0.5 ms -> 118 ms).  `st` is the stream the buffer is used on first.
void arena_take(bs2e_ctx* c, DevArena& a, size_t bytes, cudaStream_t st)
{
    bytes += 256;
    size_t cls = 64 * 1024;
    while (cls < bytes) cls <<= 1;
    ArenaBuf* pick = nullptr;
    {
        // the smallest buffer of sufficient size whose last reader has finished; a new one while the pool is small
        // (waiting for a buffer that was released behind a running fill would serialise the plan with that fill);
        // else any free one
        std::lock_guard<std::mutex> lk(c->arena_mu);
        const int passes = c->arenas.size() < 24 ? 1 : 2;
        for (int pass = 0; pass < passes && !pick; ++pass)
            for (ArenaBuf* q : c->arenas) {
                if (q->busy || q->size < bytes || (pick && q->size >= pick->size)) continue;
                if (pass == 0 && cudaEventQuery(q->free_after) != cudaSuccess) { cudaGetLastError(); continue; }
                pick = q;
            }
        if (pick) pick->busy = true;
    }
    if (pick) {
        BS2E_CUDA(cudaEventSynchronize(pick->free_after));
    } else {
        std::unique_ptr<ArenaBuf> q(new ArenaBuf());
        q->size = cls;
        if (!c->arena_pool) {   // a pool of their own: plan buffers never split the free blocks of the CSR arrays
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = c->device;
            BS2E_CUDA(cudaMemPoolCreate(&c->arena_pool, &props));
            unsigned long long keep = ~0ull;
            BS2E_CUDA(cudaMemPoolSetAttribute(c->arena_pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        BS2E_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&q->base), q->size, c->arena_pool, st));
        BS2E_CUDA(cudaEventCreateWithFlags(&q->free_after, cudaEventDisableTiming));
        q->busy = true;
        pick = q.release();
        std::lock_guard<std::mutex> lk(c->arena_mu);
        c->arenas.push_back(pick);
    }
    a.buf = pick;
    a.used = 0;
}

void arena_give(bs2e_ctx* c, DevArena& a, cudaStream_t last_use)
{
    if (!a.buf) return;
    cudaEventRecord(a.buf->free_after, last_use);
    std::lock_guard<std::mutex> lk(c->arena_mu);
    a.buf->busy = false;
    a.buf = nullptr;
    a.used = 0;
}

GroupStructure::~GroupStructure() { cudaFree(d_mem); }

void ctx_release_plan_state(bs2e_ctx* c)
{
    c->ang_cache.clear();
    for (ArenaBuf* q : c->arenas) {
        cudaFree(q->base);
        if (q->free_after) cudaEventDestroy(q->free_after);
        delete q;
    }
    c->arenas.clear();
    if (c->arena_pool) { cudaMemPoolDestroy(c->arena_pool); c->arena_pool = nullptr; }
    if (c->plan_stream) { cudaStreamSynchronize(c->plan_stream); cudaStreamDestroy(c->plan_stream); c->plan_stream = nullptr; }
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr;
    c->h_pin_bytes = 0;
}

bool site_kernel_usable(const bs2e_ctx* c, int nblk, int lmax);   // block.cu

// The (l1,l2) group structure of a configuration list: group boundaries, per group and n(1) the range of n(2)
// and the first configuration index.  It depends on the list only, so a list kept resident on the device
// (bs2e_configs) carries it from its first plan on; a plan then costs the radial-site enumeration, the count
// pass and ONE host synchronisation.
std::shared_ptr<GroupStructure> build_structure(bs2e_ctx* c, long long n_config, const long long* d_conf_n,
                                                const long long* d_conf_l, cudaStream_t st)
{
    const Geom& hg = c->hg;
    auto gs = std::make_shared<GroupStructure>();
    gs->ctx = c;
    const int stride = hg.nb + 1;
    // counters + boundary list first (their size does not depend on the number of groups)
    DevArena tmp;
    arena_take(c, tmp, DevArena::need(sizeof(int4) * kBoundCap) + DevArena::need(sizeof(int) * kCounters) + 1024, st);
    struct Give { bs2e_ctx* c; DevArena& a; cudaStream_t st; ~Give() { arena_give(c, a, st); } } give{c, tmp, st};
    int4* d_bounds = tmp.take<int4>(kBoundCap);
    int* d_counters = tmp.take<int>(kCounters);
    ensure_pin(c, 64 + sizeof(int4) * kBoundCap + sizeof(BlockDesc) * kBoundCap + sizeof(int) * (kBoundCap + 1) + 256);
    BS2E_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(int) * kCounters, st));
    conf_scan_kernel<<<(unsigned)((n_config + 255) / 256), 256, 0, st>>>(n_config, d_conf_n, d_conf_l, hg.nb, d_counters,
                                                                        d_bounds);
    BS2E_LAUNCHED();
    int* h_cnt = c->h_pin;
    int4* h_bounds = reinterpret_cast<int4*>(c->h_pin + 16);
    BS2E_CUDA(cudaMemcpyAsync(h_cnt, d_counters, sizeof(int) * kCounters, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(int4) * kBoundCap, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
    {
        const int err = h_cnt[kCntErr];
        if (err & kErrRangeN) throw Error("block_plan: configuration n outside 1..n_b");
        if (err & kErrOrderL) throw Error("block_plan: configurations must have l(1) >= l(2) >= 0 (count_configs order)");
        if ((err & kErrBounds) || h_cnt[kCntBounds] > kMaxBlocks) throw Error("block_plan: too many (l1,l2) blocks");
    }
    const int nblk = h_cnt[kCntBounds];
    gs->nblk = nblk;
    gs->max_nd = h_cnt[kCntMaxNd];
    gs->lmax = h_cnt[kCntLmax];
    std::vector<int4> bounds(h_bounds, h_bounds + nblk);
    std::sort(bounds.begin(), bounds.end(), [](const int4& a, const int4& q) { return a.x < q.x; });
    gs->blocks.resize(nblk);
    std::vector<int> blk_start(nblk + 1);
    {
        std::set<std::pair<int, int>> seen;
        for (int q = 0; q < nblk; ++q) {
            gs->blocks[q] = BlockDesc{bounds[q].y, bounds[q].z, 0, 0};
            blk_start[q] = bounds[q].x;
            if (!seen.insert({bounds[q].y, bounds[q].z}).second)
                throw Error("block_plan: configurations of one (l1,l2) pair are not contiguous");
            if (((bounds[q].y + bounds[q].z) & 1) != ((bounds[0].y + bounds[0].z) & 1))
                throw Error("block_plan: configurations of both parities in one symmetry block");
        }
        blk_start[nblk] = (int)n_config;
    }
    {
        const size_t b0 = DevArena::need(sizeof(BlockDesc) * nblk), b1 = DevArena::need(sizeof(int) * (nblk + 1));
        BS2E_CUDA(cudaMalloc(&gs->d_mem, b0 + b1 + DevArena::need(sizeof(NcRow) * (size_t)nblk * stride)));
        gs->d_blk = reinterpret_cast<BlockDesc*>(gs->d_mem);
        gs->d_blk_start = reinterpret_cast<int*>(gs->d_mem + b0);
        gs->d_ncrow = reinterpret_cast<NcRow*>(gs->d_mem + b0 + b1);
    }
    {   // small host -> device tables through the pinned staging buffer
        char* hp = reinterpret_cast<char*>(c->h_pin);
        std::memcpy(hp, gs->blocks.data(), sizeof(BlockDesc) * nblk);
        BS2E_CUDA(cudaMemcpyAsync(gs->d_blk, hp, sizeof(BlockDesc) * nblk, cudaMemcpyHostToDevice, st));
        char* hp2 = hp + ((sizeof(BlockDesc) * nblk + 15) & ~(size_t)15);
        std::memcpy(hp2, blk_start.data(), sizeof(int) * (nblk + 1));
        BS2E_CUDA(cudaMemcpyAsync(gs->d_blk_start, hp2, sizeof(int) * (nblk + 1), cudaMemcpyHostToDevice, st));
    }
    const size_t ncells = (size_t)nblk * stride;
    ncrow_init_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(ncells, gs->d_ncrow);
    BS2E_LAUNCHED();
    BS2E_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(int) * kCounters, st));
    ncrow_build_kernel<<<(unsigned)((n_config + 255) / 256), 256, 0, st>>>(n_config, d_conf_n, nblk, gs->d_blk_start, stride,
                                                                          gs->d_ncrow, gs->d_blk, d_counters);
    BS2E_LAUNCHED();
    BS2E_CUDA(cudaMemcpyAsync(h_cnt, d_counters, sizeof(int) * kCounters, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
    {
        const int err = h_cnt[kCntErr];
        if (err & kErrN1) throw Error("block_plan: n(1) not ascending inside an (l1,l2) block");
        if (err & kErrN2) throw Error("block_plan: n(2) not consecutive inside an n(1) row");
    }
    return gs;
}

bs2e_block* block_plan(bs2e_ctx* c, int L, long long n_config, const int64_t* conf_n, const int64_t* conf_l,
                       const bs2e_configs* cfg, int full, long long n_ranges, const int64_t* range_lo,
                       const int64_t* range_hi)
{
    std::lock_guard<std::mutex> lk(c->plan_mu);
    const Geom& hg = c->hg;
    if (cfg) {
        if (cfg->ctx != c) throw Error("block_plan: configuration list of another context");
        n_config = cfg->n;
    }
    if (n_config < 0) throw Error("block_plan: negative n_config");
    if (n_config > 2147483000LL) throw Error("block_plan: n_config exceeds 32-bit row indices");
    if (hg.nb > 65535) throw Error("block_plan: n_b exceeds 16-bit storage");
    if (L < 0 || L > 255) throw Error("block_plan: L out of range");
    std::unique_ptr<bs2e_block, void (*)(bs2e_block*)> guard(new bs2e_block(), block_free);
    bs2e_block* b = guard.get();
    b->ctx = c;
    b->L = L;
    b->full = full ? 1 : 0;
    b->n_config = n_config;
    Plan& pl = b->dplan;
    pl.n_config = (int)n_config;
    pl.full = b->full;
    pl.L = L;
    if (n_config == 0) {   // a symmetry without configurations: empty CSR blocks, index_ptr = [1]
        b->nrows = 0;
        return guard.release();
    }
    std::vector<int> rlo, rhi, roff;
    int nrows = 0;
    try {
        if (n_ranges < 1) throw std::invalid_argument("block_plan: no row range given");
        check_row_ranges(n_config, n_ranges, range_lo, range_hi, rlo, rhi, roff, &nrows);
    } catch (const std::invalid_argument& e) {
        throw Error(e.what());
    }
    b->nrows = nrows;
    if (!c->plan_stream) {   // highest priority: the small plan kernels must not queue behind the CTAs of a running fill
        int lo = 0, hi = 0;
        BS2E_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        BS2E_CUDA(cudaStreamCreateWithPriority(&c->plan_stream, cudaStreamNonBlocking, hi));
    }
    cudaStream_t st = c->plan_stream;   // not the main stream: the plan overlaps the fill of the previous block
    const int nr = (int)rlo.size();

    // ---- the configuration list on the device and its group structure ----
    if (cfg) {
        b->d_conf_n = cfg->d_n;
        b->d_conf_l = cfg->d_l;
        if (!cfg->gs) cfg->gs = build_structure(c, n_config, cfg->d_n, cfg->d_l, st);
        b->gs = cfg->gs;
    } else {
        if (!conf_n || !conf_l) throw Error("block_plan: null configuration arrays");
        arena_take(c, b->arena0, 2 * DevArena::need(sizeof(long long) * 2 * n_config) + 1024, st);
        long long* dn = b->arena0.take<long long>(2 * (size_t)n_config);
        long long* dl = b->arena0.take<long long>(2 * (size_t)n_config);
        BS2E_CUDA(cudaMemcpyAsync(dn, conf_n, sizeof(long long) * 2 * n_config, cudaMemcpyHostToDevice, st));
        BS2E_CUDA(cudaMemcpyAsync(dl, conf_l, sizeof(long long) * 2 * n_config, cudaMemcpyHostToDevice, st));
        b->d_conf_n = dn;
        b->d_conf_l = dl;
        b->gs = build_structure(c, n_config, dn, dl, st);
    }
    const GroupStructure& gs = *b->gs;
    const int nblk = gs.nblk, max_nd = gs.max_nd;
    b->lmax = gs.lmax;
    b->d_blk_start = gs.d_blk_start;
    b->ang = ang_tables(c, gs.blocks, L, st);

    // ---- arena 1: site keys, counters, row ranges, count / pointer arrays, scan and sort scratch ----
    const int site_cap = hg.nb * std::max(1, max_nd);
    b->site_cap = site_cap;
    b->use_site = site_kernel_usable(c, nblk, b->lmax);
    {   // Two site kernels fill the same CSR arrays: the tensor-core kernel (site_mma.cu) and the FMA kernel
        // (block.cu).  Measured on B200 (scripts/fill_ab.py, profiles/r02v_fill_ncu_summary.md): the tensor-core kernel is
        // ahead from 8 multipoles on (cfg3: 5.85 against 6.80 ms, cfg4: 41.3 against 42.6 ms); the FMA kernel stays
        // the choice for max_k <= 6, where a site is small (cfg1).  BS2E_FILL=mma / fma forces one of them.
        const char* mode = getenv("BS2E_FILL");
        const bool want = mode ? strcmp(mode, "mma") == 0 : site_kmax_for(hg.K1) >= 13;
        b->use_mma = b->use_site && want && !(mode && strcmp(mode, "fma") == 0) &&
                     site_mma_usable(c, nblk, b->ang->host.maxc, b->ang->host.maxrec, b->lmax);
    }
    size_t scan_tmp = 0, sort_tmp = 0;
    BS2E_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, scan_tmp, (long long*)nullptr, (long long*)nullptr, cub::Sum(), 1LL,
                                             (long long)nrows + 1, st));   // same index type as block_count_scan
    BS2E_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const unsigned long long*)nullptr,
                                             (unsigned long long*)nullptr, site_cap, 0, 43, st));
    {
        size_t bytes = DevArena::need(sizeof(int) * kCounters) + DevArena::need(sizeof(int) * 3 * nr) +
                       2 * DevArena::need(sizeof(unsigned long long) * site_cap) +
                       4 * DevArena::need(sizeof(long long) * ((size_t)nrows + 1)) + DevArena::need(scan_tmp) +
                       DevArena::need(sort_tmp) + 1024;
        if (!b->use_site) bytes += 3 * DevArena::need(sizeof(unsigned short) * n_config);
        arena_take(c, b->arena1, bytes, st);
    }
    b->d_counters = b->arena1.take<int>(kCounters);
    int* d_ranges = b->arena1.take<int>(3 * (size_t)nr);
    unsigned long long* d_keys_raw = b->arena1.take<unsigned long long>(site_cap);
    b->d_site_key = b->arena1.take<unsigned long long>(site_cap);
    b->d_cntH = b->arena1.take<long long>((size_t)nrows + 1);
    b->d_cntS = b->arena1.take<long long>((size_t)nrows + 1);
    b->d_Hptr = b->arena1.take<long long>((size_t)nrows + 1);
    b->d_Sptr = b->arena1.take<long long>((size_t)nrows + 1);
    b->d_scan_tmp = b->arena1.take<char>(scan_tmp);
    b->scan_tmp_bytes = scan_tmp;
    void* d_sort_tmp = b->arena1.take<char>(sort_tmp);
    ensure_pin(c, 64 + sizeof(int) * 3 * (size_t)nr + 256 + sizeof(int4) * kBoundCap + sizeof(BlockDesc) * kBoundCap +
                      sizeof(int) * (kBoundCap + 1));
    {   // the row ranges through the pinned staging buffer (behind the 64 bytes of the read-backs)
        int* hp = c->h_pin + 16;
        std::memcpy(hp, rlo.data(), sizeof(int) * nr);
        std::memcpy(hp + nr, rhi.data(), sizeof(int) * nr);
        std::memcpy(hp + 2 * nr, roff.data(), sizeof(int) * nr);
        BS2E_CUDA(cudaMemcpyAsync(d_ranges, hp, sizeof(int) * 3 * nr, cudaMemcpyHostToDevice, st));
    }
    BS2E_CUDA(cudaMemsetAsync(b->d_counters, 0, sizeof(int) * kCounters, st));
    pl.nblk = nblk;
    pl.max_nd = max_nd;
    pl.blk = gs.d_blk;
    pl.ncrow = gs.d_ncrow;
    pl.flags = b->ang->flags;
    pl.krange = b->ang->krange;
    pl.angD = b->ang->angD;
    pl.angX = b->ang->angX;
    pl.angP = b->ang->angP;
    pl.nkp = b->ang->host.nkp;
    pl.nrows = nrows;
    pl.rr = RowRanges{nr, d_ranges, d_ranges + nr, d_ranges + 2 * nr, rlo[0], rhi[0]};
    if (b->use_site) {
        BS2E_CUDA(cudaMemsetAsync(d_keys_raw, 0xff, sizeof(unsigned long long) * site_cap, st));
        site_enum_kernel<<<(unsigned)((site_cap + 127) / 128), 128, 0, st>>>(c->dg, pl, site_cap, d_keys_raw, b->d_counters);
        BS2E_LAUNCHED();
        BS2E_CUDA(cub::DeviceRadixSort::SortKeys(d_sort_tmp, sort_tmp, d_keys_raw, b->d_site_key, site_cap, 0, 43, st));
        g_launches.fetch_add(3);
    } else {
        unsigned short* rn1 = b->arena1.take<unsigned short>(n_config);
        unsigned short* rn2 = b->arena1.take<unsigned short>(n_config);
        unsigned short* rblk = b->arena1.take<unsigned short>(n_config);
        build_row_tables(st, n_config, b->d_conf_n, nblk, gs.d_blk_start, rn1, rn2, rblk);
        pl.row_n1 = rn1;
        pl.row_n2 = rn2;
        pl.row_blk = rblk;
    }
    // count pass + scan, then one read-back: number of sites, totals
    {
        const BlockStreams ps{st, st, nullptr, nullptr};
        block_count_scan(b, false, &ps);
    }
    int* h_cnt = c->h_pin;
    long long* h_tot = reinterpret_cast<long long*>(c->h_pin + 8);
    BS2E_CUDA(cudaMemcpyAsync(h_cnt, b->d_counters, sizeof(int) * kCounters, cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaMemcpyAsync(h_tot, b->d_Hptr + nrows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaMemcpyAsync(h_tot + 1, b->d_Sptr + nrows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
    b->nsites = h_cnt[kCntSites];
    b->nsites_x = h_cnt[kCntSitesX];
    if (b->nsites > site_cap) throw Error("internal: site list overflow");
    if (b->use_site && b->nsites > 0) {
        b->a_need_lo = hg.nb + 1 - h_cnt[kCntALo];
        b->a_need_hi = h_cnt[kCntAHi];
    } else {
        b->a_need_lo = 1;
        b->a_need_hi = hg.nb;
    }
    b->nnzH = h_tot[0] - 1;
    b->nnzS = h_tot[1] - 1;
    return guard.release();
}

}  // namespace bs2e
