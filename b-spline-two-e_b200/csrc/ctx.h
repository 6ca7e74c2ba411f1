// ctx.h -- internal state behind the opaque handles of include/bs2e.h
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "core.h"
#include "geom_host.h"
#include "plan.h"

struct bs2e_ctx;

namespace bs2e {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

void set_last_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define BS2E_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess)                                                     \
            throw ::bs2e::Error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#define BS2E_LAUNCHED()                                  \
    do {                                                 \
        ::bs2e::g_launches.fetch_add(1);                 \
        BS2E_CUDA(cudaGetLastError());                   \
    } while (0)

template <class T>
T* dev_alloc(size_t n)
{
    T* p = nullptr;
    BS2E_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
    return p;
}
// Stream-ordered allocation for the large CSR output arrays: cudaMalloc / cudaFree of
// multi-GB buffers cost 5-500 ms each (measured; cudaFree unmaps synchronously), the
// driver's memory pool with an unlimited release threshold (set in bs2e_ctx_create)
// keeps the pages and hands them to the next block.
template <class T>
T* dev_alloc_async(size_t n, cudaStream_t st)
{
    T* p = nullptr;
    BS2E_CUDA(cudaMallocAsync(&p, sizeof(T) * (n ? n : 1), st));
    return p;
}

template <class T>
T* dev_upload(const std::vector<T>& v, cudaStream_t st)
{
    T* p = dev_alloc<T>(v.size());
    if (!v.empty())
        BS2E_CUDA(cudaMemcpyAsync(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return p;
}

// Plan tables of a block: one buffer carved into aligned pieces.  The buffers come from a small per-context
// pool (arena_take / arena_give, plan_dev.cu) that is recycled by hand rather than freed stream-ordered: a plan is
// built on the context's PLAN stream while the fill of the previous block still runs on the main stream, and a
// buffer released behind that fill (event on the main stream) must not tie the two streams together.  The pool's
// buffers themselves are allocated (once, power-of-two size classes) from the device's stream-ordered memory pool.
struct ArenaBuf {
    char* base = nullptr;
    size_t size = 0;
    cudaEvent_t free_after = nullptr;   // recorded on the stream of the last reader when the buffer is given back
    bool busy = false;
};
struct DevArena {
    ArenaBuf* buf = nullptr;
    size_t used = 0;
    template <class T>
    T* take(size_t n)
    {
        const size_t at = (used + 255) & ~(size_t)255;
        const size_t bytes = sizeof(T) * (n ? n : 1);
        if (!buf || at + bytes > buf->size) throw Error("internal: plan arena overflow");
        used = at + bytes;
        return reinterpret_cast<T*>(buf->base + at);
    }
    static size_t need(size_t bytes) { return ((bytes ? bytes : 1) + 255) & ~(size_t)255; }
};

// angular tables of one (L, list of (l1,l2) groups, max_k): computed once per context with
// exact arithmetic on the host and kept on the device ("tabulated on the host, uploaded once")
struct AngDev {
    AngTables host;
    unsigned char* flags = nullptr;
    KRange* krange = nullptr;
    double *angD = nullptr, *angX = nullptr, *angP = nullptr;
    ~AngDev()
    {
        cudaFree(flags); cudaFree(krange); cudaFree(angD); cudaFree(angX); cudaFree(angP);
    }
};

// group structure of a configuration list on the device (plan_dev.cu: build_structure)
struct GroupStructure {
    bs2e_ctx* ctx = nullptr;
    int nblk = 0, max_nd = 0, lmax = 0;
    std::vector<BlockDesc> blocks;   // (l1, l2) of every group (host copy: key of the angular tables)
    char* d_mem = nullptr;           // one allocation behind the three tables
    BlockDesc* d_blk = nullptr;      // [nblk] with the n(1) ranges filled in
    int* d_blk_start = nullptr;      // [nblk+1] first configuration (0-based) of each group
    NcRow* d_ncrow = nullptr;        // [nblk][nb+1]
    ~GroupStructure();
};

}  // namespace bs2e

struct bs2e_ctx;
// a configuration list resident on the device (bs2e_configs_upload)
struct bs2e_configs {
    bs2e_ctx* ctx = nullptr;
    long long n = 0;
    long long *d_n = nullptr, *d_l = nullptr;  // (2, n) each, as given by the caller
    mutable std::shared_ptr<bs2e::GroupStructure> gs;   // built by the first plan of the list
};

struct bs2e_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // side stream + events: the two site-kernel launches of a block run concurrently
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // extra lanes for bs2e_blocks_run: consecutive blocks are issued on different
    // (main, side) stream pairs so that count passes and the last waves of the fill
    // launches of one block overlap the work of the next
    struct Lane { cudaStream_t main = nullptr, side = nullptr; cudaEvent_t fork = nullptr, join = nullptr, done = nullptr; };
    static constexpr int kLanes = 3;
    Lane lanes[kLanes];
    cudaEvent_t ev_start = nullptr;
    bool have_lanes = false;
    int max_k = 0;

    // plan state shared by the calls on this context (serialised by plan_mu)
    std::mutex plan_mu;
    std::map<std::vector<int>, std::shared_ptr<bs2e::AngDev>> ang_cache;
    cudaStream_t plan_stream = nullptr;   // plans are built here, concurrently with fills on `stream`
    std::mutex arena_mu;
    std::vector<bs2e::ArenaBuf*> arenas;  // pool of plan buffers (arena_take / arena_give)
    cudaMemPool_t arena_pool = nullptr;   // the memory pool they are allocated from
    // CSR output arrays handed back by bs2e_block_free, kept for the next block (out_take / out_give, block.cu)
    struct OutBuf { void* p; size_t bytes; cudaEvent_t ev; };
    std::vector<OutBuf> out_free;
    std::mutex out_mu;
    size_t out_cached = 0, out_cap = 0;   // bytes held in out_free / upper bound for them
    int* h_pin = nullptr;        // pinned staging for the small read-backs of a plan
    size_t h_pin_bytes = 0;
    // copy stream + pinned bounce buffers for downloads into pageable memory (download.cu)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy = nullptr;
    struct Stager;
    Stager* stager = nullptr;

    // host copy of the basis geometry
    bs2e::HostGeom host;
    bs2e::Geom hg{};  // = host.g, pointers are HOST pointers
    long long nnz_4d = 0, nnz_6d = 0;

    // device copy
    bs2e::Geom dg{};  // pointers in dg are DEVICE pointers
    double *d_t = nullptr, *d_bp = nullptr, *d_glx = nullptr, *d_glw = nullptr;
    int* d_rowoff = nullptr;
    bs2e::PairAC* d_pair = nullptr;

    // stage A products
    double *d_mom_rk = nullptr, *d_mom_rmk = nullptr;  // [K1][P][ks]
    double *d_pre = nullptr, *d_sufx = nullptr;        // [K1][P][ks+1]
    double* d_rd = nullptr;                            // [C][K1][ks^2][ks^2]
    bs2e::RkRow* d_rkrow = nullptr;                    // [K1][P] packed pair records of stage B
    bool have_cells = false;

    // stage B product
    double* d_R = nullptr;  // [K1][P][ldP]
    bool have_R = false;
    // row slice of stages A and B (bs2e_rk_rows): first spline index a of the band pairs (a, c) that are built;
    // slice_* = what the next run builds, cells_* / R_* = what the last run of stage A / B built
    int slice_lo = 1, slice_hi = 0;   // set to 1..n_b when the context is created
    int cells_lo = 1, cells_hi = 0, R_lo = 1, R_hi = 0;

    // stage C inputs (band storage)
    int lmax_1p = -1;
    double *d_Hb = nullptr, *d_Sb = nullptr;
    bool have_1p = false;
    // radial dipole matrices (band storage): gauge 'l': A = r_mat; gauge 'v': A = dr_mat, B = r_inv_mat
    int dip_gauge = 0;
    double *d_dipA = nullptr, *d_dipB = nullptr;
    bool have_dip = false;

    bs2e::CellData cell_data() const
    {
        return bs2e::CellData{d_mom_rk, d_mom_rmk, d_pre, d_sufx, d_rd};
    }
    bs2e::OneBody one_body() const { return bs2e::OneBody{d_Hb, d_Sb}; }
};

struct bs2e_block {
    bs2e_ctx* ctx = nullptr;
    int L = 0, full = 0, lmax = 0;
    long long n_config = 0, nrows = 0;   // nrows: planned rows (union of row ranges)
    bs2e::Plan dplan{};                  // device pointers
    // device-side plan storage: two stream-ordered arenas (before / after the structure read-back)
    bs2e::DevArena arena0, arena1;
    std::shared_ptr<bs2e::AngDev> ang;   // shared through the context's cache
    std::shared_ptr<bs2e::GroupStructure> gs;   // group tables (shared with the resident configuration list, if any)
    const long long *d_conf_n = nullptr, *d_conf_l = nullptr;   // configuration list on the device (owned by arena0 or by a bs2e_configs)
    int* d_blk_start = nullptr;          // [nblk+1] first configuration (0-based) of each (l1,l2) group
    // radial sites of the planned rows, sorted (plan.h: site_sort_key); counters[0] = nsites, [1] = nsites_x
    unsigned long long* d_site_key = nullptr;
    int* d_counters = nullptr;
    int site_cap = 0;
    int nsites = 0, nsites_x = 0;
    int a_need_lo = 1, a_need_hi = 0;   // first spline indices of the R^k rows the planned sites read
    bool use_site = false;               // site kernels (else: row-wise fallback with per-row tables)
    bool use_mma = false;                // tensor-core site kernel (site_mma.cu); else the FMA site kernel of block.cu
    // CSR fragment (device)
    long long *d_cntH = nullptr, *d_cntS = nullptr;  // [nrows+1] counts
    long long *d_Hptr = nullptr, *d_Sptr = nullptr;  // [nrows+1] 1-based
    long long *d_Hidx = nullptr, *d_Sidx = nullptr;
    double *d_Hdat = nullptr, *d_Sdat = nullptr;
    size_t cap_Hidx = 0, cap_Sidx = 0, cap_Hdat = 0, cap_Sdat = 0;   // capacities of the arrays (out_take)
    long long nnzH = 0, nnzS = 0;
    void* d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    bool assembled = false;
};

namespace bs2e {
// stage launchers (slater.cu, rk.cu, block.cu)
void run_slater_cells(bs2e_ctx* c);
void run_rk_build(bs2e_ctx* c);
void fetch_r_k(bs2e_ctx* c, double* r_k, double* r_m_k, int64_t* iv, int64_t* i, int64_t* j);
void fetch_r_d_k(bs2e_ctx* c, double* r_d_k, int64_t* iv, int64_t* i, int64_t* j, int64_t* ip,
                 int64_t* jp);
void fetch_rk_keys(bs2e_ctx* c, long long n_keys, const int64_t* keys, double* vals);
void fetch_rk_plane(bs2e_ctx* c, int k, double* out);

// conf_n / conf_l: HOST arrays (uploaded) when cfg == nullptr, else the device-resident list cfg
bs2e_block* block_plan(bs2e_ctx* c, int L, long long n_config, const int64_t* conf_n,
                       const int64_t* conf_l, const bs2e_configs* cfg, int full, long long n_ranges,
                       const int64_t* range_lo, const int64_t* range_hi);
bs2e_configs* configs_upload(bs2e_ctx* c, long long n_config, const int64_t* conf_n, const int64_t* conf_l);
void configs_free(bs2e_configs* cfg);
void ctx_release_plan_state(bs2e_ctx* c);
// CSR output arrays: a per-context cache in front of the stream-ordered allocator (block.cu)
void* out_take(bs2e_ctx* c, size_t bytes, cudaStream_t st, size_t* cap);
void out_give(bs2e_ctx* c, void* p, size_t cap, cudaStream_t st);
void out_release_all(bs2e_ctx* c);
void arena_take(bs2e_ctx* c, DevArena& a, size_t bytes, cudaStream_t first_use);   // a free pool buffer of at least `bytes`
void arena_give(bs2e_ctx* c, DevArena& a, cudaStream_t last_use);  // back to the pool, reusable after the work queued on last_use
// per-row tables (row_n1 / row_n2 / row_blk) of a configuration list, built on the device
void build_row_tables(cudaStream_t st, long long n_config, const long long* d_conf_n, int nblk,
                      const int* d_blk_start, unsigned short* row_n1, unsigned short* row_n2,
                      unsigned short* row_blk);
// onebody.cu
void one_particle_device(bs2e_ctx* c, int Z, int lmax, int cap_order, double cap_r0, double eta_re, double eta_im);
void radial_dipole_device(bs2e_ctx* c, int gauge);
void fetch_one_particle(bs2e_ctx* c, double* H_vec, double* S);
void fetch_radial_dipole(bs2e_ctx* c, double* A, double* B);
void download_to_host(bs2e_ctx* c, void* dst, const void* d_src, size_t bytes);   // pinned or pageable destination
void download_flush(bs2e_ctx* c);
void stager_destroy(bs2e_ctx* c);
void site_phase_cycles(unsigned long long* out8, bool reset);
// the streams a block's work is issued on (default: the context's own pair)
struct BlockStreams { cudaStream_t main, side; cudaEvent_t fork, join; };
BlockStreams default_streams(bs2e_ctx* c);
void block_count_scan(bs2e_block* b, bool read_totals, const BlockStreams* bs = nullptr);
void block_assemble(bs2e_block* b, const BlockStreams* bs = nullptr);
void blocks_run(bs2e_ctx* c, long long n, bs2e_block** blks, bool recount);
void block_download(bs2e_block* b, int64_t* H_ptr, int64_t* H_idx, double* H_dat, int64_t* S_ptr,
                    int64_t* S_idx, double* S_dat);
void block_row_counts(bs2e_block* b, int64_t* cH, int64_t* cS);
void block_checksum(bs2e_block* b, uint64_t* sH, uint64_t* sS);
void block_free(bs2e_block* b);
long long dip_block_run(bs2e_ctx* c, int q, const int64_t* sym1, long long n1, const int64_t* conf_n1,
                        const int64_t* conf_l1, const int64_t* sym2, long long n2, const int64_t* conf_n2,
                        const int64_t* conf_l2, bool compute, int64_t* index_ptr, int64_t* indices, double* data);
}  // namespace bs2e
