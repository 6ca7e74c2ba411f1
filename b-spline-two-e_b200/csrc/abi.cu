// abi.cu -- the extern "C" surface declared in include/bs2e.h.
#include <sched.h>

#include <cctype>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstring>
#include <memory>
#include <tuple>
#include <map>
#include <mutex>

#include "../../include/bs2e.h"
#include "ctx.h"
#include "files.h"
#include "geom_host.h"
#include "host.h"
#include "wigner.h"

namespace bs2e {

std::atomic<long long> g_launches{0};
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }

template <class F>
static int guarded(const char* where, F&& f)
{
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_last_error(std::string(where) + ": " + e.what());
        return 1;
    } catch (...) {
        set_last_error(std::string(where) + ": unknown error");
        return 2;
    }
}

static void use_device(const bs2e_ctx* c) { BS2E_CUDA(cudaSetDevice(c->device)); }
void purge_parked(bs2e_ctx* c);

}  // namespace bs2e

using namespace bs2e;

// count_nnz / construct_block_tensor shaped calls.  The plan made by the count
// call is parked so that the fill call that follows (hamiltonian.f90:137-139:
// count, allocate, fill) does not repeat the count pass.
namespace {
struct ParkKey {
    bs2e_ctx* c; int64_t L, n, full; uint64_t h;
    bool operator<(const ParkKey& o) const
    {
        return std::tie(c, L, n, full, h) < std::tie(o.c, o.L, o.n, o.full, o.h);
    }
};
std::mutex g_park_mu;
std::map<ParkKey, bs2e_block*> g_parked;

}  // namespace

namespace bs2e {
void purge_parked(bs2e_ctx* c)
{
    std::vector<bs2e_block*> dead;
    {
        std::lock_guard<std::mutex> lk(g_park_mu);
        for (auto it = g_parked.begin(); it != g_parked.end();) {
            if (it->first.c == c) { dead.push_back(it->second); it = g_parked.erase(it); }
            else ++it;
        }
    }
    for (bs2e_block* b : dead) block_free(b);
}
}  // namespace bs2e


extern "C" {

const char* bs2e_last_error(void) { return t_last_error.c_str(); }

int bs2e_device_count(int64_t* count)
{
    return guarded("bs2e_device_count", [&] {
        int n = 0;
        BS2E_CUDA(cudaGetDeviceCount(&n));
        *count = n;
    });
}

int bs2e_ctx_create(int64_t k_spline, int64_t n_knots, const double* knots, int64_t max_k,
                    int64_t k_GL, const double* gl_x, const double* gl_w, int64_t device,
                    bs2e_ctx** out)
{
    return guarded("bs2e_ctx_create", [&] {
        if (!knots || !gl_x || !gl_w || !out) throw Error("null argument");
        int ndev = 0;
        BS2E_CUDA(cudaGetDeviceCount(&ndev));
        if (ndev <= 0) throw Error("no CUDA device: this library has no CPU fallback");
        if (device < 0 || device >= ndev) throw Error("device index out of range");
        std::unique_ptr<bs2e_ctx> c(new bs2e_ctx());
        c->device = (int)device;
        try {
            build_host_geom(c->host, (int)k_spline, (int)n_knots, knots, (int)max_k, (int)k_GL, gl_x, gl_w);
        } catch (const std::invalid_argument& e) {
            throw Error(e.what());
        }
        c->hg = c->host.g;
        c->slice_lo = 1;
        c->slice_hi = c->hg.nb;
        c->max_k = c->host.max_k;
        c->nnz_4d = c->host.nnz_4d;
        c->nnz_6d = c->host.nnz_6d;
        use_device(c.get());
        cudaDeviceProp prop;
        BS2E_CUDA(cudaGetDeviceProperties(&prop, c->device));
        if (prop.major < 10) throw Error("this library is built for sm_100a (B200) only");
        BS2E_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
        {   // keep freed output arrays in the device's default memory pool (see dev_alloc_async)
            cudaMemPool_t pool = nullptr;
            BS2E_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->device));
            unsigned long long keep = ~0ull;
            BS2E_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        BS2E_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        BS2E_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        BS2E_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        c->d_t = dev_upload(c->host.knots, c->stream);
        c->d_bp = dev_upload(c->host.bp, c->stream);
        c->d_glx = dev_upload(c->host.glx, c->stream);
        c->d_glw = dev_upload(c->host.glw, c->stream);
        c->d_rowoff = dev_upload(c->host.rowoff, c->stream);
        c->d_pair = dev_upload(c->host.pairs, c->stream);
        c->dg = c->hg;
        c->dg.t = c->d_t;
        c->dg.bp = c->d_bp;
        c->dg.glx = c->d_glx;
        c->dg.glw = c->d_glw;
        c->dg.rowoff = c->d_rowoff;
        c->dg.pair = c->d_pair;
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        *out = c.release();
    });
}

int bs2e_ctx_destroy(bs2e_ctx* c)
{
    return guarded("bs2e_ctx_destroy", [&] {
        if (!c) return;
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        purge_parked(c);          // plans parked by bs2e_block_count that no fill call collected
        out_release_all(c);       // cached CSR output arrays
        cudaStreamSynchronize(c->stream);
        stager_destroy(c);
        ctx_release_plan_state(c);
        cudaFree(c->d_t); cudaFree(c->d_bp); cudaFree(c->d_glx); cudaFree(c->d_glw);
        cudaFree(c->d_rowoff); cudaFree(c->d_pair);
        cudaFree(c->d_mom_rk); cudaFree(c->d_mom_rmk); cudaFree(c->d_pre); cudaFree(c->d_sufx);
        cudaFree(c->d_rd); cudaFree(c->d_rkrow); cudaFree(c->d_R); cudaFree(c->d_Hb); cudaFree(c->d_Sb);
        cudaFree(c->d_dipA); cudaFree(c->d_dipB);
        if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
        if (c->have_lanes) {
            for (auto& ln : c->lanes) {
                cudaStreamSynchronize(ln.main); cudaStreamSynchronize(ln.side);
                cudaStreamDestroy(ln.main); cudaStreamDestroy(ln.side);
                cudaEventDestroy(ln.fork); cudaEventDestroy(ln.join); cudaEventDestroy(ln.done);
            }
            cudaEventDestroy(c->ev_start);
        }
        if (c->ev_fork) cudaEventDestroy(c->ev_fork);
        if (c->ev_join) cudaEventDestroy(c->ev_join);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        {   // hand the cached output pages back
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        }
        delete c;
    });
}

int bs2e_ctx_set_stream(bs2e_ctx* c, void* cuda_stream)
{
    return guarded("bs2e_ctx_set_stream", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        if (c->own_stream) cudaStreamDestroy(c->stream);
        c->stream = (cudaStream_t)cuda_stream;
        c->own_stream = false;
    });
}

int bs2e_ctx_sync(bs2e_ctx* c)
{
    return guarded("bs2e_ctx_sync", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int bs2e_sizes(bs2e_ctx* c, int64_t* n_b, int64_t* cells, int64_t* P, int64_t* nnz_4d,
               int64_t* nnz_6d)
{
    return guarded("bs2e_sizes", [&] {
        if (!c) throw Error("null context");
        if (n_b) *n_b = c->hg.nb;
        if (cells) *cells = c->hg.cells;
        if (P) *P = c->hg.P;
        if (nnz_4d) *nnz_4d = c->nnz_4d;
        if (nnz_6d) *nnz_6d = c->nnz_6d;
    });
}

int bs2e_slater_cells(bs2e_ctx* c)
{
    return guarded("bs2e_slater_cells", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        run_slater_cells(c);
    });
}

int bs2e_get_r_k(bs2e_ctx* c, double* r_k, double* r_m_k, int64_t* iv, int64_t* i, int64_t* j)
{
    return guarded("bs2e_get_r_k", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_r_k(c, r_k, r_m_k, iv, i, j);
    });
}

int bs2e_get_r_d_k(bs2e_ctx* c, double* r_d_k, int64_t* iv, int64_t* i, int64_t* j, int64_t* i_p,
                   int64_t* j_p)
{
    return guarded("bs2e_get_r_d_k", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_r_d_k(c, r_d_k, iv, i, j, i_p, j_p);
    });
}

int bs2e_rk_build(bs2e_ctx* c)
{
    return guarded("bs2e_rk_build", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        run_rk_build(c);
    });
}

int bs2e_rk_rows(bs2e_ctx* c, int64_t a_lo, int64_t a_hi)
{
    return guarded("bs2e_rk_rows", [&] {
        if (!c) throw Error("null context");
        if (a_lo < 1 || a_hi > c->hg.nb || a_lo > a_hi) throw Error("bs2e_rk_rows: need 1 <= a_lo <= a_hi <= n_b");
        c->slice_lo = (int)a_lo;
        c->slice_hi = (int)a_hi;
    });
}

int bs2e_rk_get(bs2e_ctx* c, int64_t n_keys, const int64_t* keys, double* vals)
{
    return guarded("bs2e_rk_get", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_rk_keys(c, n_keys, keys, vals);
    });
}

int bs2e_rk_plane(bs2e_ctx* c, int64_t k, double* out)
{
    return guarded("bs2e_rk_plane", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_rk_plane(c, (int)k, out);
    });
}

int bs2e_set_one_particle(bs2e_ctx* c, int64_t max_l_1p, const double* H_vec, const double* S)
{
    return guarded("bs2e_set_one_particle", [&] {
        if (!c || !H_vec || !S) throw Error("null argument");
        if (max_l_1p < 0 || max_l_1p > 120) throw Error("max_l_1p out of range");
        use_device(c);
        const Geom& g = c->hg;
        const size_t per_l = band_doubles(g);
        std::vector<double> Hb((size_t)(max_l_1p + 1) * per_l, 0.0), Sb(per_l, 0.0);
        try {
            pack_band(g, S, Sb.data(), "S");
            for (int l = 0; l <= max_l_1p; ++l)
                pack_band(g, H_vec + (size_t)l * g.nb * g.nb * 2, Hb.data() + l * per_l, "H_vec");
        } catch (const std::invalid_argument& e) {
            throw Error(e.what());
        }
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_Hb);
        cudaFree(c->d_Sb);
        c->d_Hb = c->d_Sb = nullptr;
        c->d_Hb = dev_upload(Hb, c->stream);
        c->d_Sb = dev_upload(Sb, c->stream);
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        c->lmax_1p = (int)max_l_1p;
        c->have_1p = true;
    });
}

int bs2e_one_particle_device(bs2e_ctx* c, int64_t Z, int64_t max_l_1p, int64_t CAP_order, double CAP_r_0,
                             double CAP_eta_re, double CAP_eta_im)
{
    return guarded("bs2e_one_particle_device", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        one_particle_device(c, (int)Z, (int)max_l_1p, (int)CAP_order, CAP_r_0, CAP_eta_re, CAP_eta_im);
    });
}

int bs2e_get_one_particle(bs2e_ctx* c, double* H_vec, double* S)
{
    return guarded("bs2e_get_one_particle", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_one_particle(c, H_vec, S);
    });
}

int bs2e_radial_dipole_device(bs2e_ctx* c, int64_t gauge)
{
    return guarded("bs2e_radial_dipole_device", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        radial_dipole_device(c, (int)gauge);
    });
}

int bs2e_get_radial_dipole(bs2e_ctx* c, double* A, double* B)
{
    return guarded("bs2e_get_radial_dipole", [&] {
        if (!c) throw Error("null context");
        use_device(c);
        fetch_radial_dipole(c, A, B);
    });
}

int bs2e_set_radial_dipole(bs2e_ctx* c, int64_t gauge, const double* A, const double* B)
{
    return guarded("bs2e_set_radial_dipole", [&] {
        if (!c || !A) throw Error("null argument");
        if (gauge != 'l' && gauge != 'v') throw Error("gauge must be 'l' (108) or 'v' (118)");
        if (gauge == 'v' && !B) throw Error("the velocity gauge needs r_inv_mat");
        use_device(c);
        const Geom& g = c->hg;
        const size_t per = band_doubles(g);
        std::vector<double> Ab(per, 0.0), Bb(gauge == 'v' ? per : 0, 0.0);
        try {
            pack_band(g, A, Ab.data(), gauge == 'l' ? "r_mat" : "dr_mat");
            if (gauge == 'v') pack_band(g, B, Bb.data(), "r_inv_mat");
        } catch (const std::invalid_argument& e) {
            throw Error(e.what());
        }
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_dipA);
        cudaFree(c->d_dipB);
        c->d_dipA = c->d_dipB = nullptr;
        c->d_dipA = dev_upload(Ab, c->stream);
        if (gauge == 'v') c->d_dipB = dev_upload(Bb, c->stream);
        BS2E_CUDA(cudaStreamSynchronize(c->stream));
        c->dip_gauge = (int)gauge;
        c->have_dip = true;
    });
}

int bs2e_dip_block_count(bs2e_ctx* c, int64_t q, const int64_t* sym1, int64_t n_config1, const int64_t* conf_n1,
                         const int64_t* conf_l1, const int64_t* sym2, int64_t n_config2, const int64_t* conf_n2,
                         const int64_t* conf_l2, int64_t compute, int64_t* nnz)
{
    return guarded("bs2e_dip_block_count", [&] {
        if (!c || !sym1 || !sym2 || !conf_n1 || !conf_l1 || !conf_n2 || !conf_l2 || !nnz) throw Error("null argument");
        use_device(c);
        *nnz = dip_block_run(c, (int)q, sym1, n_config1, conf_n1, conf_l1, sym2, n_config2, conf_n2, conf_l2,
                             compute != 0, nullptr, nullptr, nullptr);
    });
}

int bs2e_dip_block_fill(bs2e_ctx* c, int64_t q, const int64_t* sym1, int64_t n_config1, const int64_t* conf_n1,
                        const int64_t* conf_l1, const int64_t* sym2, int64_t n_config2, const int64_t* conf_n2,
                        const int64_t* conf_l2, int64_t compute, int64_t* index_ptr, int64_t* indices, double* data)
{
    return guarded("bs2e_dip_block_fill", [&] {
        if (!c || !sym1 || !sym2 || !conf_n1 || !conf_l1 || !conf_n2 || !conf_l2 || !index_ptr) throw Error("null argument");
        use_device(c);
        dip_block_run(c, (int)q, sym1, n_config1, conf_n1, conf_l1, sym2, n_config2, conf_n2, conf_l2, compute != 0,
                      index_ptr, indices, data);
    });
}

int bs2e_block_plan(bs2e_ctx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                    const int64_t* conf_l, int64_t full, int64_t row_lo, int64_t row_hi,
                    bs2e_block** blk)
{
    return guarded("bs2e_block_plan", [&] {
        if (!c || !conf_n || !conf_l || !blk) throw Error("null argument");
        use_device(c);
        *blk = block_plan(c, (int)L, n_config, conf_n, conf_l, nullptr, full != 0, 1, &row_lo, &row_hi);
    });
}

int bs2e_block_plan_ranges(bs2e_ctx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                           const int64_t* conf_l, int64_t full, int64_t n_ranges,
                           const int64_t* range_lo, const int64_t* range_hi, bs2e_block** blk)
{
    return guarded("bs2e_block_plan_ranges", [&] {
        if (!c || !conf_n || !conf_l || !blk || !range_lo || !range_hi) throw Error("null argument");
        use_device(c);
        *blk = block_plan(c, (int)L, n_config, conf_n, conf_l, nullptr, full != 0, n_ranges, range_lo, range_hi);
    });
}

int bs2e_configs_upload(bs2e_ctx* c, int64_t n_config, const int64_t* conf_n, const int64_t* conf_l,
                        bs2e_configs** cfg)
{
    return guarded("bs2e_configs_upload", [&] {
        if (!c || !cfg || (n_config > 0 && (!conf_n || !conf_l))) throw Error("null argument");
        use_device(c);
        *cfg = configs_upload(c, n_config, conf_n, conf_l);
    });
}

int bs2e_configs_free(bs2e_configs* cfg)
{
    return guarded("bs2e_configs_free", [&] {
        if (!cfg) return;
        cudaSetDevice(cfg->ctx->device);
        cudaStreamSynchronize(cfg->ctx->stream);
        configs_free(cfg);
    });
}

int bs2e_block_plan_dev(bs2e_ctx* c, int64_t L, bs2e_configs* cfg, int64_t full, int64_t n_ranges,
                        const int64_t* range_lo, const int64_t* range_hi, bs2e_block** blk)
{
    return guarded("bs2e_block_plan_dev", [&] {
        if (!c || !cfg || !blk) throw Error("null argument");
        use_device(c);
        const int64_t one = 1, all = cfg->n;
        if (n_ranges <= 0 || !range_lo || !range_hi) { n_ranges = 1; range_lo = &one; range_hi = &all; }
        *blk = block_plan(c, (int)L, cfg->n, nullptr, nullptr, cfg, full != 0, n_ranges, range_lo, range_hi);
    });
}

int bs2e_block_nnz(bs2e_block* b, int64_t* nnz_H, int64_t* nnz_S)
{
    return guarded("bs2e_block_nnz", [&] {
        if (!b) throw Error("null block");
        if (nnz_H) *nnz_H = b->nnzH;
        if (nnz_S) *nnz_S = b->nnzS;
    });
}

int bs2e_block_row_counts(bs2e_block* b, int64_t* cnt_H, int64_t* cnt_S)
{
    return guarded("bs2e_block_row_counts", [&] {
        if (!b) throw Error("null block");
        use_device(b->ctx);
        block_row_counts(b, cnt_H, cnt_S);
    });
}

int bs2e_block_recount(bs2e_block* b)
{
    return guarded("bs2e_block_recount", [&] {
        if (!b) throw Error("null block");
        use_device(b->ctx);
        block_count_scan(b, false);
    });
}

int bs2e_block_assemble(bs2e_block* b)
{
    return guarded("bs2e_block_assemble", [&] {
        if (!b) throw Error("null block");
        use_device(b->ctx);
        block_assemble(b);
    });
}

int bs2e_blocks_run(bs2e_ctx* c, int64_t n, bs2e_block** blks, int64_t recount)
{
    return guarded("bs2e_blocks_run", [&] {
        if (!c || (n > 0 && !blks)) throw Error("null argument");
        use_device(c);
        blocks_run(c, n, blks, recount != 0);
    });
}

int bs2e_block_download(bs2e_block* b, int64_t* H_ptr, int64_t* H_idx, double* H_dat,
                        int64_t* S_ptr, int64_t* S_idx, double* S_dat)
{
    return guarded("bs2e_block_download", [&] {
        if (!b) throw Error("null block");
        use_device(b->ctx);
        block_download(b, H_ptr, H_idx, H_dat, S_ptr, S_idx, S_dat);
    });
}

int bs2e_block_checksum(bs2e_block* b, uint64_t* sum_H, uint64_t* sum_S)
{
    return guarded("bs2e_block_checksum", [&] {
        if (!b) throw Error("null block");
        use_device(b->ctx);
        block_checksum(b, sum_H, sum_S);
    });
}

int bs2e_block_free(bs2e_block* b)
{
    return guarded("bs2e_block_free", [&] {
        if (!b) return;
        cudaSetDevice(b->ctx->device);
        block_free(b);   // stream-ordered on the context's stream: no host synchronisation
    });
}

namespace {
uint64_t conf_hash(int64_t n, const int64_t* a, const int64_t* b)
{
    uint64_t h = 1469598103934665603ull;
    for (int64_t q = 0; q < 2 * n; ++q) {
        h = (h ^ (uint64_t)a[q]) * 1099511628211ull;
        h = (h ^ (uint64_t)b[q]) * 1099511628211ull;
    }
    return h;
}
}  // namespace

int bs2e_block_count(bs2e_ctx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                     const int64_t* conf_l, int64_t full, int64_t* nnz_H, int64_t* nnz_S)
{
    return guarded("bs2e_block_count", [&] {
        if (!c || (n_config > 0 && (!conf_n || !conf_l))) throw Error("null argument");
        use_device(c);
        const int64_t one = 1;
        const bool trace = getenv("BS2E_TRACE") != nullptr;
        const auto tc0 = std::chrono::steady_clock::now();
        bs2e_block* b = block_plan(c, (int)L, n_config, conf_n, conf_l, nullptr, full != 0, 1, &one, &n_config);
        if (trace)
            fprintf(stderr, "bs2e_block_count L=%lld: %.2f ms\n", (long long)L,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
        if (nnz_H) *nnz_H = b->nnzH;
        if (nnz_S) *nnz_S = b->nnzS;
        const ParkKey key{c, L, n_config, full != 0, conf_hash(n_config, conf_n, conf_l)};
        bs2e_block* old = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_park_mu);
            auto it = g_parked.find(key);
            if (it != g_parked.end()) { old = it->second; it->second = b; }
            else g_parked.emplace(key, b);
        }
        if (old) block_free(old);
    });
}

int bs2e_block_fill(bs2e_ctx* c, int64_t L, int64_t n_config, const int64_t* conf_n,
                    const int64_t* conf_l, int64_t full, int64_t* H_ptr, int64_t* H_idx,
                    double* H_dat, int64_t* S_ptr, int64_t* S_idx, double* S_dat)
{
    return guarded("bs2e_block_fill", [&] {
        if (!c || (n_config > 0 && (!conf_n || !conf_l))) throw Error("null argument");
        use_device(c);
        const ParkKey key{c, L, n_config, full != 0, conf_hash(n_config, conf_n, conf_l)};
        bs2e_block* b = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_park_mu);
            auto it = g_parked.find(key);
            if (it != g_parked.end()) { b = it->second; g_parked.erase(it); }
        }
        const int64_t one = 1;
        const bool trace = getenv("BS2E_TRACE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        if (!b) b = block_plan(c, (int)L, n_config, conf_n, conf_l, nullptr, full != 0, 1, &one, &n_config);
        double t1 = 0, t2 = 0, t3 = 0;
        try {
            t1 = now();
            block_assemble(b);
            if (trace) { cudaStreamSynchronize(c->stream); t2 = now(); }
            block_download(b, H_ptr, H_idx, H_dat, S_ptr, S_idx, S_dat);
            t3 = now();
        } catch (...) {
            block_free(b);
            throw;
        }
        block_free(b);
        if (trace)
            fprintf(stderr, "bs2e_block_fill L=%lld: plan %.2f ms, alloc+fill %.2f ms, download %.2f ms, free %.2f ms\n",
                    (long long)L, t1 - t0, t2 - t1, t3 - t2, now() - t3);
    });
}

// Pinned pages are placed by the kernel's local-allocation policy on the NUMA
// node of the calling thread.  A buffer on the socket the GPU is not attached
// to halves the D2H rate (every byte crosses the inter-socket link), so the
// calling thread is moved onto the CPUs of the GPU's node for the duration of
// the allocation.  Best effort: any failure leaves the default placement.
namespace {
bool gpu_node_cpus(cpu_set_t* set)
{
    int dev = 0;
    char bus[32] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return false;
    for (char* q = bus; *q; ++q) *q = (char)tolower(*q);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a = 0, b = 0, n = 0;
    for (;;) {  // "0-15,32-47"
        if (fscanf(f, "%d", &a) != 1) break;
        b = a;
        int ch = fgetc(f);
        if (ch == '-') {
            if (fscanf(f, "%d", &b) != 1) break;
            ch = fgetc(f);
        }
        for (int cpu = a; cpu <= b && cpu < CPU_SETSIZE; ++cpu) { CPU_SET(cpu, set); ++n; }
        if (ch != ',') break;
    }
    fclose(f);
    return n > 0;
}
}  // namespace

int bs2e_host_alloc(int64_t bytes, void** ptr)
{
    return guarded("bs2e_host_alloc", [&] {
        if (!ptr || bytes < 0) throw Error("bad argument");
        cpu_set_t saved, want;
        const bool have_saved = sched_getaffinity(0, sizeof(saved), &saved) == 0;
        bool moved = false;
        if (have_saved && gpu_node_cpus(&want)) {
            cpu_set_t both;
            CPU_AND(&both, &want, &saved);  // stay inside the CPUs this process may use
            if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof(both), &both) == 0;
        }
        const cudaError_t e = cudaHostAlloc(ptr, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
        if (moved) sched_setaffinity(0, sizeof(saved), &saved);
        BS2E_CUDA(e);
    });
}

int bs2e_host_free(void* ptr)
{
    return guarded("bs2e_host_free", [&] {
        if (ptr) BS2E_CUDA(cudaFreeHost(ptr));
    });
}

int64_t bs2e_launch_count(void) { return g_launches.load(); }

int bs2e_debug_site_phase_cycles(uint64_t* out8, int64_t reset)
{
    return guarded("bs2e_debug_site_phase_cycles", [&] {
        if (!out8) throw Error("null argument");
        site_phase_cycles(reinterpret_cast<unsigned long long*>(out8), reset != 0);
    });
}

// ---- result files (host only) ----------------------------------------------
struct bs2e_file {
    std::unique_ptr<files::Writer> w;
    std::unique_ptr<files::Reader> r;
    std::vector<char> rec;
};

int bs2e_file_create_block_diag(const char* path, int64_t n_blocks, const int64_t* block_rows, bs2e_file** f)
{
    return guarded("bs2e_file_create_block_diag", [&] {
        if (!path || !f || n_blocks < 0 || (n_blocks > 0 && !block_rows)) throw Error("bad argument");
        std::unique_ptr<bs2e_file> h(new bs2e_file());
        h->w.reset(new files::Writer(path));
        h->w->block_diag_header(n_blocks, block_rows);
        *f = h.release();
    });
}

int bs2e_file_create_block_matrix(const char* path, int64_t n_block_rows, int64_t n_block_cols,
                                  const int64_t* block_rows, const int64_t* block_cols, bs2e_file** f)
{
    return guarded("bs2e_file_create_block_matrix", [&] {
        if (!path || !f || n_block_rows < 1 || n_block_cols < 1 || !block_rows || !block_cols) throw Error("bad argument");
        std::unique_ptr<bs2e_file> h(new bs2e_file());
        h->w.reset(new files::Writer(path));
        h->w->block_matrix_header(n_block_rows, n_block_cols, block_rows, block_cols);
        *f = h.release();
    });
}

int bs2e_file_write_block(bs2e_file* f, int64_t rows, int64_t cols, int64_t nnz, const int64_t* index_ptr,
                          const int64_t* indices, const double* data)
{
    return guarded("bs2e_file_write_block", [&] {
        if (!f || !f->w) throw Error("file not open for writing");
        if (nnz > 0 && (!index_ptr || !indices || !data)) throw Error("null array");
        f->w->csr_block(rows, cols, nnz, index_ptr, indices, data);
    });
}

int bs2e_file_write_block_fragments(bs2e_file* f, int64_t rows, int64_t cols, int64_t n_frag,
                                    const int64_t* frag_rows, const int64_t* const* frag_ptr,
                                    const int64_t* const* frag_idx, const double* const* frag_dat)
{
    return guarded("bs2e_file_write_block_fragments", [&] {
        if (!f || !f->w) throw Error("file not open for writing");
        if (n_frag < 1 || !frag_rows || !frag_ptr || !frag_idx || !frag_dat) throw Error("bad argument");
        std::vector<long long> fr(frag_rows, frag_rows + n_frag);
        f->w->csr_block_fragments(rows, cols, (int)n_frag, fr.data(), frag_ptr, frag_idx, frag_dat);
    });
}

int bs2e_file_close(bs2e_file* f)
{
    return guarded("bs2e_file_close", [&] {
        if (!f) return;
        std::unique_ptr<bs2e_file> h(f);
        if (h->w) h->w->close();
    });
}

int bs2e_file_write_basis(const char* path, int64_t max_l_1p, int64_t max_L, int64_t two_el, int64_t n_sym,
                          const int64_t* sym_l, const int64_t* sym_m, const int64_t* sym_pi,
                          const int64_t* n_config, const int64_t* const* conf_n, const int64_t* const* conf_l,
                          const int64_t* const* conf_eqv)
{
    return guarded("bs2e_file_write_basis", [&] {
        if (!path || n_sym < 0) throw Error("bad argument");
        files::write_basis(path, max_l_1p, max_L, two_el != 0, n_sym, sym_l, sym_m, sym_pi, n_config, conf_n,
                           conf_l, conf_eqv);
    });
}

int bs2e_file_write_splines(const char* path, int64_t k, int64_t n_knots, const double* knots)
{
    return guarded("bs2e_file_write_splines", [&] {
        if (!path || !knots) throw Error("bad argument");
        files::write_splines(path, k, n_knots, knots);
    });
}

int bs2e_file_open(const char* path, bs2e_file** f)
{
    return guarded("bs2e_file_open", [&] {
        if (!path || !f) throw Error("bad argument");
        std::unique_ptr<bs2e_file> h(new bs2e_file());
        h->r.reset(new files::Reader(path));
        *f = h.release();
    });
}

int bs2e_file_next_record(bs2e_file* f, int64_t* nbytes)
{
    return guarded("bs2e_file_next_record", [&] {
        if (!f || !f->r || !nbytes) throw Error("file not open for reading");
        f->r->next_record(f->rec);
        *nbytes = (int64_t)f->rec.size();
    });
}

int bs2e_file_record_data(bs2e_file* f, void* dst, int64_t nbytes)
{
    return guarded("bs2e_file_record_data", [&] {
        if (!f || !f->r || (nbytes > 0 && !dst)) throw Error("bad argument");
        if ((size_t)nbytes != f->rec.size()) throw Error("size does not match the current record");
        if (nbytes) std::memcpy(dst, f->rec.data(), (size_t)nbytes);
    });
}

int bs2e_file_set_max_subrecord(int64_t bytes)
{
    files::set_max_subrecord(bytes);
    return 0;
}

// ---- host companions -------------------------------------------------------
int64_t bs2e_host_generate_grid(int64_t k, int64_t m, int64_t Z, double h_max, double r_max,
                                double* grid, int64_t cap)
{
    int64_t n = -1;
    guarded("bs2e_host_generate_grid", [&] {
        auto g = host::generate_grid((int)k, (int)m, (int)Z, h_max, r_max);
        n = (int64_t)g.size();
        if (grid && n <= cap) std::memcpy(grid, g.data(), sizeof(double) * g.size());
    });
    return n;
}

int bs2e_host_gauss_legendre(int64_t N, double a, double b, double* x, double* w)
{
    return guarded("bs2e_host_gauss_legendre", [&] { host::gauss_legendre((int)N, a, b, x, w); });
}

int64_t bs2e_host_find_max_n_b(int64_t k, int64_t n_knots, const double* knots, double x)
{
    std::vector<double> t(knots, knots + n_knots);
    return host::find_max_n_b((int)k, t, x);
}

int bs2e_host_setup_S(int64_t k, int64_t n_knots, const double* knots, int64_t k_GL, double* S)
{
    return guarded("bs2e_host_setup_S", [&] {
        std::vector<double> t(knots, knots + n_knots);
        host::setup_S((int)k, t, (int)k_GL, reinterpret_cast<std::complex<double>*>(S));
    });
}

int bs2e_host_setup_H_one_particle(int64_t k, int64_t n_knots, const double* knots, int64_t Z,
                                   int64_t l, int64_t CAP_order, double CAP_r_0, double CAP_eta_re,
                                   double CAP_eta_im, int64_t k_GL, double* H)
{
    return guarded("bs2e_host_setup_H_one_particle", [&] {
        std::vector<double> t(knots, knots + n_knots);
        host::setup_H_one_particle((int)k, t, (int)Z, (int)l, (int)CAP_order, CAP_r_0,
                                   std::complex<double>(CAP_eta_re, CAP_eta_im), (int)k_GL,
                                   reinterpret_cast<std::complex<double>*>(H));
    });
}

int bs2e_host_setup_radial_dip(int64_t k, int64_t n_knots, const double* knots, int64_t k_GL, int64_t gauge,
                               double* A, double* B)
{
    return guarded("bs2e_host_setup_radial_dip", [&] {
        if (!knots || !A || (gauge == 'v' && !B)) throw Error("null argument");
        std::vector<double> t(knots, knots + n_knots);
        try {
            host::setup_radial_dip((int)k, t, (int)k_GL, (int)gauge, reinterpret_cast<std::complex<double>*>(A),
                                   reinterpret_cast<std::complex<double>*>(B));
        } catch (const std::invalid_argument& e) {
            throw Error(e.what());
        }
    });
}

int64_t bs2e_host_basis_syms(int64_t max_L, int64_t z_pol, int64_t* sym_l, int64_t* sym_m,
                             int64_t* sym_pi, int64_t cap)
{
    auto s = host::basis_syms((int)max_L, z_pol != 0);
    if ((int64_t)s.size() <= cap)
        for (size_t q = 0; q < s.size(); ++q) {
            sym_l[q] = s[q].l;
            sym_m[q] = s[q].m;
            sym_pi[q] = s[q].pi ? 1 : 0;
        }
    return (int64_t)s.size();
}

int64_t bs2e_host_count_configs(int64_t term_l, int64_t term_pi, int64_t max_l_1p, int64_t n_b,
                                int64_t k_spline, int64_t max_n_b, int64_t n_all_l,
                                int64_t l_2_max, int64_t* conf_n, int64_t* conf_l,
                                int64_t* conf_eqv, int64_t cap)
{
    std::vector<int64_t> cn, cl, ce;
    host::count_configs((int)term_l, term_pi != 0, (int)max_l_1p, (int)n_b, (int)k_spline,
                        (int)max_n_b, (int)n_all_l, (int)l_2_max, cn, cl, ce);
    const int64_t n = (int64_t)ce.size();
    if (n <= cap && n > 0) {
        if (conf_n) std::memcpy(conf_n, cn.data(), sizeof(int64_t) * cn.size());
        if (conf_l) std::memcpy(conf_l, cl.data(), sizeof(int64_t) * cl.size());
        if (conf_eqv) std::memcpy(conf_eqv, ce.data(), sizeof(int64_t) * ce.size());
    }
    return n;
}

double bs2e_host_three_j0(int64_t ja, int64_t jb, int64_t jc) { return three_j0((int)ja, (int)jb, (int)jc); }
double bs2e_host_six_j(int64_t ja, int64_t jb, int64_t jc, int64_t jd, int64_t je, int64_t jf)
{
    return six_j((int)ja, (int)jb, (int)jc, (int)jd, (int)je, (int)jf);
}
double bs2e_host_ang_k_LS(int64_t k, int64_t la, int64_t lb, int64_t lc, int64_t ld, int64_t L)
{
    return ang_k_LS((int)k, (int)la, (int)lb, (int)lc, (int)ld, (int)L);
}

}  // extern "C"
