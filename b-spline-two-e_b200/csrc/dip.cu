// dip.cu -- dipole blocks <sym1| d_q |sym2> as CSR matrices on the device.
//
// Stands in for construct_dip_block_tensor + init_dip_block
// (src/mat_els/dipole.f90:8-47,87-146).  The reference scans all n1 x n2 configuration
// pairs twice; here the stored columns of a row are generated from the (l1,l2) group
// structure of the column list (dip_core.h), a count kernel + exclusive scan give
// index_ptr, and one warp per row writes indices and values in ascending column order.
// HBM-write bound (24 B per stored element), no R^k involved.
#include <cub/device/device_scan.cuh>

#include "ctx.h"
#include "dip_plan.h"

namespace bs2e {

__global__ void dip_count_kernel(Geom g, Plan plC, DipTables dt, long long* __restrict__ cnt)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > dt.nrows) return;
    cnt[idx] = idx < dt.nrows ? dip_row_count(g, plC, dt, (int)idx + 1) : 0;  // slot nrows: total after the scan
}

constexpr int kDipWarps = 8;

namespace {
struct DipArena {   // device arrays of one call, released on every exit path
    std::vector<void*> p;
    ~DipArena() { for (void* q : p) cudaFree(q); }
    template <class T> T* up(const std::vector<T>& v, cudaStream_t s) { T* d = dev_upload(v, s); p.push_back(d); return d; }
    template <class T> T* alloc(size_t n) { T* d = dev_alloc<T>(n); p.push_back(d); return d; }
};
}  // namespace

// One warp per row.  Per flagged column group the lanes first take one n_c slot each (clipped n_d intervals of
// both windows, their union, its size), a warp scan turns the sizes into offsets, and then the lanes take
// consecutive stored ENTRIES of the (row, column group) pair: every store instruction writes 32 consecutive
// entries of the CSR row, whatever the width of the windows (<= 2w+1 columns per n_c; the first version walked
// the n_c slots one after the other with one lane per n_d: 15 of 32 lanes at best and the whole interval
// arithmetic repeated by every lane for every slot -- 42 warp instructions per stored entry, 0.08 of the HBM
// rate, profiles/r01t_*).
__global__ void __launch_bounds__(kDipWarps * 32, 4)
dip_fill_kernel(const __grid_constant__ Geom g, const __grid_constant__ Plan plC, const __grid_constant__ DipTables dt,
                const __grid_constant__ DipBand bd, const long long* __restrict__ ptr, long long* __restrict__ idx,
                double2* __restrict__ dat)
{
    extern __shared__ double2 dtab[];   // per warp: the band rows n_a and n_b of A, B, S ([6][2w+2], last slot = 0)
    const long long wrow = (long long)blockIdx.x * kDipWarps + (threadIdx.x >> 5);
    if (wrow >= dt.nrows) return;
    const int lane = threadIdx.x & 31;
    constexpr unsigned kAll = 0xffffffffu;
    RowInfo r = dip_row(dt, (int)wrow + 1);
    const int bi = r.bi;
    r.bi = -1;   // never the "own" group of the column list: no triangle cut (dip_for_each_chunk)
    long long pos = ptr[wrow] - 1;
    // Every band-matrix value an entry of this row needs has n_a or n_b as its first index (dip_value_cf: A(n_a,.),
    // A(n_b,.), S(n_a,.), S(n_b,.), the same for B): the six band rows are copied to shared memory once per row,
    // and the twelve look-ups per entry become shared-memory reads (slot 2w+1 holds the zero outside the band).
    const int W = g.w, W2 = 2 * W + 2;
    double2* tb = dtab + (size_t)(threadIdx.x >> 5) * 6 * W2;
    for (int i = lane; i < 6 * W2; i += 32) {
        const int which = i / W2, d = i - which * W2;
        const int n = which < 3 ? r.na : r.nb;
        const double* M = (which % 3 == 0) ? bd.A : ((which % 3 == 1) ? bd.B : bd.S);
        double2 v = make_double2(0.0, 0.0);
        if (d <= 2 * W) {
            const double* q = M + ((size_t)n * (2 * W + 1) + d) * 2;
            v = make_double2(q[0], q[1]);
        }
        tb[i] = v;
    }
    __syncwarp();
    auto look = [&](int which, int n, int np) {
        const int d = np - n + W;
        const double2 v = tb[which * W2 + ((unsigned)d <= (unsigned)(2 * W) ? d : 2 * W + 1)];
        return Cplx{v.x, v.y};
    };
    // the statements of dip_value_cf with the look-ups above
    auto value = [&](const double* cf, int nc, int nd) {
        Cplx acc = Cplx{0.0, 0.0};
        const int sel[4] = {0, 1, 0, 1};                       // first index of the dipole factor: n_a / n_b
        const int n_[4] = {r.na, r.nb, r.na, r.nb}, np_[4] = {nc, nd, nd, nc};
        const int m_[4] = {r.nb, r.na, r.nb, r.na}, mp_[4] = {nd, nc, nc, nd};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double al = cf[2 * t], be = cf[2 * t + 1];
            if (al == 0.0 && be == 0.0) continue;
            const Cplx a = look(3 * sel[t], n_[t], np_[t]);
            Cplx d = Cplx{al * a.re, al * a.im};
            if (be != 0.0) {
                const Cplx b = look(3 * sel[t] + 1, n_[t], np_[t]);
                d = Cplx{al * a.re + be * b.re, al * a.im + be * b.im};
            }
            acc = cadd(acc, cmul(d, look(3 * (1 - sel[t]) + 2, m_[t], mp_[t])));
        }
        return acc;
    };
    for (int bj = 0; bj < plC.nblk; ++bj) {
        if (!dt.flag[(size_t)bi * dt.nblkC + bj]) continue;
        const Union2 win = nc_windows(g, plC, r, bj);
        const int n0 = win.n > 0 ? win.hi[0] - win.lo[0] + 1 : 0;
        const int nslots = n0 + (win.n > 1 ? win.hi[1] - win.lo[1] + 1 : 0);
        double cf[8];   // the pair's folded coefficients, in registers for all of its entries
#pragma unroll
        for (int t = 0; t < 8; ++t) cf[t] = dt.coef[((size_t)bi * dt.nblkC + bj) * 8 + t];
        for (int s0 = 0; s0 < nslots; s0 += 32) {
            // lane = n_c slot
            const int sl = s0 + lane;
            int nc = 0, cnt = 0, lo0 = 0, c0 = 0, lo1 = 0, jbase = 0;
            if (sl < nslots) {
                nc = sl < n0 ? win.lo[0] + sl : win.lo[1] + (sl - n0);
                const Segment sg = segment(g, plC, r, bj, nc);
                const Union2 u = union2(sg.dlo, sg.dhi, sg.xlo, sg.xhi);
                if (u.n > 0) { lo0 = u.lo[0]; c0 = u.hi[0] - u.lo[0] + 1; }
                cnt = c0;
                if (u.n > 1) { lo1 = u.lo[1]; cnt += u.hi[1] - u.lo[1] + 1; }
                jbase = sg.jbase;
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kAll, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(kAll, incl, 31);
            const int excl = incl - cnt;
            // lane = stored entry
            for (int p0 = 0; p0 < total; p0 += 32) {
                const int p = p0 + lane;
                // slot of entry p: the last slot whose offset is <= p (offsets ascend; empty slots share the offset of
                // their successor and lose against it)
                int at = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int e = __shfl_sync(kAll, excl, (at + step) & 31);
                    if (at + step < 32 && e <= p) at += step;
                }
                const int s_nc = __shfl_sync(kAll, nc, at), s_ex = __shfl_sync(kAll, excl, at);
                const int s_lo0 = __shfl_sync(kAll, lo0, at), s_c0 = __shfl_sync(kAll, c0, at);
                const int s_lo1 = __shfl_sync(kAll, lo1, at), s_jb = __shfl_sync(kAll, jbase, at);
                if (p < total) {
                    const int off = p - s_ex;
                    const int nd = off < s_c0 ? s_lo0 + off : s_lo1 + (off - s_c0);
                    const Cplx v = value(cf, s_nc, nd);
                    __stcs(idx + pos + p, (long long)s_jb + nd);
                    __stcs(dat + pos + p, make_double2(v.re, v.im));
                }
            }
            pos += total;
        }
    }
}

// count -> scan -> (optionally) fill + download.  index_ptr may be NULL (count only).
long long dip_block_run(bs2e_ctx* c, int q, const int64_t* sym1, long long n1, const int64_t* conf_n1,
                        const int64_t* conf_l1, const int64_t* sym2, long long n2, const int64_t* conf_n2,
                        const int64_t* conf_l2, bool compute, int64_t* index_ptr, int64_t* indices, double* data)
{
    if (!c->have_dip) throw Error("dipole block: call bs2e_set_radial_dipole first");
    if (!c->have_1p) throw Error("dipole block: call bs2e_set_one_particle first (overlap matrix)");
    HostDipPlan hp;
    try {
        hp = build_dip_plan(c->hg, c->dip_gauge, q, sym1, n1, conf_n1, conf_l1, sym2, n2, conf_n2, conf_l2, compute);
    } catch (const std::invalid_argument& e) {
        throw Error(e.what());
    }
    if (hp.empty) return 0;   // dip_block%init(shape, 0): no arrays are written (dipole.f90:26-30)
    cudaStream_t st = c->stream;
    DipArena dev;
    Plan plC{};
    plC.nblk = hp.cols.nblk;
    plC.n_config = (int)n2;
    plC.full = 1;
    plC.L = (int)sym2[0];
    plC.max_nd = hp.cols.max_nd;
    plC.blk = dev.up(hp.cols.blocks, st);
    plC.ncrow = dev.up(hp.cols.ncrow, st);
    DipTables dt = hp.tables();
    dt.flag = dev.up(hp.flag, st);
    dt.coef = dev.up(hp.coef, st);
    dt.row_n1 = dev.up(hp.rows.row_n1, st);
    dt.row_n2 = dev.up(hp.rows.row_n2, st);
    dt.row_blk = dev.up(hp.rows.row_blk, st);
    const long long nrows = n1;
    long long* d_cnt = dev.alloc<long long>(nrows + 1);
    long long* d_ptr = dev.alloc<long long>(nrows + 1);
    dip_count_kernel<<<(unsigned)((nrows + 1 + 127) / 128), 128, 0, st>>>(c->dg, plC, dt, d_cnt);
    BS2E_LAUNCHED();
    size_t tmp = 0;
    BS2E_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp, d_cnt, d_ptr, cub::Sum(), 1LL, nrows + 1, st));
    void* d_tmp = dev.alloc<char>(tmp ? tmp : 1);
    BS2E_CUDA(cub::DeviceScan::ExclusiveScan(d_tmp, tmp, d_cnt, d_ptr, cub::Sum(), 1LL, nrows + 1, st));
    g_launches.fetch_add(2);
    long long last = 0;
    BS2E_CUDA(cudaMemcpyAsync(&last, d_ptr + nrows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    BS2E_CUDA(cudaStreamSynchronize(st));
    const long long nnz = last - 1;
    if (!index_ptr || nnz == 0) return nnz;
    long long* d_idx = dev_alloc_async<long long>((size_t)nnz, st);
    double* d_dat = dev_alloc_async<double>(2 * (size_t)nnz, st);
    try {
        const DipBand bd{c->d_dipA, c->d_dipB ? c->d_dipB : c->d_dipA, c->d_Sb};
        const size_t tab_bytes = sizeof(double2) * kDipWarps * 6 * (2 * (size_t)c->hg.w + 2);
        dip_fill_kernel<<<(unsigned)((nrows + kDipWarps - 1) / kDipWarps), kDipWarps * 32, tab_bytes, st>>>(
            c->dg, plC, dt, bd, d_ptr, d_idx, reinterpret_cast<double2*>(d_dat));
        BS2E_LAUNCHED();
        BS2E_CUDA(cudaMemcpyAsync(index_ptr, d_ptr, sizeof(long long) * (nrows + 1), cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaMemcpyAsync(indices, d_idx, sizeof(long long) * nnz, cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaMemcpyAsync(data, d_dat, sizeof(double) * 2 * nnz, cudaMemcpyDeviceToHost, st));
        BS2E_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFreeAsync(d_idx, st);
        cudaFreeAsync(d_dat, st);
        throw;
    }
    cudaFreeAsync(d_idx, st);
    cudaFreeAsync(d_dat, st);
    return nnz;
}

}  // namespace bs2e
