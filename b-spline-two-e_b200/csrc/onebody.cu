// onebody.cu -- one-particle matrices and radial dipole integrals on the device:
// removes the last host-side numerical inputs (H_vec, S, radial_dipole) from the boundary.
// Stands in for setup_S / setup_H_one_particle / setup_radial_dip (src/mat_els/mat_els.f90:47-170); the
// results land in the band storage the stage-C kernels read, and can be fetched back as the dense Fortran
// matrices.  Latency-bound, microseconds of work: one thread per band entry (onebody_core.h).
#include "ctx.h"
#include "onebody_core.h"

namespace bs2e {

__global__ void one_body_kernel(Geom g, OneBodyParams p, OneBodyOut o)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int bw = 2 * g.w + 1;
    if (idx >= g.nb * bw) return;
    one_body_entry(g, p, o, idx / bw + 1, idx % bw);
}

static void run_one_body(bs2e_ctx* c, const OneBodyParams& p)
{
    const Geom& g = c->dg;
    const size_t per = band_doubles(c->hg);
    OneBodyOut o{nullptr, nullptr, nullptr, nullptr};
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
    if (p.want_1p) {
        cudaFree(c->d_Hb); cudaFree(c->d_Sb);
        c->d_Hb = c->d_Sb = nullptr;
        c->d_Hb = dev_alloc<double>((size_t)(p.lmax + 1) * per);
        c->d_Sb = dev_alloc<double>(per);
        BS2E_CUDA(cudaMemsetAsync(c->d_Hb, 0, sizeof(double) * (size_t)(p.lmax + 1) * per, c->stream));
        BS2E_CUDA(cudaMemsetAsync(c->d_Sb, 0, sizeof(double) * per, c->stream));
        o.Hb = c->d_Hb;
        o.Sb = c->d_Sb;
    }
    if (p.gauge) {
        cudaFree(c->d_dipA); cudaFree(c->d_dipB);
        c->d_dipA = c->d_dipB = nullptr;
        c->d_dipA = dev_alloc<double>(per);
        BS2E_CUDA(cudaMemsetAsync(c->d_dipA, 0, sizeof(double) * per, c->stream));
        o.A = c->d_dipA;
        if (p.gauge == 'v') {
            c->d_dipB = dev_alloc<double>(per);
            BS2E_CUDA(cudaMemsetAsync(c->d_dipB, 0, sizeof(double) * per, c->stream));
            o.B = c->d_dipB;
        }
    }
    const int n = g.nb * (2 * g.w + 1);
    one_body_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(g, p, o);
    BS2E_LAUNCHED();
}

void one_particle_device(bs2e_ctx* c, int Z, int lmax, int cap_order, double cap_r0, double eta_re, double eta_im)
{
    if (lmax < 0 || lmax >= kMaxL1p) throw Error("bs2e_one_particle_device: max_l_1p out of range");
    run_one_body(c, OneBodyParams{Z, lmax, cap_order, cap_r0, eta_re, eta_im, 1, 0});
    c->lmax_1p = lmax;
    c->have_1p = true;
}

void radial_dipole_device(bs2e_ctx* c, int gauge)
{
    if (gauge != 'l' && gauge != 'v') throw Error("bs2e_radial_dipole_device: gauge must be 'l' (108) or 'v' (118)");
    run_one_body(c, OneBodyParams{0, 0, 0, 0.0, 0.0, 0.0, 0, gauge});
    c->dip_gauge = gauge;
    c->have_dip = true;
}

// band storage -> dense complex n_b x n_b column-major (zeros outside the band), as the Fortran types hold it
static void fetch_band(bs2e_ctx* c, const double* d_band, double* dense)
{
    const Geom& g = c->hg;
    const size_t per = band_doubles(g);
    std::vector<double> band(per);
    BS2E_CUDA(cudaMemcpyAsync(band.data(), d_band, sizeof(double) * per, cudaMemcpyDeviceToHost, c->stream));
    BS2E_CUDA(cudaStreamSynchronize(c->stream));
    const int nb = g.nb, w = g.w, bw = 2 * w + 1;
    for (size_t q = 0; q < (size_t)nb * nb * 2; ++q) dense[q] = 0.0;
    for (int n = 1; n <= nb; ++n)
        for (int d = 0; d < bw; ++d) {
            const int np = n + d - w;
            if (np < 1 || np > nb) continue;
            const size_t at = 2 * ((size_t)(n - 1) + (size_t)nb * (np - 1));
            dense[at] = band[((size_t)n * bw + d) * 2];
            dense[at + 1] = band[((size_t)n * bw + d) * 2 + 1];
        }
}

void fetch_one_particle(bs2e_ctx* c, double* H_vec, double* S)
{
    if (!c->have_1p) throw Error("bs2e_get_one_particle: no one-particle matrices on the device");
    const Geom& g = c->hg;
    if (S) fetch_band(c, c->d_Sb, S);
    if (H_vec)
        for (int l = 0; l <= c->lmax_1p; ++l)
            fetch_band(c, c->d_Hb + (size_t)l * band_doubles(g), H_vec + (size_t)l * g.nb * g.nb * 2);
}

void fetch_radial_dipole(bs2e_ctx* c, double* A, double* B)
{
    if (!c->have_dip) throw Error("bs2e_get_radial_dipole: no radial dipole integrals on the device");
    if (A) fetch_band(c, c->d_dipA, A);
    if (B) {
        if (c->dip_gauge != 'v') throw Error("bs2e_get_radial_dipole: r_inv_mat exists in the velocity gauge only");
        fetch_band(c, c->d_dipB, B);
    }
}

}  // namespace bs2e
