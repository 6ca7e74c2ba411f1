/*
 * bs2e.h -- C ABI of libbs2e_gpu.so: the B200 (sm_100a) implementation of the
 * two-electron matrix-element hot path of edvinolo/b-spline-two-e.
 *
 * The reference has no FFI layer for this path; the seams are three Fortran
 * calls in src/apps/main_basis_setup.f90:
 *     :80   call setup_Slater_integrals(splines,max_k,k_GL,r_k,r_m_k,r_d_k)
 *     :85   call compute_R_k_map(r_d_k,r_k,r_m_k,splines,max_k,R_p)
 *     :108  call construct_block_tensor(H_vec,S,splines,bas%syms(i),max_k,R_p,
 *                                       H_diag%blocks(i),S_diag%blocks(i),full)
 * Each entry point below names the reference routine it stands in for.  The
 * ISO_C_BINDING module a maintainer adds on the Fortran side is
 * fortran/bs2e_gpu_binding.f90.
 *
 * Conventions: every integer is int64_t (the reference is built with
 * -fdefault-integer-8), reals are double, complex numbers are interleaved
 * (re,im) doubles, arrays are column-major, spline / configuration / CSR
 * indices are 1-based exactly as the reference stores them.  Every function
 * returns 0 on success and a non-zero status otherwise; the message is
 * available from bs2e_last_error() (thread local).  There is no CPU fallback:
 * without a usable CUDA device every compute entry point fails.
 */
#ifndef BS2E_H
#define BS2E_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bs2e_ctx bs2e_ctx;      /* device-resident basis + R^k tensor */
typedef struct bs2e_block bs2e_block;  /* one symmetry block being assembled */

const char *bs2e_last_error(void);
int bs2e_device_count(int64_t *count);

/* Stands in for the state the reference keeps in type(b_spline)
 * (src/tools/bspline_tools.f90:4-40) plus the Gauss-Legendre rule that
 * setup_GL(k_GL,-1,1,x,w) returns (src/tools/quad_tools.f90:14-27; the nodes
 * come from fortran-stdlib on the caller's side so they are bit-identical to
 * what the reference integrates with).  knots has n_knots = n + k entries.   */
int bs2e_ctx_create(int64_t k_spline, int64_t n_knots, const double *knots,
                    int64_t max_k, int64_t k_GL, const double *gl_x,
                    const double *gl_w, int64_t device, bs2e_ctx **ctx);
/* Destroys the context and everything it holds on the device.  Block, configuration
 * and dipole-block handles made from it must be freed BEFORE this call; blocks parked
 * by bs2e_block_count for a later bs2e_block_fill are released here.               */
int bs2e_ctx_destroy(bs2e_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t). */
int bs2e_ctx_set_stream(bs2e_ctx *ctx, void *cuda_stream);
int bs2e_ctx_sync(bs2e_ctx *ctx);

/* sparse_4d%count_nnz / sparse_6d%count_nnz / count_nnz_R_k
 * (src/tools/sparse_array_tools.f90:276-366); P*P = count_nnz_R_k.          */
int bs2e_sizes(bs2e_ctx *ctx, int64_t *n_b, int64_t *cells, int64_t *P,
               int64_t *nnz_4d, int64_t *nnz_6d);

/* ---- stage A: setup_Slater_integrals (src/mat_els/mat_els.f90:172-292,
 *      392-491).  Results stay on the device.                               */
int bs2e_slater_cells(bs2e_ctx *ctx);
/* Copy the cell integrals back in the reference's entry order so that the
 * Fortran sparse_4d / sparse_6d objects can be populated:
 * r_k, r_m_k: data(nnz_4d, 0:max_k); iv,i,j: (nnz_4d)  (mat_els.f90:430-438)
 * r_d_k: data(nnz_6d, 0:max_k); iv,i,j,i_p,j_p: (nnz_6d) (mat_els.f90:484-490)
 * Any pointer may be NULL to skip that array.                               */
int bs2e_get_r_k(bs2e_ctx *ctx, double *r_k, double *r_m_k,
                 int64_t *iv, int64_t *i, int64_t *j);
int bs2e_get_r_d_k(bs2e_ctx *ctx, double *r_d_k, int64_t *iv, int64_t *i,
                   int64_t *j, int64_t *i_p, int64_t *j_p);

/* ---- stage B: compute_R_K_map (src/tools/sparse_array_tools.f90:452-493).
 *      The R^k tensor stays on the device; the Fortran Nd_DOK becomes a
 *      holder of the context handle.                                        */
int bs2e_rk_build(bs2e_ctx *ctx);
/* Multi-GPU (the reference runs this stage once per process,
 * src/apps/main_basis_setup.f90:80-85): a GPU that assembles a share of the rows of
 * every symmetry block reads only the rows R^k(a .; . .) of the tensor whose first
 * spline index a belongs to its radial sites.  After bs2e_rk_rows(ctx, a_lo, a_hi)
 * bs2e_slater_cells computes the same-cell integrals of the cells those rows live on
 * and bs2e_rk_build the rows a_lo <= a <= a_hi only; the default is 1..n_b.
 * bs2e_block_fill of rows that read outside the built slice is an error, and so are
 * the getters (bs2e_get_r_d_k, bs2e_rk_get, bs2e_rk_plane) on a partial tensor.     */
int bs2e_rk_rows(bs2e_ctx *ctx, int64_t a_lo, int64_t a_hi);
/* Nd_DOK%get_val (sparse_array_tools.f90:536-555) for n_keys keys (a,b,c,d):
 * keys is (4, n_keys), vals is (max_k+1, n_keys); a key outside the band
 * structure is an error, as in the reference.                               */
int bs2e_rk_get(bs2e_ctx *ctx, int64_t n_keys, const int64_t *keys, double *vals);
/* One multipole plane as a dense (P,P) matrix, out[p1*P + p2]; pair numbering
 * is a-major / c ascending over ordered band pairs (a,c).                   */
int bs2e_rk_plane(bs2e_ctx *ctx, int64_t k, double *out);

/* ---- stage C inputs: H_vec(0:max_l_1p)%data and S as passed to
 *      construct_block_tensor (hamiltonian.f90:106-108): dense complex
 *      n_b x n_b column-major matrices, H_vec concatenated over l.          */
int bs2e_set_one_particle(bs2e_ctx *ctx, int64_t max_l_1p, const double *H_vec,
                          const double *S);

/* The same matrices computed ON THE DEVICE: setup_S and setup_H_one_particle for
 * l = 0..max_l_1p (src/mat_els/mat_els.f90:47-118,294-346; V = l(l+1)/(2r^2) - Z/r,
 * src/mat_els/potentials.f90:35-43; CAP -i eta (r-r0)^order for r >= r0, src/tools/CAP_tools.f90:24-34)
 * with the Gauss-Legendre rule of the context.  Replaces bs2e_set_one_particle; bs2e_get_one_particle
 * copies the dense n_b x n_b column-major matrices back (H_vec concatenated over l; either may be NULL)
 * so that the Fortran H_vec / S stay populated for the consumers that write them out.                 */
int bs2e_one_particle_device(bs2e_ctx *ctx, int64_t Z, int64_t max_l_1p, int64_t CAP_order,
                             double CAP_r_0, double CAP_eta_re, double CAP_eta_im);
int bs2e_get_one_particle(bs2e_ctx *ctx, double *H_vec, double *S);

/* ---- stage C: one symmetry block.
 * conf_n / conf_l are (2, n_config): term%configs(:)%n and %l
 * (src/tools/orbital_tools.f90:15-19) in the order count_configs generates
 * (orbital_tools.f90:157-193).  L is term%l; full as in the reference.
 *
 * bs2e_block_count   = count_nnz            (hamiltonian.f90:348-416)
 * bs2e_block_fill    = construct_block_tensor (hamiltonian.f90:106-283),
 *                      filling arrays the caller allocated from the counts:
 *                      index_ptr(n_config+1), indices(nnz), data(nnz).
 * The counts are those of the pattern the reference EMITS (they coincide with
 * count_nnz whenever the reference itself does not overrun its arrays, see
 * SURVEY.md F5).                                                            */
int bs2e_block_count(bs2e_ctx *ctx, int64_t L, int64_t n_config,
                     const int64_t *conf_n, const int64_t *conf_l, int64_t full,
                     int64_t *nnz_H, int64_t *nnz_S);
int bs2e_block_fill(bs2e_ctx *ctx, int64_t L, int64_t n_config,
                    const int64_t *conf_n, const int64_t *conf_l, int64_t full,
                    int64_t *H_ptr, int64_t *H_idx, double *H_dat,
                    int64_t *S_ptr, int64_t *S_idx, double *S_dat);

/* Split form of the same work, used for device-resident pipelines and for
 * sharding rows over GPUs: rows row_lo..row_hi (1-based, inclusive) of the
 * block are planned/counted, assembled on the device and downloaded as a CSR
 * fragment whose index_ptr starts at 1.                                     */
int bs2e_block_plan(bs2e_ctx *ctx, int64_t L, int64_t n_config,
                    const int64_t *conf_n, const int64_t *conf_l, int64_t full,
                    int64_t row_lo, int64_t row_hi, bs2e_block **blk);
/* The same for a union of n_ranges ascending, disjoint row ranges
 * [range_lo[q], range_hi[q]]; the fragment holds those rows in ascending order.
 * This is the unit of the multi-GPU partition: rows are dealt to GPUs by the
 * first radial index n(1) of their configuration, so that all rows of a radial
 * site (which share their R^k values) stay on one GPU; inside every (l1,l2)
 * group of the configuration list such rows form one contiguous range.      */
int bs2e_block_plan_ranges(bs2e_ctx *ctx, int64_t L, int64_t n_config,
                           const int64_t *conf_n, const int64_t *conf_l, int64_t full,
                           int64_t n_ranges, const int64_t *range_lo,
                           const int64_t *range_hi, bs2e_block **blk);
/* A configuration list kept resident on the device (both arrays (2, n_config) as above),
 * for pipelines whose inputs already live in HBM: bs2e_block_plan_dev plans the rows
 * [range_lo[q], range_hi[q]] (n_ranges = 0: the whole block) without touching host copies
 * of the list.  The plan itself is built on the device in every case (group structure,
 * radial sites, row counts); the only host arithmetic is the 3j/6j table of the (l1,l2)
 * group pairs, cached per context.                                                    */
typedef struct bs2e_configs bs2e_configs;
int bs2e_configs_upload(bs2e_ctx *ctx, int64_t n_config, const int64_t *conf_n,
                        const int64_t *conf_l, bs2e_configs **cfg);
int bs2e_configs_free(bs2e_configs *cfg);
int bs2e_block_plan_dev(bs2e_ctx *ctx, int64_t L, bs2e_configs *cfg, int64_t full,
                        int64_t n_ranges, const int64_t *range_lo, const int64_t *range_hi,
                        bs2e_block **blk);
int bs2e_block_nnz(bs2e_block *blk, int64_t *nnz_H, int64_t *nnz_S);
/* per-row entry counts of the planned rows (one value per planned row each) */
int bs2e_block_row_counts(bs2e_block *blk, int64_t *cnt_H, int64_t *cnt_S);
/* Repeat the count pass + scan of an existing plan with everything already
 * resident on the device (no host synchronisation); used for device-timed
 * benchmarking of the count stage. */
int bs2e_block_recount(bs2e_block *blk);
int bs2e_block_assemble(bs2e_block *blk);
/* Count pass (when recount != 0) and bs2e_block_assemble of n planned blocks of
 * one context in one call.  Consecutive blocks are issued on different internal
 * streams, so the count pass and the last wave of one block's fill overlap the
 * next block; the work is ordered after what is queued on the context's stream,
 * which in turn waits for all of it.  Stands in for the loop over symmetries of
 * src/apps/main_basis_setup.f90:105-116 when the CSR fragments stay on the device. */
int bs2e_blocks_run(bs2e_ctx *ctx, int64_t n, bs2e_block **blks, int64_t recount);
int bs2e_block_download(bs2e_block *blk, int64_t *H_ptr, int64_t *H_idx, double *H_dat,
                        int64_t *S_ptr, int64_t *S_idx, double *S_dat);
/* 64-bit checksums of the device-resident fragment (indices and raw data
 * bits), for runs whose output is too large to bring to the host.           */
int bs2e_block_checksum(bs2e_block *blk, uint64_t *sum_H, uint64_t *sum_S);
/* Releases the fragment and the plan; the device memory goes back to the pool in stream order
 * (after the work already queued on the context's stream), the call does not wait for the device. */
int bs2e_block_free(bs2e_block *blk);

/* Pinned host memory for callers that want full-speed transfers; the pages are
 * placed on the NUMA node the current CUDA device is attached to.           */
int bs2e_host_alloc(int64_t bytes, void **ptr);
int bs2e_host_free(void *ptr);

/* ---- dipole blocks (the next stage of basis_setup, main_basis_setup.f90:125-152).
 * bs2e_set_radial_dipole: type(radial_dipole) as setup_radial_dip fills it
 *   (src/mat_els/mat_els.f90:14-19,120-170): dense complex n_b x n_b column-major;
 *   gauge 'l' (108): A = r_mat, B ignored; gauge 'v' (118): A = dr_mat, B = r_inv_mat.
 * bs2e_dip_block_count / _fill = init_dip_block + construct_dip_block_tensor
 *   (src/mat_els/dipole.f90:8-47,87-146) for the block <sym1| d_q |sym2>: sym = (l, m, pi),
 *   rows are the configurations of sym1, columns those of sym2, q in {-1,0,1}; compute as
 *   in the reference (a block that is forbidden or not computed has nnz = 0 and no arrays).
 *   The overlap matrix is the one given to bs2e_set_one_particle.  index_ptr has
 *   n_config1+1 entries; indices / data have the nnz of the count call.              */
int bs2e_set_radial_dipole(bs2e_ctx *ctx, int64_t gauge, const double *A, const double *B);
/* setup_radial_dip on the device (mat_els.f90:120-170,348-390), in place of bs2e_set_radial_dipole;
 * bs2e_get_radial_dipole copies A (and B in the velocity gauge) back as dense matrices.            */
int bs2e_radial_dipole_device(bs2e_ctx *ctx, int64_t gauge);
int bs2e_get_radial_dipole(bs2e_ctx *ctx, double *A, double *B);
int bs2e_dip_block_count(bs2e_ctx *ctx, int64_t q, const int64_t *sym1, int64_t n_config1,
                         const int64_t *conf_n1, const int64_t *conf_l1, const int64_t *sym2,
                         int64_t n_config2, const int64_t *conf_n2, const int64_t *conf_l2,
                         int64_t compute, int64_t *nnz);
int bs2e_dip_block_fill(bs2e_ctx *ctx, int64_t q, const int64_t *sym1, int64_t n_config1,
                        const int64_t *conf_n1, const int64_t *conf_l1, const int64_t *sym2,
                        int64_t n_config2, const int64_t *conf_n2, const int64_t *conf_l2,
                        int64_t compute, int64_t *index_ptr, int64_t *indices, double *data);

/* ---- result files of basis_setup, written natively (no Fortran runtime):
 *      gfortran unformatted sequential records, 8-byte default integers.  Lets a
 *      driver hand the GPU-built matrices to the reference's consumers (diag,
 *      quasi, time_prop read them with CS_block_diag_load, load_basis,
 *      load_bsplines) and stream a block that does not fit host staging.
 * H_diag.dat / S_diag.dat: block_diag_CS%store (src/tools/block_tools.f90:458-485);
 * basis.dat: basis%store (src/tools/orbital_tools.f90:364-387);
 * splines.dat: b_spline%store (src/tools/bspline_tools.f90:375-386).            */
typedef struct bs2e_file bs2e_file;
int bs2e_file_create_block_diag(const char *path, int64_t n_blocks,
                                const int64_t *block_rows, bs2e_file **f);
/* D_q.dat: block_CS%store (src/tools/block_tools.f90:386-415); the blocks are then written
 * with bs2e_file_write_block in column-major order (block column outer, block row inner) */
int bs2e_file_create_block_matrix(const char *path, int64_t n_block_rows, int64_t n_block_cols,
                                  const int64_t *block_rows, const int64_t *block_cols,
                                  bs2e_file **f);
int bs2e_file_write_block(bs2e_file *f, int64_t rows, int64_t cols, int64_t nnz,
                          const int64_t *index_ptr, const int64_t *indices,
                          const double *data);
/* the same block given as CSR fragments of consecutive row ranges (each with an
 * index_ptr starting at 1, as bs2e_block_download returns them)              */
int bs2e_file_write_block_fragments(bs2e_file *f, int64_t rows, int64_t cols,
                                    int64_t n_frag, const int64_t *frag_rows,
                                    const int64_t *const *frag_ptr,
                                    const int64_t *const *frag_idx,
                                    const double *const *frag_dat);
int bs2e_file_close(bs2e_file *f);
/* conf_n[q], conf_l[q]: (2, n_config[q]); conf_eqv[q]: (n_config[q]) */
int bs2e_file_write_basis(const char *path, int64_t max_l_1p, int64_t max_L,
                          int64_t two_el, int64_t n_sym, const int64_t *sym_l,
                          const int64_t *sym_m, const int64_t *sym_pi,
                          const int64_t *n_config, const int64_t *const *conf_n,
                          const int64_t *const *conf_l, const int64_t *const *conf_eqv);
int bs2e_file_write_splines(const char *path, int64_t k, int64_t n_knots,
                            const double *knots);
/* record-level reader of the same files (CS_block_diag_load, block_tools.f90:487-524) */
int bs2e_file_open(const char *path, bs2e_file **f);
int bs2e_file_next_record(bs2e_file *f, int64_t *nbytes);
int bs2e_file_record_data(bs2e_file *f, void *dst, int64_t nbytes);
/* test hook: split records into subrecords of at most this many bytes (0: default 2^31-9) */
int bs2e_file_set_max_subrecord(int64_t bytes);

/* Number of kernels this library has launched since load (all contexts). */
int64_t bs2e_launch_count(void);
/* Profiling aid: cycles per phase of the site fill kernel summed over its CTAs (zeros unless the
 * library was built with -DBS2E_PHASE_TIMING): [0] phases 0-1, [1] phase 2, [2] 3a, [3] 3b, [4] 3c,
 * [5] phase 4.                                                                      */
int bs2e_debug_site_phase_cycles(uint64_t *out8, int64_t reset);

/* ---- host-side companions of the path (no GPU needed).  They restate the
 *      cheap reference routines whose OUTPUT feeds the hot path, so that a
 *      driver without the Fortran program (tests, bench.py) can produce the
 *      same inputs.                                                         */
/* grid_tools.f90:6-55 generate_grid; returns the number of knots (or -1). */
int64_t bs2e_host_generate_grid(int64_t k, int64_t m, int64_t Z, double h_max,
                                double r_max, double *grid, int64_t cap);
/* quad_tools.f90:14-27 setup_GL on [a,b] */
int bs2e_host_gauss_legendre(int64_t N, double a, double b, double *x, double *w);
/* bspline_tools.f90:364-373 find_max_n_b */
int64_t bs2e_host_find_max_n_b(int64_t k, int64_t n_knots, const double *knots, double x);
/* mat_els.f90:85-118 setup_S and :47-83 setup_H_one_particle (hydrogenic
 * potential, potentials.f90:35-43; CAP, CAP_tools.f90:24-34)                */
int bs2e_host_setup_S(int64_t k, int64_t n_knots, const double *knots, int64_t k_GL, double *S);
int bs2e_host_setup_H_one_particle(int64_t k, int64_t n_knots, const double *knots,
                                   int64_t Z, int64_t l, int64_t CAP_order, double CAP_r_0,
                                   double CAP_eta_re, double CAP_eta_im, int64_t k_GL, double *H);
/* mat_els.f90:120-170 setup_radial_dip: gauge 'l' (108): A = r_mat; 'v' (118): A = dr_mat, B = r_inv_mat */
int bs2e_host_setup_radial_dip(int64_t k, int64_t n_knots, const double *knots, int64_t k_GL,
                               int64_t gauge, double *A, double *B);
/* orbital_tools.f90:245-343 init_basis (two_el): symmetry list, then
 * count_configs (:119-216) per symmetry.                                    */
int64_t bs2e_host_basis_syms(int64_t max_L, int64_t z_pol, int64_t *sym_l,
                             int64_t *sym_m, int64_t *sym_pi, int64_t cap);
int64_t bs2e_host_count_configs(int64_t term_l, int64_t term_pi, int64_t max_l_1p,
                                int64_t n_b, int64_t k_spline, int64_t max_n_b,
                                int64_t n_all_l, int64_t l_2_max, int64_t *conf_n,
                                int64_t *conf_l, int64_t *conf_eqv, int64_t cap);
/* wigner_tools.f90:30-60,126-138 (exact arithmetic instead of GSL) */
double bs2e_host_three_j0(int64_t ja, int64_t jb, int64_t jc);
double bs2e_host_six_j(int64_t ja, int64_t jb, int64_t jc, int64_t jd, int64_t je, int64_t jf);
double bs2e_host_ang_k_LS(int64_t k, int64_t la, int64_t lb, int64_t lc, int64_t ld, int64_t L);

#ifdef __cplusplus
}
#endif
#endif /* BS2E_H */
