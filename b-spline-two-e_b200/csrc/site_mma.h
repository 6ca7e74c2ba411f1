// site_mma.h -- launcher of the tensor-core site kernel (site_mma.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

struct bs2e_ctx;
struct bs2e_block;

namespace bs2e {

constexpr int kMmaThreads = 256;
constexpr size_t kMmaSmemLimit = 227 * 1024;   // per CTA; two CTAs per SM up to 113 KB
constexpr int kMmaStageRow = 512 + 16;          // bytes per row of a staging tile: 32 values + 16 B, so that the rows a
                                                // quarter warp writes to (0,2,4,6 / 1,3,5,7) start in different banks
constexpr int kMmaStageBytes = 8 * kMmaStageRow; // staging tile of one warp: 8 rows of 32 values
// sites without exchange windows stage four rows at a time (rows l3 = 0..3 of a quarter warp in different banks)
constexpr int kMmaHalfRow = 512 + 32;
constexpr int kMmaHalfBytes = 4 * kMmaHalfRow;

struct MmaSmem {    // element counts of the dynamic shared memory carve-up
    int ncmax;      // n_c slots
    int nseg;       // candidate segments (mask words per table row)
    int mstr;       // stride of the mask table rows in words (nseg | 1)
    int G;          // rows per group
    int cap;        // records per parity list
    int chrec;      // records per staging buffer (multiple of 8)
    int nl;         // l_max + 1 of the one-particle matrices
    int rowb;       // bytes per row of the CTA-level value staging (sites without exchange windows)
    // byte offsets of the pieces (kept in the constant bank by the kernel)
    int off_cfs, off_ob, off_recs, off_rcache, off_sblk, off_mtab, off_mraw, off_jb, off_ncq, off_cand, off_cprefix, off_stage,
        off_srow, off_tot, off_mbar, off_misc;
    size_t bytes;
};

MmaSmem mma_layout(const bs2e_ctx* c, int nblk, int maxc, int maxrec, bool wx, int lmax, bool bulk);
bool site_mma_usable(const bs2e_ctx* c, int nblk, int maxc, int maxrec, int lmax);
// the two launches of a block: sites with exchange windows on stX, the others on stD
void launch_site_mma(bs2e_block* b, cudaStream_t stX, cudaStream_t stD);

}  // namespace bs2e
