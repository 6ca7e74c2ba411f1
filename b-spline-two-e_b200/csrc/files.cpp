// files.cpp -- native writer / reader of the result files of basis_setup, so that a
// driver without the Fortran program can hand the GPU-built matrices to the
// reference's consumers (diag, quasi, time_prop, ...) and stream large outputs block by
// block.  Host code only.
//
// Formats: Fortran unformatted SEQUENTIAL files as gfortran writes them
// with -fdefault-integer-8: every WRITE statement is one record framed by 4-byte
// length markers; default INTEGER and LOGICAL are 8 bytes.  A record longer than
// 2^31-9 bytes is split into subrecords: the leading marker of a subrecord is
// negative when another subrecord follows, the trailing marker is negative when the
// subrecord has a predecessor (libgfortran, GFC_MAX_SUBRECORD_LENGTH = 2147483639).
//
//   H_diag.dat / S_diag.dat   CS_block_diag_store   src/tools/block_tools.f90:458-485
//   single CSR matrix         store_CS              src/tools/sparse_array_tools.f90:695-708
//   basis.dat                 store_basis           src/tools/orbital_tools.f90:364-387
//   splines.dat               store_bsplines        src/tools/bspline_tools.f90:375-386
#include "files.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace bs2e {
namespace files {

namespace {
long long g_max_subrecord = 2147483639LL;
}
void set_max_subrecord(long long bytes) { g_max_subrecord = bytes > 0 ? bytes : 2147483639LL; }

struct Writer::Impl {
    FILE* f = nullptr;
    std::string path;
    long long blocks_expected = -1, blocks_written = 0;
};

static void put(FILE* f, const void* p, size_t n, const std::string& path)
{
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("write to " + path + " failed");
}

// one Fortran record made of the given byte ranges
static void record(FILE* f, const std::string& path, const void* const* parts, const size_t* sizes, int nparts)
{
    unsigned long long total = 0;
    for (int q = 0; q < nparts; ++q) total += sizes[q];
    // stream the concatenation through subrecords of at most g_max_subrecord bytes
    int part = 0;
    size_t off = 0;
    unsigned long long left = total;
    bool first = true;
    do {
        const long long len = (long long)(left < (unsigned long long)g_max_subrecord ? left : (unsigned long long)g_max_subrecord);
        const bool more = (unsigned long long)len < left;
        const int32_t lead = (int32_t)(more ? -len : len), trail = (int32_t)(first ? len : -len);
        put(f, &lead, 4, path);
        long long todo = len;
        while (todo > 0) {
            const size_t chunk = std::min<size_t>((size_t)todo, sizes[part] - off);
            put(f, (const char*)parts[part] + off, chunk, path);
            off += chunk;
            todo -= (long long)chunk;
            if (off == sizes[part]) { ++part; off = 0; }
        }
        put(f, &trail, 4, path);
        left -= (unsigned long long)len;
        first = false;
    } while (left > 0);
}
static void record1(FILE* f, const std::string& path, const void* p, size_t n)
{
    const void* parts[1] = {p};
    const size_t sizes[1] = {n};
    record(f, path, parts, sizes, 1);
}

Writer::Writer(const std::string& path) : impl(new Impl())
{
    impl->path = path;
    impl->f = fopen(path.c_str(), "wb");
    if (!impl->f) {
        delete impl;
        throw std::runtime_error("cannot open " + path + " for writing");
    }
}
Writer::~Writer()
{
    if (impl->f) fclose(impl->f);
    delete impl;
}
void Writer::close()
{
    if (impl->blocks_expected >= 0 && impl->blocks_written != impl->blocks_expected)
        throw std::runtime_error(impl->path + ": fewer blocks written than announced");
    if (impl->f && fclose(impl->f) != 0) { impl->f = nullptr; throw std::runtime_error("close of " + impl->path + " failed"); }
    impl->f = nullptr;
}
void Writer::raw_record(const void* p, size_t n) { record1(impl->f, impl->path, p, n); }

// CS_block_diag_store header: 'CSR', block_shape(2), shape(2)   (block_tools.f90:468-470)
void Writer::block_diag_header(long long n_blocks, const int64_t* block_rows)
{
    raw_record("CSR", 3);
    const int64_t bshape[2] = {n_blocks, n_blocks};
    int64_t tot = 0;
    for (long long q = 0; q < n_blocks; ++q) tot += block_rows[q];
    const int64_t shape[2] = {tot, tot};   // compute_shape_block_diag_CS: sum of the (square) block shapes
    raw_record(bshape, sizeof(bshape));
    raw_record(shape, sizeof(shape));
    impl->blocks_expected = n_blocks;
}
// CS_block_store header (D_q.dat, block_tools.f90:386-396): 'CSR', block_shape(2), shape(2) with
// shape = sums over the diagonal blocks (compute_shape_block_CS, :127-140); the blocks follow in
// column-major order (j outer, i inner), each written like a block of the block-diagonal file
void Writer::block_matrix_header(long long nbr, long long nbc, const int64_t* block_rows, const int64_t* block_cols)
{
    raw_record("CSR", 3);
    const int64_t bshape[2] = {nbr, nbc};
    int64_t r = 0, c = 0;
    for (long long q = 0; q < nbr; ++q) r += block_rows[q];
    for (long long q = 0; q < nbc; ++q) c += block_cols[q];
    const int64_t shape[2] = {r, c};
    raw_record(bshape, sizeof(bshape));
    raw_record(shape, sizeof(shape));
    impl->blocks_expected = nbr * nbc;
}
// one block: shape(2); nnz; if nnz > 0: index_ptr; indices; data   (block_tools.f90:472-481)
void Writer::csr_block(long long rows, long long cols, long long nnz, const int64_t* index_ptr,
                       const int64_t* indices, const double* data)
{
    if (nnz > 0 && (index_ptr[0] != 1 || index_ptr[rows] != nnz + 1))
        throw std::runtime_error(impl->path + ": index_ptr does not match nnz (1-based CSR expected)");
    const int64_t shape[2] = {rows, cols};
    const int64_t n = nnz;
    raw_record(shape, sizeof(shape));
    raw_record(&n, sizeof(n));
    if (nnz > 0) {
        raw_record(index_ptr, sizeof(int64_t) * (size_t)(rows + 1));
        raw_record(indices, sizeof(int64_t) * (size_t)nnz);
        raw_record(data, sizeof(double) * 2 * (size_t)nnz);
    }
    ++impl->blocks_written;
}
// the same block delivered as fragments of consecutive row ranges (the streaming form:
// a block that does not fit host staging is downloaded range by range): the three
// records of the block are written from the pieces without assembling them.
void Writer::csr_block_fragments(long long rows, long long cols, int nfrag, const long long* frag_rows,
                                 const int64_t* const* frag_ptr, const int64_t* const* frag_idx,
                                 const double* const* frag_dat)
{
    long long nnz = 0, r = 0;
    for (int q = 0; q < nfrag; ++q) { nnz += frag_ptr[q][frag_rows[q]] - 1; r += frag_rows[q]; }
    if (r != rows) throw std::runtime_error(impl->path + ": fragments do not cover the rows of the block");
    const int64_t shape[2] = {rows, cols};
    const int64_t n = nnz;
    raw_record(shape, sizeof(shape));
    raw_record(&n, sizeof(n));
    if (nnz > 0) {
        std::vector<int64_t> ptr((size_t)rows + 1);
        long long run = 0, row = 0;
        for (int q = 0; q < nfrag; ++q) {
            if (frag_ptr[q][0] != 1) throw std::runtime_error(impl->path + ": fragment index_ptr must start at 1");
            for (long long i = 0; i < frag_rows[q]; ++i) ptr[(size_t)row++] = frag_ptr[q][i] + run;
            run += frag_ptr[q][frag_rows[q]] - 1;
        }
        ptr[(size_t)rows] = run + 1;
        raw_record(ptr.data(), sizeof(int64_t) * ptr.size());
        std::vector<const void*> parts(nfrag);
        std::vector<size_t> sizes(nfrag);
        for (int q = 0; q < nfrag; ++q) { parts[q] = frag_idx[q]; sizes[q] = sizeof(int64_t) * (size_t)(frag_ptr[q][frag_rows[q]] - 1); }
        record(impl->f, impl->path, parts.data(), sizes.data(), nfrag);
        for (int q = 0; q < nfrag; ++q) { parts[q] = frag_dat[q]; sizes[q] = sizeof(double) * 2 * (size_t)(frag_ptr[q][frag_rows[q]] - 1); }
        record(impl->f, impl->path, parts.data(), sizes.data(), nfrag);
    }
    ++impl->blocks_written;
}
// store_CS: type tag, shape, nnz, index_ptr, indices, data   (sparse_array_tools.f90:695-708)
void Writer::single_csr(long long rows, long long cols, long long nnz, const int64_t* index_ptr,
                        const int64_t* indices, const double* data)
{
    const int64_t shape[2] = {rows, cols};
    const int64_t n = nnz;
    raw_record("CSR", 3);
    raw_record(shape, sizeof(shape));
    raw_record(&n, sizeof(n));
    raw_record(index_ptr, sizeof(int64_t) * (size_t)(rows + 1));
    raw_record(indices, sizeof(int64_t) * (size_t)nnz);
    raw_record(data, sizeof(double) * 2 * (size_t)nnz);
}

// store_basis (orbital_tools.f90:373-387); a config is n(2), l(2), eqv = 5 x 8 bytes
void write_basis(const std::string& path, long long max_l_1p, long long max_L, bool two_el, long long n_sym,
                 const int64_t* sym_l, const int64_t* sym_m, const int64_t* sym_pi, const int64_t* n_config,
                 const int64_t* const* conf_n, const int64_t* const* conf_l, const int64_t* const* conf_eqv)
{
    Writer w(path);
    const int64_t a = max_l_1p, b = max_L, c = two_el ? 1 : 0, d = n_sym;
    int64_t n_states = 0;
    std::vector<int64_t> sym_ptr((size_t)n_sym + 1);
    for (long long q = 0; q < n_sym; ++q) { sym_ptr[(size_t)q] = n_states + 1; n_states += n_config[q]; }
    sym_ptr[(size_t)n_sym] = n_states + 1;
    w.raw_record(&a, 8); w.raw_record(&b, 8); w.raw_record(&c, 8); w.raw_record(&d, 8);
    w.raw_record(&n_states, 8);
    w.raw_record(sym_ptr.data(), 8 * sym_ptr.size());
    for (long long q = 0; q < n_sym; ++q) {
        const int64_t pi = sym_pi[q] ? 1 : 0;
        w.raw_record(&sym_l[q], 8); w.raw_record(&sym_m[q], 8); w.raw_record(&pi, 8); w.raw_record(&n_config[q], 8);
        std::vector<int64_t> cf((size_t)n_config[q] * 5);
        for (long long i = 0; i < n_config[q]; ++i) {
            cf[(size_t)i * 5 + 0] = conf_n[q][2 * i]; cf[(size_t)i * 5 + 1] = conf_n[q][2 * i + 1];
            cf[(size_t)i * 5 + 2] = conf_l[q][2 * i]; cf[(size_t)i * 5 + 3] = conf_l[q][2 * i + 1];
            cf[(size_t)i * 5 + 4] = conf_eqv[q][i] ? 1 : 0;
        }
        w.raw_record(cf.data(), 8 * cf.size());
    }
    w.close();
}

// store_bsplines (bspline_tools.f90:381-385)
void write_splines(const std::string& path, long long k, long long n_knots, const double* knots)
{
    Writer w(path);
    const int64_t a = k, b = n_knots;
    w.raw_record(&a, 8); w.raw_record(&b, 8);
    w.raw_record(knots, 8 * (size_t)n_knots);
    w.close();
}

// ---- reader (CS_block_diag_load, block_tools.f90:487-524) --------------------------
struct Reader::Impl { FILE* f = nullptr; std::string path; };

Reader::Reader(const std::string& path) : impl(new Impl())
{
    impl->path = path;
    impl->f = fopen(path.c_str(), "rb");
    if (!impl->f) {
        delete impl;
        throw std::runtime_error("cannot open " + path);
    }
}
Reader::~Reader()
{
    if (impl->f) fclose(impl->f);
    delete impl;
}
// next record (all its subrecords) appended to out
void Reader::next_record(std::vector<char>& out)
{
    out.clear();
    bool more = true, first = true;
    while (more) {
        int32_t lead = 0, trail = 0;
        if (fread(&lead, 4, 1, impl->f) != 1) throw std::runtime_error(impl->path + ": unexpected end of file");
        more = lead < 0;
        const size_t len = (size_t)(lead < 0 ? -(long long)lead : lead);
        const size_t at = out.size();
        out.resize(at + len);
        if (len && fread(out.data() + at, 1, len, impl->f) != len) throw std::runtime_error(impl->path + ": truncated record");
        if (fread(&trail, 4, 1, impl->f) != 1) throw std::runtime_error(impl->path + ": truncated record marker");
        const long long want = first ? (long long)len : -(long long)len;
        if (trail != want) throw std::runtime_error(impl->path + ": record markers do not match");
        first = false;
    }
}

}  // namespace files
}  // namespace bs2e
