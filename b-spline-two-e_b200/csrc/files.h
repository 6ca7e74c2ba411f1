// files.h -- Fortran unformatted sequential result files of basis_setup (see files.cpp)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bs2e {
namespace files {

// test hook: split records into subrecords of at most this many bytes (default 2^31-9)
void set_max_subrecord(long long bytes);

class Writer {
public:
    explicit Writer(const std::string& path);
    ~Writer();
    Writer(const Writer&) = delete;
    Writer& operator=(const Writer&) = delete;
    void raw_record(const void* p, size_t n);
    void block_diag_header(long long n_blocks, const int64_t* block_rows);
    void block_matrix_header(long long nbr, long long nbc, const int64_t* block_rows, const int64_t* block_cols);
    void csr_block(long long rows, long long cols, long long nnz, const int64_t* index_ptr,
                   const int64_t* indices, const double* data);
    void csr_block_fragments(long long rows, long long cols, int nfrag, const long long* frag_rows,
                             const int64_t* const* frag_ptr, const int64_t* const* frag_idx,
                             const double* const* frag_dat);
    void single_csr(long long rows, long long cols, long long nnz, const int64_t* index_ptr,
                    const int64_t* indices, const double* data);
    void close();

private:
    struct Impl;
    Impl* impl;
};

void write_basis(const std::string& path, long long max_l_1p, long long max_L, bool two_el, long long n_sym,
                 const int64_t* sym_l, const int64_t* sym_m, const int64_t* sym_pi, const int64_t* n_config,
                 const int64_t* const* conf_n, const int64_t* const* conf_l, const int64_t* const* conf_eqv);
void write_splines(const std::string& path, long long k, long long n_knots, const double* knots);

class Reader {
public:
    explicit Reader(const std::string& path);
    ~Reader();
    Reader(const Reader&) = delete;
    Reader& operator=(const Reader&) = delete;
    void next_record(std::vector<char>& out);

private:
    struct Impl;
    Impl* impl;
};

}  // namespace files
}  // namespace bs2e
