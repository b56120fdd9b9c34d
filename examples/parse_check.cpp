// examples/parse_check.cpp -- the reference's I/O pair through the drop-in headers, host only (no device is touched):
//   parse_data(<MatrixMarket file>)   src/Parse.cpp:10-62   -> "<MGCR_DATA_DIR>/parsed.txt" in the CRS text format
//   read_data("parsed.txt")           src/Parse.cpp:65-91   -> Sparse<long>
// and dumps the arrays read_data produced as raw binary (row.bin / col.bin / val.bin next to parsed.txt) for
// tests/test_parse.py, which holds the text and the arrays to the unmodified reference's output on the same file.
//
//   MGCR_DATA_DIR=<dir> ./parse_check <dir>/in.mtx
#include <cstdio>
#include <string>

#include "Parse.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: parse_check <file.mtx>\n"); return 1; }
    parse_data(argv[1]);
    Sparse<long> m = read_data("parsed.txt");
    const std::string dir = mgcr_data_dir();
    auto dump = [&](const char* name, const void* p, size_t bytes) {
        std::FILE* f = std::fopen((dir + name).c_str(), "wb");
        std::fwrite(p, 1, bytes, f);
        std::fclose(f);
    };
    const long nrow = m.get_nrow(), nnz = m.get_nnz();
    std::vector<long> row((size_t)nrow + 1), col((size_t)nnz);
    std::vector<std::complex<double>> val((size_t)nnz);
    for (long r = 0; r <= nrow; r++) row[(size_t)r] = m.get_ROW(r);
    for (long l = 0; l < nnz; l++) { col[(size_t)l] = m.get_COL(l); val[(size_t)l] = m.val_at(l); }
    dump("row.bin", row.data(), 8 * row.size());
    dump("col.bin", col.data(), 8 * col.size());
    dump("val.bin", val.data(), 16 * val.size());
    std::printf("PARSED %ld %ld %ld\n", nrow, (long)m.get_dim(), nnz);
    return 0;
}
