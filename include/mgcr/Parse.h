// include/mgcr/Parse.h -- drop-in for the reference's src/Parse.h / Parse.cpp: the text CRS format the sample operators ship
// in (header `nrow ncol nnz`; one line of nrow row offsets -- the final offset is implied; nnz lines `col (re,im)`:
// src/Parse.cpp:46-58, 65-91) and the MatrixMarket -> CRS-text converter (src/Parse.cpp:10-62).  Header-only here.
// The reference hard-codes the directory "../../data/sample_matrix/" (src/Parse.cpp:67); that stays the default and
// the environment variable MGCR_DATA_DIR overrides it.
#ifndef MGCR_DROPIN_PARSE_H
#define MGCR_DROPIN_PARSE_H

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "Operator.h"

inline std::string mgcr_data_dir() {
    const char* d = std::getenv("MGCR_DATA_DIR");
    std::string s = d ? d : "../../data/sample_matrix/";
    if (!s.empty() && s.back() != '/') s += '/';
    return s;
}

// CRS text -> Sparse<long> (the arrays are malloc'd and adopted by the returned operator)
inline Sparse<long> read_data(const std::string& filename) {
    const std::string path = mgcr_data_dir() + filename;
    std::FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) {
        std::printf("File read is unsuccessful!\n");   // the reference prints and carries on (src/Parse.cpp:68-69)
        return Sparse<long>((long)0, (long)0, (long)0);
    }
    std::printf("File read is successful.\n");
    std::fseek(fp, 0, SEEK_END);
    const long bytes = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<char> buf((size_t)bytes + 1);
    const size_t got = std::fread(buf.data(), 1, (size_t)bytes, fp);
    std::fclose(fp);
    buf[got] = 0;
    char* p = buf.data();
    const long nrow = std::strtol(p, &p, 10), ncol = std::strtol(p, &p, 10), nnz = std::strtol(p, &p, 10);
    Sparse<long> out(nrow, ncol, nnz);
    for (long r = 0; r < nrow; r++) out.mod_ROW_at(r, std::strtol(p, &p, 10));
    for (long l = 0; l < nnz; l++) {
        const long c = std::strtol(p, &p, 10);
        while (*p && *p != '(') p++;
        p++;
        const double re = std::strtod(p, &p);
        while (*p && *p != ',') p++;
        p++;
        const double im = std::strtod(p, &p);
        while (*p && *p != ')') p++;
        if (*p) p++;
        out.mod_COL_at(l, c);
        out.mod_VAL_at(l, std::complex<double>(re, im));
    }
    return out;
}

// MatrixMarket coordinate file (1-based, complex; duplicates summed) -> "<data dir>/parsed.txt" in CRS text
// (src/Parse.cpp:10-62: header, the nrow row offsets on one line, then one "col (re,im)" line per entry)
inline void parse_data(const std::string& file_loc) {
    std::ifstream in(file_loc);
    if (in) std::printf("File read is successful.\n");
    else { std::printf("File read is unsuccessful!\n"); return; }
    while (in.peek() == '%') in.ignore(1 << 20, '\n');
    long nrow = 0, ncol = 0, n = 0;
    in >> nrow >> ncol >> n;
    typedef std::pair<std::complex<double>, std::pair<long, long>> Triplet;
    std::vector<Triplet> t((size_t)n);
    for (long l = 0; l < n; l++) {
        long r, c; double re, im;
        in >> r >> c >> re >> im;
        t[(size_t)l] = Triplet(std::complex<double>(re, im), std::pair<long, long>(r - 1, c - 1));
    }
    Sparse<long> sparse(nrow, ncol, t.data(), n);
    std::ofstream out(mgcr_data_dir() + "parsed.txt");
    out << sparse.get_nrow() << " " << sparse.get_dim() << " " << sparse.get_nnz() << "\n";
    for (long r = 0; r < nrow; r++) out << sparse.get_ROW(r) << " ";
    for (long l = 0; l < sparse.get_nnz(); l++) out << "\n" << sparse.get_COL(l) << " " << sparse.val_at(l);
}

#endif  // MGCR_DROPIN_PARSE_H
