// dataset.hpp -- text inputs of the reference: PLINK .fam, optional phenotype file, covariate file,
// and the .bed payload.  Parsing rules follow Data::read_fam / read_y / read_e / read_g
// (src/data.cpp:141-273): whitespace-separated token streams, records with unreadable or NaN fields
// are skipped, the (FID, IID) pair of the .fam file defines the row order, and a final count check
// raises the reference's error messages.
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace bmg {

struct Dataset {
  size_t n = 0, m_g = 0, m_e = 0;   // m_e INCLUDES the constant column (data.hpp:50)
  std::vector<double> y;            // n
  std::vector<double> e;            // n x m_e column-major, column 0 = 1
  std::vector<uint8_t> bed;         // ceil(n/4) * m_g payload bytes (header stripped)

  Dataset(size_t n_, size_t m_g_, size_t m_e_cov, const std::string& fam, const std::string& bedfile,
          const std::string& efile, const std::string& yfile, bool load_bed = true)
  : n(n_), m_g(m_g_), m_e(m_e_cov + 1), y(n_, 0.0), e(n_ * (m_e_cov + 1), 1.0)
  {
    read_fam(fam);
    if (!yfile.empty()) read_y(yfile);
    if (m_e > 1) read_e(efile);
    if (load_bed) read_bed(bedfile);
  }

 private:
  std::map<std::pair<std::string, std::string>, size_t> row_of_;

  void read_fam(const std::string& path)
  {
    std::ifstream f(path.c_str());
    size_t i = 0;
    while (f.good()) {
      std::string fid, iid, tmp;
      double ph = NAN;
      f >> fid >> iid >> tmp >> tmp >> tmp >> ph;
      if (f.good() && !fid.empty() && !iid.empty() && !std::isnan(ph)) {
        row_of_[std::make_pair(fid, iid)] = i;
        if (i < n) y[i] = ph;
        ++i;
      }
    }
    if (i != n || row_of_.size() != n) throw std::runtime_error("FAM file size does not match given n");
  }
  void read_y(const std::string& path)
  {
    std::ifstream f(path.c_str());
    size_t i = 0;
    while (f.good()) {
      std::string fid, iid;
      double ph = NAN;
      f >> fid >> iid >> ph;
      if (f.good() && !fid.empty() && !iid.empty() && !std::isnan(ph)) {
        auto it = row_of_.find(std::make_pair(fid, iid));
        if (it != row_of_.end()) { y[it->second] = ph; ++i; }
      }
    }
    if (i != n) throw std::runtime_error("Alternate phenotype file does not contain all phenotypes");
  }
  void read_e(const std::string& path)
  {
    std::ifstream f(path.c_str());
    size_t i = 0;
    const size_t nc = m_e - 1;
    while (f.good()) {
      std::string fid, iid;
      std::vector<double> row(nc, NAN);
      f >> fid >> iid;
      for (size_t j = 0; j < nc; ++j) f >> row[j];
      if (f.good() && !fid.empty() && !iid.empty()) {
        bool bad = false;
        for (size_t j = 0; j < nc; ++j) bad = bad || std::isnan(row[j]);
        if (bad) continue;
        auto it = row_of_.find(std::make_pair(fid, iid));
        if (it != row_of_.end()) {
          for (size_t j = 0; j < nc; ++j) e[(j + 1) * n + it->second] = row[j];
          ++i;
        }
      }
    }
    if (i != n) throw std::runtime_error("Covariate file does not contain all covariates");
  }
  void read_bed(const std::string& path)
  {
    std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
    if (!f.good()) throw std::runtime_error("BED file could not be opened");
    char b = 0;
    f.read(&b, 1);
    if (b != 0x6C) throw std::runtime_error("BED file not recognised (magic number does not match)");
    f.read(&b, 1);
    if (b != 0x1B) throw std::runtime_error("BED file not recognised (magic number does not match)");
    f.read(&b, 1);
    if (b != 0x01) throw std::runtime_error("BED file not in snp-major format");
    const size_t len = ((n + 3) / 4) * m_g;
    bed.resize(len);
    f.read(reinterpret_cast<char*>(bed.data()), (std::streamsize)len);
    if (!f.good() || (size_t)f.gcount() != len) throw std::runtime_error("Reading the BED file failed");
  }
};

}  // namespace bmg
